"""Experiment (measured and left off): K2 with its warp-uniform branches marked ``bra.uni``.

ptxas cannot prove that the branches of the sampler's step loop are warp-uniform (the chain index comes from
``threadIdx.x >> 5``), so it guards every vote and shuffle with ``BRA.DIV`` / ``WARPSYNC`` pairs.  This tool rebuilds the
float32 Voigt K2 object with a PTX rewrite between cicc and ptxas: every conditional branch whose ``.loc`` line is in
``WHITELIST`` (lines of vamp_k2.cu / vamp_device.cuh whose conditions are warp-uniform by construction - no lane index,
no lane-strided loop) becomes ``bra.uni``.  Static SASS 4096 -> 3488 instructions, no BRA.DIV left; measured on B200:
115.6 -> 114.9 ms per 1000 iterations of the C2 chain set, identical chains.  Not part of the build.

    python tools/ptx_uniform.py            # -> tools/variants/b_uni.so   (time it with tools/k2_variants.py run)

The whitelist holds LINE NUMBERS; the tool checks the SHA-256 of the two sources it was derived for and refuses any
other version (``--list`` prints every conditional branch location of the kernels to re-derive it).
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
WHITELIST = {"vamp_k2.cu": {357, 430, 540, 585, 597, 614, 658, 692, 693, 697, 699, 701, 733, 763, 764, 767, 782, 783, 786, 789,
                            793, 801, 820, 848, 871, 908, 911, 933, 935, 1472, 1475, 1476, 1477, 1499},
             "vamp_device.cuh": {684}}
# the sources the line numbers above belong to: the tool refuses any other (a stale line could mark a divergent branch)
SOURCE_SHA256 = {"vamp_k2.cu": "de4f80f2ea70295fe18e624d3a05e1c794320af446c7dcd69aaaf289db6cbeac",
                 "vamp_device.cuh": "b338939e4b04261f69fab737dbc34d1058aa65031379af0a9aac81b743b921c8"}


def patch(ptx_in, ptx_out, list_only=False):
    src = open(ptx_in).read().split("\n")
    files = {}
    for line in src:
        m = re.match(r'\s*\.file\s+(\d+)\s+"([^"]+)"', line)
        if m:
            files[m.group(1)] = os.path.basename(m.group(2))
    out, cur, n, inside, seen = [], None, 0, False, {}
    for line in src:
        if line.startswith(".entry"):
            inside = "k2_mcmc_kernel" in line
        m = re.match(r"\s*\.loc\s+(\d+)\s+(\d+)", line)
        if m:
            cur = (files.get(m.group(1)), int(m.group(2)))
        if inside and cur and re.match(r"\s*@!?%p\d+\s+bra\s+\$L", line):
            seen[cur] = seen.get(cur, 0) + 1
            if cur[1] in WHITELIST.get(cur[0], ()):
                line = line.replace("bra ", "bra.uni ", 1)
                n += 1
        out.append(line)
    if list_only:
        for k in sorted(seen):
            print("%-18s %5d  x%d%s" % (k[0], k[1], seen[k], "  [uni]" if k[1] in WHITELIST.get(k[0], ()) else ""))
        return n
    open(ptx_out, "w").write("\n".join(out))
    return n


def main():
    import hashlib
    from vamp_b200 import build as B
    for name, want in SOURCE_SHA256.items():
        got = hashlib.sha256(open(os.path.join(B.CSRC, name), "rb").read()).hexdigest()
        if got != want and "--list" not in sys.argv:
            raise SystemExit("%s changed since the whitelist was derived: re-derive WHITELIST (--list) and SOURCE_SHA256" % name)
    B.build()
    nvcc = B._nvcc()
    keep = os.path.join(ROOT, "tools", "variants", "_uni")
    os.makedirs(keep, exist_ok=True)
    cu = os.path.join(B.CSRC, "vamp_k2.cu")
    defs = ["-DVAMP_INST_F64=0", "-DVAMP_INST_VOIGT=1"]
    # the whole pipeline once (intermediates kept), then ptxas / fatbinary / host compile again on the patched PTX
    subprocess.run([nvcc, "--keep", "--keep-dir", keep] + B.ARCH + B.FLAGS + defs + ["-c", cu, "-o", os.path.join(keep, "plain.o")],
                   check=True, capture_output=True)
    ptx = os.path.join(keep, "vamp_k2.ptx")
    if "--list" in sys.argv:
        patch(ptx, None, list_only=True)
        return
    n = patch(ptx, os.path.join(keep, "vamp_k2_uni.ptx"))
    cuda = os.path.dirname(os.path.dirname(nvcc))
    subprocess.run([os.path.join(cuda, "bin", "ptxas"), "-arch=sm_100a", "-m64", "--generate-line-info",
                    os.path.join(keep, "vamp_k2_uni.ptx"), "-o", os.path.join(keep, "vamp_k2.cubin")], check=True)
    subprocess.run([os.path.join(cuda, "bin", "fatbinary"), "-64", "--image3=kind=elf,sm=100a,file=" + os.path.join(keep, "vamp_k2.cubin"),
                    "--embedded-fatbin=" + os.path.join(keep, "vamp_k2.fatbin.c")], check=True)
    obj = os.path.join(keep, "k2_uni.o")
    subprocess.run(["gcc", "-std=c++17", "-D__CUDA_ARCH__=1000", "-D__CUDA_ARCH_FEAT_SM100_ALL", "-D__CUDA_ARCH_SPECIFIC__=1000",
                    "-D__CUDA_ARCH_FAMILY_SPECIFIC__=1000", "-D__CUDA_ARCH_LIST__=1000", "-D__NV_LEGACY_LAUNCH", "-c", "-x", "c++",
                    "-DCUDA_DOUBLE_MATH_FUNCTIONS", "-fPIC", "-O3", "-Wno-psabi", "-I" + os.path.join(cuda, "targets", "x86_64-linux", "include"),
                    "-m64", os.path.join(keep, "vamp_k2.cudafe1.cpp"), "-o", obj], check=True)
    objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith(".o") and f != "vamp_k2_f32_voigt.o"]
    so = os.path.join(ROOT, "tools", "variants", "b_uni.so")
    subprocess.run([nvcc] + B.ARCH + ["-shared", "--cudart", "shared", "-Xcompiler", "-fPIC", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
                                      "-o", so, obj] + objs, check=True)
    print("%d branches marked uniform -> %s" % (n, so))


if __name__ == "__main__":
    main()
