"""Build libvamp_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library).

Each kernel file is compiled once per (precision, profile kind) so the objects build in
parallel (the sampler once more for its multiplet instantiations); `python -m vamp_b200.build` or `__graft_entry__.build()` runs this.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvamp_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
KERNEL_FILES = ("vamp_k1", "vamp_k2", "vamp_k3")
HEADERS = ("vamp_device.cuh", "vamp_kernels.h", "faddeeva_tables.inc")


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: cannot build libvamp_b200.so")


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _jobs():
    hdr = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(HERE, "..", "include", "vamp_b200.h")]
    jobs = []
    for base in KERNEL_FILES:
        for f64 in (0, 1):
            for voigt in (0, 1):
                src = os.path.join(CSRC, base + ".cu")
                obj = os.path.join(OBJ, "%s_%s_%s.o" % (base, "f64" if f64 else "f32", "voigt" if voigt else "gauss"))
                defs = ["-DVAMP_INST_F64=%d" % f64, "-DVAMP_INST_VOIGT=%d" % voigt]
                jobs.append((src, obj, defs, hdr))
                if base == "vamp_k2":       # the sampler's multiplet instantiations (two tied lines per component)
                    jobs.append((src, obj[:-2] + "_mult.o", defs + ["-DVAMP_INST_MULT=1"], hdr))
                    # ... and its adaptive-Metropolis instantiations (step_kind = VAMP_STEP_ADAPTIVE)
                    jobs.append((src, obj[:-2] + "_am.o", defs + ["-DVAMP_INST_AM=1"], hdr))
                    # ... and its pair-tail instantiations (regions with an odd number of 64-pixel row pairs)
                    jobs.append((src, obj[:-2] + "_tail.o", defs + ["-DVAMP_INST_TAIL=1"], hdr))
                    # ... and its long-region instantiations (257 .. 2048 pixels, one chain per CTA)
                    jobs.append((src, obj[:-2] + "_long.o", defs + ["-DVAMP_INST_LONG=1"], hdr))
    jobs.append((os.path.join(CSRC, "vamp_k4.cu"), os.path.join(OBJ, "vamp_k4.o"), [], hdr))
    jobs.append((os.path.join(CSRC, "vamp_cabi.cu"), os.path.join(OBJ, "vamp_cabi.o"), [], hdr))
    return jobs


def _compile(job, nvcc, verbose):
    src, obj, defs, hdr = job
    stamp = obj + ".sha"
    dig = _digest([src] + hdr, " ".join(ARCH + FLAGS + defs))
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False, ""
    cmd = [nvcc] + ARCH + FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    with open(stamp, "w") as fh:
        fh.write(dig)
    return obj, True, r.stderr


def build(force=False, verbose=False, jobs=None):
    """Compile (if stale) and link vamp_b200/libvamp_b200.so.  Returns the library path."""
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    work = _jobs()
    nthreads = jobs or min(len(work), os.cpu_count() or 4)
    with ThreadPoolExecutor(nthreads) as ex:
        res = list(ex.map(lambda j: _compile(j, nvcc, verbose), work))
    objs = [r[0] for r in res]
    if verbose:
        for r in res:
            if r[2]:
                sys.stderr.write(r[2])
    if any(r[1] for r in res) or not os.path.exists(LIB):
        cmd = [nvcc] + ARCH + ["-shared", "--cudart", "shared", "-Xcompiler", "-fPIC", "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
