"""Build K2 variants (extra -D flags on vamp_k2.cu, f32 Voigt object only) into gpurun_out/variants/*.so and,
on a GPU box, time each one:  python tools/k2_variants.py build "NAME:-DX=1 -DY=2" ... ; python tools/k2_variants.py run [iters]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "tools", "variants")

def build(specs):
    from vamp_b200 import build as B
    B.build()
    os.makedirs(VDIR, exist_ok=True)
    nvcc = B._nvcc()
    objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith(".o") and f != "vamp_k2_f32_voigt.o"]
    for spec in specs:
        name, flags = spec.split(":", 1)
        obj = os.path.join(VDIR, name + ".o")
        cmd = [nvcc] + B.ARCH + B.FLAGS + ["-DVAMP_INST_F64=0", "-DVAMP_INST_VOIGT=1", "-Xptxas", "-v"] + flags.split() + ["-c", os.path.join(B.CSRC, "vamp_k2.cu"), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode: raise SystemExit(r.stderr[-3000:])
        info = [l for l in r.stderr.splitlines() if "Used" in l or "spill" in l]
        print(name, "|", info[5].strip() if len(info) > 5 else "", "|", info[4].strip() if len(info) > 4 else "")
        so = os.path.join(VDIR, name + ".so")
        r = subprocess.run([nvcc] + B.ARCH + ["-shared", "--cudart", "shared", "-Xcompiler", "-fPIC", "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", so, obj] + objs, capture_output=True, text=True)
        if r.returncode: raise SystemExit(r.stderr[-3000:])
        os.remove(obj)

def run(iters):
    names = sorted(f[:-3] for f in os.listdir(VDIR) if f.endswith(".so"))
    for rep in range(2):
        for n in names:
            env = dict(os.environ, VAMP_B200_LIB=os.path.join(VDIR, n + ".so"))
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "k2_time.py"), str(iters)], env=env, capture_output=True, text=True)
            print("%-28s %s" % (n, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1]), flush=True)

if __name__ == "__main__":
    if sys.argv[1] == "build": build(sys.argv[2:])
    else: run(int(sys.argv[2]) if len(sys.argv) > 2 else 1000)
