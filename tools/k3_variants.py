"""K3w variants: python tools/k3_variants.py build "NAME:-Dflags" ... ; python tools/k3_variants.py run"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "tools", "variants")
def build(specs):
    from vamp_b200 import build as B
    B.build(); os.makedirs(VDIR, exist_ok=True); nvcc = B._nvcc()
    objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith(".o") and f != "vamp_k3_f32_voigt.o"]
    for spec in specs:
        name, flags = spec.split(":", 1)
        obj = os.path.join(VDIR, name + ".o")
        r = subprocess.run([nvcc] + B.ARCH + B.FLAGS + ["-DVAMP_INST_F64=0", "-DVAMP_INST_VOIGT=1", "-Xptxas", "-v"] + flags.split() + ["-c", os.path.join(B.CSRC, "vamp_k3.cu"), "-o", obj], capture_output=True, text=True)
        if r.returncode: raise SystemExit(r.stderr[-3000:])
        print(name, [l.strip() for l in r.stderr.splitlines() if "Used" in l][:1])
        so = os.path.join(VDIR, name + ".so")
        r = subprocess.run([nvcc] + B.ARCH + ["-shared", "--cudart", "shared", "-Xcompiler", "-fPIC", "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", so, obj] + objs, capture_output=True, text=True)
        if r.returncode: raise SystemExit(r.stderr[-3000:])
        os.remove(obj)
def run():
    for n in sorted(f[:-3] for f in os.listdir(VDIR) if f.endswith(".so")):
        env = dict(os.environ, VAMP_B200_LIB=os.path.join(VDIR, n + ".so"), POLISH="mixed")
        a = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "step_phases.py"), "3"], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
        b = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_regions.py"), "256", "8"], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
        print(n, "|", a[a.find("K3 "):a.find("K3 ") + 9], "|", b[b.find("K2 "):b.find("rounds")], flush=True)
if __name__ == "__main__":
    build(sys.argv[2:]) if sys.argv[1] == "build" else run()
