"""BASELINE config C4 (scaled down): R mixed synthetic regions (P ~ U{32..256}, 1-8 true components, S/N 10/30/100)
fitted with ``fit_regions`` - sharded over the ranks when launched under torchrun.
    python tools/bench_regions.py [R] [chains_per_restart]      (or torchrun --nproc-per-node N ...)
Prints regions/s (max over ranks) and the K2 / K3 split of rank 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vamp_b200 import engine, synth
from vamp_b200.fit_regions import fit_regions

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cpr = int(sys.argv[2]) if len(sys.argv) > 2 else 8
world, rank, lrank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
# workload generation first (host processes are forked before this process touches CUDA)
if R >= 4096:
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(16, max(1, (os.cpu_count() or 8) // max(world, 1)))) as pool:
        regs = pool.map(synth.config_c4_region, range(10_000, 10_000 + R), chunksize=256)
else:
    regs = [synth.config_c4_region(10_000 + i) for i in range(R)]
torch.cuda.set_device(lrank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
ctx = engine.default_context(lrank)
host = [(r["nu"], r["flux"], r["noise"]) for r in regs]
ntrue = np.array([len(r["truth"]) for r in regs])
import gc
for rep in range(int(os.environ.get("REPS", 3))):
    if os.environ.get("GCOFF"):
        gc.collect(); gc.disable()
    st = {}
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    res = fit_regions(host, uids=[10_000 + i for i in range(R)], voigt=True, iterations=3000, burn=300, thin=15, restarts=3,
                      chains_per_restart=cpr, window=eval(os.environ.get("WINDOW", "3")), nmax=8, seed=rep, stats=st,
                      polish_dtype=os.environ.get("POLISH", "auto"))
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        tm = st.get("timing", {})
        print("   fit_models: K2 %.0f ms, polish %.0f ms (device); host set-up %.3f s, launch %.3f s, wait %.3f s, post %.3f s" % (
            tm.get("k2_ms", 0), tm.get("polish_ms", 0), tm.get("host_setup_s", 0), tm.get("host_launch_s", 0),
            tm.get("host_wait_s", 0), tm.get("host_post_s", 0)), flush=True)
        print("   fit_local: upload %.3f s, ladder %.3f s, records %.3f s, close %.3f s" % (
            tm.get("upload_s", 0), tm.get("ladder_s", 0), tm.get("records_s", 0), tm.get("close_s", 0)), flush=True)
        nfit = np.array([r.n for r in res])
        print("rep %d: %d regions on %d GPU(s): %.2f s -> %.1f regions/s | rounds %d chains %d | checksum %s | "
              "n == truth %.2f, |n - truth| <= 1 %.2f, status0 %.2f" % (
                  rep, R, world, dt, R / dt, st.get("launches", 0) // 4, st.get("chains", 0), res.checksum()[:12],
                  np.mean(nfit == ntrue), np.mean(np.abs(nfit - ntrue) <= 1), np.mean([r.status == 0 for r in res])), flush=True)
if world > 1:
    dist.destroy_process_group()
