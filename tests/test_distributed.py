"""CPU (gloo, world_size 2): the sharding / gather logic of fit_regions with a fake per-shard
fitter - LPT partition, fixed-geometry records, uid ordering, independence of world size."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from vamp_b200 import batch, fit_regions as FR


def test_lpt_partition_balanced_and_deterministic():
    costs = [5, 9, 1, 7, 3, 3, 8, 2]
    parts = FR.lpt_partition(costs, 3)
    assert sorted(sum(parts, [])) == list(range(8))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(costs)
    assert parts == FR.lpt_partition(costs, 3)
    assert FR.lpt_partition(costs, 1) == [list(range(8))]


def test_record_roundtrip():
    tables = {2: (np.array([-10., -11., -12.]), np.array([3., 3., 3.])), 3: (np.array([-20., -21., -22.]), np.array([1., 1., 1.]))}
    th = np.arange(12.).reshape(3, 4)
    rec = FR.pack_record(77, 0, 2, 3, 3, tables, 4, 3, 7, th, 0.02, th + 1, th + 2, 0.021, 0.001, -5.5)
    assert rec.shape == (FR._record_len(4, 3, 7),)
    r = FR.unpack_record(rec, 4, 3, 7)
    assert (r.uid, r.status, r.n0, r.n, r.exit_n) == (77, 0, 2, 3, 3)
    np.testing.assert_array_equal(r.theta, th)
    np.testing.assert_array_equal(r.bic[3], [-20, -21, -22])
    assert r.lnL == -5.5 and r.sd == 0.02


def test_ladder_replay_matches_oracle_rule():
    from oracle import region as OR
    rng = np.random.default_rng(0)
    for trial in range(200):
        n0 = int(rng.integers(1, 4))
        tables = {}
        b = 0.0
        for n in range(n0, n0 + 8):
            b += rng.normal(-5, 10)
            tables[n] = (b + rng.normal(0, 2, 3), rng.uniform(0.5, 4.0, 3))
        kept_o, exit_o, _ = OR.region_fit_ladder(n0, lambda n: tables[n], chi_limit=1.5, n_cap=n0 + 7)
        kept, exit_n, decided = batch.ladder_replay(n0, tables, 1.5)
        if decided:
            assert (kept, exit_n) == (kept_o, exit_o), trial
        # partial tables: never decides something different from the full replay
        part = {n: tables[n] for n in range(n0, n0 + 3)}
        k2, e2, d2 = batch.ladder_replay(n0, part, 1.5)
        if d2:
            assert (k2, e2) == (kept_o, exit_o)


def _fake_fit(regions_host, uids, n0=None, cmax_rec=None, nmax=12, restarts=3, multiplet=None, **kw):
    """Stand-in for the GPU fit: a deterministic function of (uid, data) only (and of the region's multiplet
    entry, which must arrive sliced like the regions: it is written to ``sd_mean``)."""
    out = np.empty((len(uids), FR._record_len(nmax, restarts, cmax_rec)))
    assert multiplet is None or len(multiplet) == len(uids)
    for i, uid in enumerate(uids):
        nu, flux, noise = regions_host[i]
        kept = n0[i] + (uid % 2)
        tables = {n0[i] + j: (np.full(restarts, -10.0 * j - uid), np.full(restarts, 2.0 / (j + 1))) for j in range(2)}
        th = np.tile(np.array([flux.mean(), nu[0], noise[0], float(uid)]), (kept, 1))
        ar = 0.0 if (multiplet is None or multiplet[i] is None) else float(multiplet[i][1])
        out[i] = FR.pack_record(uid, 0, n0[i], kept, kept, tables, nmax, restarts, cmax_rec, th, 0.05, th, th * 0.1,
                                0.05 if multiplet is None else ar, 0.001, -float(uid))
    return out


def _regions(R):
    rng = np.random.default_rng(5)
    regs = []
    for i in range(R):
        P = int(rng.integers(20, 200))
        nu = 1e15 + np.arange(P) * 4e10
        regs.append((nu, rng.uniform(0, 1, P), np.full(P, 0.01 * (i + 1))))
    return regs


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    regs = _regions(11)
    res = FR.fit_regions(regs, uids=list(range(100, 111)), fit_fn=_fake_fit, n0=[1 + (i % 3) for i in range(11)])
    # per-region multiplet table: every rank's shard (and every device batch of it) gets ITS regions' entries
    mult = [None if i % 4 == 0 else (0.998, 0.1 * i, 1.0) for i in range(11)]
    resm = FR.fit_regions(regs, uids=list(range(100, 111)), fit_fn=_fake_fit, n0=[1 + (i % 3) for i in range(11)],
                          multiplet=mult, max_batch_regions=2)
    assert [r.sd_mean for r in resm] == [0.0 if m is None else m[1] for m in mult]
    # no n0 given: every rank counts the minima of its slice, one all-gather makes the counts global
    auto = FR.fit_regions(regs, uids=list(range(100, 111)), fit_fn=_fake_fit)
    q.put((rank, [(r.uid, r.n, float(r.theta[0, 0]), float(r.lnL), sorted(r.bic)) for r in res], auto.checksum(),
           [int(v) for v in auto.records[:, 2]]))
    dist.destroy_process_group()


def test_fit_regions_gloo_world2_equals_single_process():
    regs = _regions(11)
    n0 = [1 + (i % 3) for i in range(11)]
    single = FR.fit_regions(regs, uids=list(range(100, 111)), fit_fn=_fake_fit, n0=n0, distributed=False)
    ref = [(r.uid, r.n, float(r.theta[0, 0]), float(r.lnL), sorted(r.bic)) for r in single]
    assert [r[0] for r in ref] == list(range(100, 111))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows = [q.get(timeout=120) for _ in range(2)]
    got = {r[0]: r[1] for r in rows}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == ref and got[1] == ref
    # the sliced estimate_n + gather gives the single-process counts, and the records' checksum is world-size independent
    auto1 = FR.fit_regions(regs, uids=list(range(100, 111)), fit_fn=_fake_fit, distributed=False)
    assert [r[3] for r in rows] == [[FR.estimate_n(r[1]) for r in regs]] * 2
    assert {r[2] for r in rows} == {auto1.checksum()}


def test_emulated_ranks_reassemble_the_single_process_result():
    """``emulate=(rank, world)``: the records a rank of a W-rank run computes; their union is the 1-rank result."""
    regs = _regions(23)
    uids = list(range(500, 523))
    one = FR.fit_regions(regs, uids=uids, fit_fn=_fake_fit, distributed=False)
    for world in (2, 3, 8):
        parts = [FR.fit_regions(regs, uids=uids, fit_fn=_fake_fit, emulate=(r, world)).records for r in range(world)]
        assert sum(len(p) for p in parts) == 23
        allr = np.concatenate(parts)
        allr = allr[np.argsort(allr[:, 0])]
        assert FR.RegionResults(allr, one.nmax, one.G, one.cmax).checksum() == one.checksum()


def test_region_costs_vectorised_equals_scalar():
    rng = np.random.default_rng(1)
    P, n0 = rng.integers(17, 257, 50), rng.integers(1, 9, 50)
    for voigt in (False, True):
        for w in (1, 3, 4):
            np.testing.assert_allclose(FR.region_costs(P, n0, w, voigt), [FR.region_cost(p, n, w, voigt) for p, n in zip(P, n0)])


def test_chain_uids_unique_and_world_independent():
    a = batch.chain_uids(12345, 3, 3, 32)
    b = batch.chain_uids(12345, 4, 3, 32)
    c = batch.chain_uids(12346, 3, 3, 32)
    allu = np.concatenate([a, b, c])
    assert len(np.unique(allu)) == allu.size


def test_model_chain_uids_equal_the_per_model_definition():
    rng = np.random.default_rng(0)
    region_uids = rng.integers(0, 2 ** 39, 50)
    models = [(int(rng.integers(0, 50)), int(rng.integers(1, 18))) for _ in range(400)]
    want = np.concatenate([batch.chain_uids(region_uids[r], n, 3, 8) for r, n in models])
    got = batch.model_chain_uids(region_uids, models, 3, 8)
    assert got.dtype == np.uint64 and np.array_equal(got, want)


def test_empty_inputs_need_no_device():
    """No regions (an empty catalogue, or a rank whose shard is empty): an empty result, no GPU touched."""
    assert FR.fit_regions([], distributed=False) == []
    out = FR.fit_local([], [], cmax_rec=5, nmax=4, restarts=3)
    assert out.shape == (0, FR._record_len(4, 3, 5))


def test_chunked_batches_equal_single_batch():
    """A shard above ``max_batch_regions`` is fitted in consecutive device batches; the records are the same."""
    regs = _regions(11)
    n0 = [1 + (i % 3) for i in range(11)]
    calls = []

    def counting(regions_host, uids, **kw):
        calls.append(len(regions_host))
        return _fake_fit(regions_host, uids, **kw)

    one = FR.fit_regions(regs, uids=list(range(100, 111)), fit_fn=_fake_fit, n0=n0, distributed=False)
    many = FR.fit_regions(regs, uids=list(range(100, 111)), fit_fn=counting, n0=n0, distributed=False, max_batch_regions=4)
    assert calls == [4, 4, 3]
    for a, b in zip(one, many):
        assert a.uid == b.uid and a.n == b.n
        np.testing.assert_array_equal(a.theta, b.theta)


def test_multiplet_table_follows_the_regions():
    """One tuple for all regions or one per region; sliced with the regions into device batches."""
    regs = _regions(7)
    n0 = [1] * 7
    one = FR.fit_regions(regs, fit_fn=_fake_fit, n0=n0, distributed=False, multiplet=(0.998, 0.5, 1.0), max_batch_regions=3)
    assert [r.sd_mean for r in one] == [0.5] * 7
    per = [(0.998, 0.1 * (i + 1)) if i % 2 else None for i in range(7)]
    res = FR.fit_regions(regs, fit_fn=_fake_fit, n0=n0, distributed=False, multiplet=per, max_batch_regions=3)
    assert [r.sd_mean for r in res] == [0.0 if m is None else m[1] for m in per]
    with pytest.raises(ValueError):
        FR.fit_regions(regs, fit_fn=_fake_fit, n0=n0, distributed=False, multiplet=[(0.9, 0.5)] * 3)
    # without n0 the absorber count of a doublet region comes from the significant minima
    from vamp_b200 import synth
    d = synth.make_doublet_region(256, 2, 60.0, 901)
    got = {}

    def spy(regions_host, uids, n0=None, **kw):
        got["n0"] = list(n0)
        return _fake_fit(regions_host, uids, n0=n0, **kw)

    FR.fit_regions([(d["nu"], d["flux"], d["noise"])] * 2, fit_fn=spy, distributed=False, multiplet=[d["multiplet"], None])
    assert got["n0"][0] == 2 and got["n0"][1] == FR.estimate_n(d["flux"]) >= 4
