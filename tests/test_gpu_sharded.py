"""GPU parity of the sharded engine entry points (gl_ns_shard_*): G CUDA shards emulated in one
process on one device (LocalComm) against numpy oracle shards run through the same SPMD loop."""
import numpy as np
import pytest

from gleipnir_b200.sharded import LocalComm, run_sharded, shard_bounds
from oracle import priors as pr
from oracle.likelihoods import Model
from oracle.sharded import OracleShard

pytestmark = pytest.mark.gpu


def _run(world, n, k, deaths, table, dev_like, ora_model, skw, seed=21, sigmas=8.0):
    from gleipnir_b200.engine import DeviceEngine
    dev, ora = [], []
    for r in range(world):
        off, n_loc = shard_bounds(n, world, r)
        dkw = dict(skw)
        dkw["proposal"] = 1 if dkw.get("proposal") == "normal" else 0
        dev.append(DeviceEngine(table, dev_like, n_loc, k, deaths, seed=seed, row_offset=off,
                                global_population=n, **dkw))
        ora.append(OracleShard(table, ora_model, n_loc, batch_k=k, seed=seed, row_offset=off, global_population=n,
                               **skw))
    run_sharded(dev, LocalComm(world), deaths, k, candidate_sigmas=sigmas)
    run_sharded(ora, LocalComm(world), deaths, k, candidate_sigmas=0.0)      # ground truth: full exchange
    return dev, ora


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_device_run_matches_oracle_shards(cuda, world):
    from gleipnir_b200 import likelihoods as lk
    n, k, deaths = 400, 40, 1200
    table = [(pr.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * 2
    dev, ora = _run(world, n, k, deaths, table, lk.EggCartonLogLikelihood(2), Model("eggcarton", 2),
                    dict(iterations=6, tuning_cycles=1))
    states = [e.state() for e in dev]
    assert len({(s.log_z, s.information, s.log_x, s.n_dead) for s in states}) == 1      # redundant & identical
    total_rows = 0
    for e, o, s in zip(dev, ora, states):
        tab, idx, step = e.dead_table()
        want = o.result
        nd = len(want["dead_index"])
        # the device table is in death order, the oracle shard's in slot order: compare per batch
        od = np.lexsort((idx[:nd], step[:nd]))
        oo = np.lexsort((want["dead_index"], want["dead_batch"]))
        assert np.array_equal(idx[:nd][od], want["dead_index"][oo])
        assert np.array_equal(step[:nd][od], want["dead_batch"][oo])
        np.testing.assert_allclose(tab[:nd][od], want["dead"][:nd][oo], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(tab[nd:], want["dead"][nd:], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(s.log_z, want["logZ"], rtol=1e-12)
        assert s.n_evals == want["n_evals"]
        total_rows += tab.shape[0]
    assert total_rows == deaths + (n - k)
    w = np.concatenate([e.dead_table()[0][:, 1] for e in dev])
    assert abs(w.sum() - 1.0) < 1e-12


def test_world_of_one_sharded_equals_plain_engine(cuda):
    from gleipnir_b200 import likelihoods as lk
    from gleipnir_b200.engine import DeviceEngine
    n, k, deaths = 300, 30, 900
    table = [(pr.GL_PRIOR_UNIFORM, -5.0, 10.0)] * 3
    like = lk.IIDGaussianLogLikelihood(3)
    a = DeviceEngine(table, like, n, k, deaths, seed=4, iterations=5, tuning_cycles=1)
    run_sharded([a], LocalComm(1), deaths, k)
    b = DeviceEngine(table, like, n, k, deaths, seed=4, iterations=5, tuning_cycles=1)
    b.init(); b.advance(deaths); b.finalize()
    ta, tb = a.dead_table(), b.dead_table()
    assert np.array_equal(ta[0], tb[0]) and np.array_equal(ta[1], tb[1])
    assert a.state().log_z == b.state().log_z and a.state().information == b.state().information


def test_sharded_logz_is_unbiased(cuda):
    """4 shards, rank-local seeding: 2-D Gaussian evidence over 8 seeds."""
    from gleipnir_b200 import likelihoods as lk
    from gleipnir_b200.engine import DeviceEngine
    n, k, deaths, world = 2000, 200, 20000, 4
    table = [(pr.GL_PRIOR_UNIFORM, -5.0, 10.0)] * 2
    like = lk.IIDGaussianLogLikelihood(2)
    zs = []
    for seed in range(8):
        dev = []
        for r in range(world):
            off, n_loc = shard_bounds(n, world, r)
            dev.append(DeviceEngine(table, like, n_loc, k, deaths, seed=seed, row_offset=off,
                                    global_population=n, iterations=10, tuning_cycles=1))
        run_sharded(dev, LocalComm(world), deaths, k)
        zs.append(dev[0].state().log_z)
    assert abs(np.mean(zs) - (-4.60517)) < 3 * 0.03 / np.sqrt(8) + 0.01, zs


def test_large_gathered_set_uses_global_key_scratch(cuda):
    """world*k = 28 000 > 24 576: the fused commit keeps its keys in a global scratch; results must
    equal the oracle shards exactly as in the shared-memory case."""
    from gleipnir_b200 import likelihoods as lk
    n, k, deaths, world = 32000, 7000, 14000, 4
    table = [(pr.GL_PRIOR_UNIFORM, -5.0, 10.0)] * 2
    dev, ora = _run(world, n, k, deaths, table, lk.IIDGaussianLogLikelihood(2), Model("iid_gaussian", 2),
                    dict(iterations=2, tuning_cycles=0), sigmas=0.0)
    for e, o in zip(dev, ora):
        tab, idx, step = e.dead_table()
        want = o.result
        nd = len(want["dead_index"])
        od = np.lexsort((idx[:nd], step[:nd]))
        oo = np.lexsort((want["dead_index"], want["dead_batch"]))
        assert np.array_equal(idx[:nd][od], want["dead_index"][oo])
        np.testing.assert_allclose(e.state().log_z, want["logZ"], rtol=1e-12)


@pytest.mark.parametrize("multipass", [False, True])
def test_truncated_exchange_matches_full_exchange_oracle(cuda, multipass):
    """m < k candidates per rank (k = 400 over 4 ranks: m = 178): the committed selection is
    proven exact on the device and equals the oracle's full exchange bit for bit."""
    from gleipnir_b200 import likelihoods as lk
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.sharded import _run_sharded_once, candidate_count
    n, k, deaths, world = 4000, 400, 2400, 4
    assert candidate_count(k, world, n // world, 8.0) == 178
    table = [(pr.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * 2
    skw = dict(iterations=4, tuning_cycles=1)
    dev, ora = [], []
    for r in range(world):
        off, n_loc = shard_bounds(n, world, r)
        e = DeviceEngine(table, lk.EggCartonLogLikelihood(2), n_loc, k, deaths, seed=5, row_offset=off,
                         global_population=n, **skw)
        if multipass:
            e.force_multipass_select(True)
        dev.append(e)
        ora.append(OracleShard(table, Model("eggcarton", 2), n_loc, batch_k=k, seed=5, row_offset=off,
                               global_population=n, **skw))
    _run_sharded_once(dev, LocalComm(world), deaths, k, 8.0)
    assert not any(e.overflowed() for e in dev)
    _run_sharded_once(ora, LocalComm(world), deaths, k, 0.0)
    for e, o in zip(dev, ora):
        tab, idx, step = e.dead_table()
        want = o.result
        nd = len(want["dead_index"])
        od = np.lexsort((idx[:nd], step[:nd]))
        oo = np.lexsort((want["dead_index"], want["dead_batch"]))
        assert np.array_equal(idx[:nd][od], want["dead_index"][oo])
        np.testing.assert_allclose(tab[:nd][od], want["dead"][:nd][oo], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(e.state().log_z, want["logZ"], rtol=1e-12)


@pytest.mark.parametrize("multipass", [False, True])
def test_truncated_exchange_overflow_is_detected_and_repeated(cuda, multipass):
    """A margin of ~0 sigmas cannot cover the binomial spread of deaths over the ranks: the device
    raises shard_overflow (same verdict on every rank) and run_sharded repeats the run in full."""
    from gleipnir_b200 import likelihoods as lk
    from gleipnir_b200.engine import DeviceEngine
    from gleipnir_b200.sharded import _run_sharded_once
    n, k, deaths, world = 4000, 1200, 6000, 2
    table = [(pr.GL_PRIOR_UNIFORM, -5.0, 10.0)] * 2
    skw = dict(iterations=3, tuning_cycles=0)

    def engines():
        out = []
        for r in range(world):
            off, n_loc = shard_bounds(n, world, r)
            e = DeviceEngine(table, lk.IIDGaussianLogLikelihood(2), n_loc, k, deaths, seed=9, row_offset=off,
                             global_population=n, **skw)
            if multipass:
                e.force_multipass_select(True)
            out.append(e)
        return out
    probe = engines()
    _run_sharded_once(probe, LocalComm(world), deaths, k, 1e-9)
    assert all(e.overflowed() for e in probe)
    auto, full = engines(), engines()
    run_sharded(auto, LocalComm(world), deaths, k, candidate_sigmas=1e-9)
    run_sharded(full, LocalComm(world), deaths, k, candidate_sigmas=0.0)
    for a, b in zip(auto, full):
        ta, tb = a.dead_table(), b.dead_table()
        assert np.array_equal(ta[0], tb[0]) and np.array_equal(ta[1], tb[1])
        assert a.state().log_z == b.state().log_z and not a.overflowed()


@pytest.mark.parametrize("world", [2, 4])
def test_gathered_table_assembled_on_the_device_matches_host_ordering(cuda, world):
    """gl_export_gathered_table (sharded run(): one kernel instead of a host lexsort) against
    engine.death_order applied to the same gathered rows."""
    import torch
    from gleipnir_b200 import _lib
    from gleipnir_b200 import likelihoods as lk
    from gleipnir_b200.engine import DeviceEngine, death_order
    n, k, deaths = 2000, 300, 2000           # short last batch (2000 = 6*300 + 200)
    table = [(pr.GL_PRIOR_UNIFORM, 0.0, 10.0 * np.pi)] * 2
    dev = []
    for r in range(world):
        off, n_loc = shard_bounds(n, world, r)
        dev.append(DeviceEngine(table, lk.EggCartonLogLikelihood(2), n_loc, k, deaths, seed=3, row_offset=off,
                                global_population=n, iterations=3, tuning_cycles=1))
    run_sharded(dev, LocalComm(world), deaths, k)
    d = 2
    counts = [int(e.state().n_rows) for e in dev]
    mx = max(counts)
    packed = torch.zeros((world, mx, d + 4), dtype=torch.float64, device=dev[0].device)
    for r, e in enumerate(dev):
        c = counts[r]
        packed[r, :c, 0] = e.dead_logl[:c]
        packed[r, :c, 1] = e.dead_w[:c]
        packed[r, :c, 2:2 + d] = e.dead_x[:c]
        packed[r, :c, 2 + d] = e.dead_idx[:c].to(torch.float64)
        packed[r, :c, 3 + d] = e.dead_step[:c].to(torch.float64)
    total = sum(counts)
    out = torch.full((total, d + 2), float("nan"), dtype=torch.float64, device=dev[0].device)
    cnt = torch.tensor(counts, dtype=torch.int64, device=dev[0].device)
    n_steps = (deaths + k - 1) // k
    lib = _lib.load()
    _lib.check(lib.gl_export_gathered_table(packed.data_ptr(), cnt.data_ptr(), world, mx, d, k, n_steps, deaths,
                                            out.data_ptr(), None))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    rows = np.concatenate([packed[r, :counts[r]].cpu().numpy() for r in range(world)])
    order = death_order(rows[:, 0], rows[:, d + 2].astype(np.int64), rows[:, d + 3].astype(np.int64))
    want = rows[order][:, :d + 2]
    assert total == deaths + (n - 200)
    assert np.array_equal(got, want)
    assert abs(got[:, 1].sum() - 1.0) < 1e-12
