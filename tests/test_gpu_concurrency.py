"""Two handles on ONE device. The fused per-scan kernel's blocks poll for each other's partial rows, so two such
grids must never be half resident together; the library orders fused launches of different handles of a process
(lk_api.cu: FusedGate). These tests interleave two handles — asynchronously from one thread, and from two host
threads — and require every result to equal the single-handle result bit for bit (and, above all, to finish).
Also covers back-to-back launches of one stream with programmatic dependent launch on and off."""
import threading

import numpy as np
import pytest

import scenes
from legkilo_b200 import Engine, abi

pytestmark = pytest.mark.gpu


def _ring(n=6, stream0=4100):
    cfg, blob, scans = scenes.box_scene(batch=n, stream0=stream0)
    pts = np.concatenate(scans)
    offs = np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.uint32)
    return cfg, blob, scans, pts, offs


def _staged_engine(cfg, blob, pts, offs, n, **params):
    eng = Engine(cfg)
    for k, v in params.items():
        eng.set_param(k, v)
    eng.map_upload(blob)
    eng.stage(abi.default_states(n), abi.init_cov(n), abi.process_cov_Q(cfg), np.zeros(n, abi.CLOCK_DTYPE), pts, offs, np.zeros(n))
    return eng


def test_two_handles_interleaved_async_launches():
    n = 6
    cfg, blob, scans, pts, offs = _ring(n)
    ref_eng = _staged_engine(cfg, blob, pts, offs, n)
    for i in range(n):
        ref_eng.run_range(i, 1, iters=3)
    ref_eng.sync()
    ref = ref_eng.fetch()
    a = _staged_engine(cfg, blob, pts, offs, n)
    b = _staged_engine(cfg, blob, pts, offs, n)
    for rep in range(20):  # 240 launches alternating between two streams, no host sync in between
        for i in range(n):
            a.run_range(i, 1, iters=3)
            b.run_range((i + 3) % n, 1, iters=3)
    a.sync(); b.sync()
    for eng in (a, b):
        out = eng.fetch()
        assert out["x"].tobytes() == ref["x"].tobytes()
        assert out["P"].tobytes() == ref["P"].tobytes()
        assert np.array_equal(out["n_eff"], ref["n_eff"])
        np.testing.assert_array_equal(out["world"], ref["world"])


def test_two_handles_two_host_threads():
    n = 4
    cfg, blob, scans, pts, offs = _ring(n, stream0=4300)
    Q = abi.process_cov_Q(cfg)
    engs = [Engine(cfg), Engine(cfg)]
    for e in engs:
        e.map_upload(blob)
    x0 = abi.default_states(1); P0 = abi.init_cov(1); clk = np.zeros(1, abi.CLOCK_DTYPE)
    want = [engs[0].scan_update(x0, P0, Q, clk, s, [0, len(s)], [0.0], iters=3, pinned=True) for s in scans]
    got = [[None] * n, [None] * n]
    errs = []

    def worker(t):
        try:
            for rep in range(25):
                for i, s in enumerate(scans):
                    got[t][i] = engs[t].scan_update(x0, P0, Q, clk, s, [0, len(s)], [0.0], iters=3, pinned=(rep % 2 == 0))
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
        assert not t.is_alive(), "a fused launch never finished (two grids polling each other?)"
    assert not errs, errs
    for t in range(2):
        for i in range(n):
            assert got[t][i]["x"].tobytes() == want[i]["x"].tobytes()
            assert got[t][i]["P"].tobytes() == want[i]["P"].tobytes()


@pytest.mark.parametrize("params", [dict(pdl=0), dict(pdl=1), dict(coop_launch=1), dict(lane_cache=0)])
def test_back_to_back_launch_modes_bitwise(params):
    n = 6
    cfg, blob, scans, pts, offs = _ring(n, stream0=4500)
    ref_eng = _staged_engine(cfg, blob, pts, offs, n, fused=0)
    for i in range(n):
        ref_eng.run_range(i, 1, iters=3)
    ref_eng.sync()
    ref = ref_eng.fetch()
    eng = _staged_engine(cfg, blob, pts, offs, n, **params)
    for rep in range(30):
        for i in range(n):
            eng.run_range(i, 1, iters=3)
    eng.sync()
    out = eng.fetch()
    assert out["x"].tobytes() == ref["x"].tobytes()
    assert out["P"].tobytes() == ref["P"].tobytes()
    assert np.array_equal(out["n_eff"], ref["n_eff"]) and int(ref["n_eff"].min()) > 0
    np.testing.assert_array_equal(out["world"], ref["world"])
