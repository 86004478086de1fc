"""GPU parity: batched Sim3Solver (CUDA, through the C ABI) vs the CPU oracle.

All Sim3 arithmetic is FP32 with the reference's operation order on both sides, so hypotheses,
counts, masks and the returned R/t/s are compared bit for bit (north_star tolerance would be
1e-4 relative on R/t/s)."""
import numpy as np
import pytest

from ransac_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _problems(seeds, n, outl, scale):
    return [synth.sim3_problem(int(s), n, outl, scale) for s in seeds]


def _cat(ps, key):
    return np.concatenate([p[key] for p in ps])


def _run(engine, oracle, ps, prm, seeds, fix_scale):
    sizes = [p["x1c"].shape[0] for p in ps]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    K = np.array([ps[0]["K"]], np.float32)
    res, masks = engine.sim3_solve(offsets, _cat(ps, "x1c"), _cat(ps, "x2c"), _cat(ps, "sigma2_1"), _cat(ps, "sigma2_2"),
                                   K, K, capi.Sim3Params(prm[0], prm[1], prm[2], 1 if fix_scale else 0), seeds=np.asarray(seeds, np.uint32))
    ml = engine.split_masks(masks, offsets)
    hyp_words = 0
    Hs = []
    for n in sizes:
        H = capi.sim3_ransac_setup(n, capi.Sim3Params(prm[0], prm[1], prm[2], 1)) if n >= max(prm[1], 3) else 0
        Hs.append(H)
        hyp_words += H * ((n + 31) // 32)
    poses, counts, hm = engine.sim3_hypotheses(hyp_words)
    h0 = 0
    for c, p in enumerate(ps):
        n, H = sizes[c], Hs[c]
        pb = oracle.sim3_problem(p["x1c"], p["x2c"], p["sigma2_1"], p["sigma2_2"], p["K"], p["K"], fix_scale=fix_scale)
        if H == 0:
            assert res[c]["ok"] == 0 and res[c]["no_more"] == 1
            continue
        tab = oracle.index_table(int(seeds[c]), n, 3, H)
        o = oracle.sim3_ransac(pb, prm[0], prm[1], prm[2], tab, oracle.FLAG_EXHAUSTIVE, per_hyp=True)
        gp, op = poses[h0:h0 + H], o["hyp_pose"]
        same = (gp.view(np.uint32) == op.view(np.uint32)) | (np.isnan(gp) & np.isnan(op))
        assert same.all(), f"problem {c}: hypotheses {np.argwhere(~same.all(axis=1)).ravel()[:5]} differ"
        assert (counts[h0:h0 + H] == o["hyp_counts"]).all(), c
        r = res[c]
        for k in ("ok", "no_more", "n_inliers", "best_hyp", "best_count"):
            assert r[k] == o[k], (c, k, r[k], o[k])
        assert (ml[c] == o["mask"]).all(), c
        assert (r["R"].reshape(3, 3).view(np.uint32) == o["T"][:3, :3].view(np.uint32)).all()
        assert (r["t"].view(np.uint32) == o["T"][:3, 3].view(np.uint32)).all()
        assert np.float32(r["s"]) == np.float32(o["scale"])
        h0 += H
    return res


def test_sim3_cfg3_fixed_scale(engine, oracle):
    """cfg3: N=200, 40% outliers, SetRansacParameters(0.99,20,300) => H=300 (LoopClosing.cpp:261)"""
    seeds = [3000 + i for i in range(6)]
    ps = _problems(seeds, 200, 0.4, 1.0)
    res = _run(engine, oracle, ps, (0.99, 20, 300), seeds, True)
    assert res["ok"].sum() >= 5
    for c, p in enumerate(ps):
        if res[c]["ok"]:
            assert np.abs(res[c]["R"].reshape(3, 3) - p["R12"]).max() < 0.05
            assert np.abs(res[c]["t"] - p["t12"]).max() < 0.3


def test_sim3_monocular_scale(engine, oracle):
    """free scale (Horn's scale step, upstream ORB-SLAM2; absent from the reference, parity unpinned)"""
    seeds = [3100 + i for i in range(4)]
    ps = _problems(seeds, 200, 0.4, 1.6)
    res = _run(engine, oracle, ps, (0.99, 20, 300), seeds, False)
    ok = res[res["ok"] == 1]
    assert len(ok) >= 3 and np.all(np.abs(ok["s"] - 1.6) < 0.1)


def test_sim3_ragged_and_failures(engine, oracle):
    """ragged sizes incl. empty / n < minInliers, and an all-outlier set that must exhaust its budget"""
    seeds = [3200 + i for i in range(7)]
    sizes = [0, 2, 19, 20, 33, 129, 200]
    outl = [0.0, 0.0, 0.0, 0.3, 1.0, 0.5, 0.95]
    ps = []
    for s, n, o in zip(seeds, sizes, outl):
        p = synth.sim3_problem(s, max(n, 1), o, 1.0)
        for k in ("x1c", "x2c", "sigma2_1", "sigma2_2"):
            p[k] = p[k][:n]
        ps.append(p)
    res = _run(engine, oracle, ps, (0.99, 20, 300), seeds, True)
    assert res[4]["ok"] == 0 and res[4]["no_more"] == 1
