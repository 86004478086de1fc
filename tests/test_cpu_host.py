"""CPU suite, part 2: host logic of the product and the C-ABI library itself (no compute calls
that need a GPU): every symbol declared in include/ransac_b200.h is exported, the glibc rand()
restatement equals the real libc stream, the RANSAC parameter arithmetic equals the oracle's, the
device numerical core compiled for the host is bit-identical to the oracle, and engine creation
fails loudly without a CUDA device (there is no CPU fallback)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from ransac_b200 import capi, shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _has_gpu():
    return capi.lib().rsac_device_count() > 0


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "ransac_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(rsac_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 40
    L = capi.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, f"declared in include/ransac_b200.h but not exported: {missing}"
    assert L.rsac_version() == 100


def test_result_record_layout(built_lib):
    assert C.sizeof(capi.Result) == 96 and capi.RESULT_DTYPE.itemsize == 96
    assert shard.REC_WORDS == 24


def test_glibc_rand_restatement_equals_libc(built_lib):
    libc = C.CDLL("libc.so.6")
    for seed in (0, 1, 2, 7, 1000, 123456789, 2 ** 31 - 1, 2 ** 32 - 1):
        libc.srand(C.c_uint(seed))
        ref = np.array([libc.rand() for _ in range(3000)], np.int32)
        assert (capi.rand_stream(seed, 3000) == ref).all(), seed
    g = json.load(open(os.path.join(GOLD, "rng_known_answers.json")))
    for seed, vals in g["rand"].items():
        assert capi.rand_stream(int(seed), 16).tolist() == vals


def test_index_tables_equal_reference_idiom(built_lib, oracle):
    for (seed, n, k, H) in ((1, 500, 4, 300), (1000, 200, 3, 300), (5, 1000, 6, 300), (9, 7, 6, 50), (9, 4, 4, 10), (3, 3, 3, 5)):
        assert (capi.index_table(seed, n, k, H) == oracle.index_table(seed, n, k, H)).all()
    assert capi.index_table(1, 500, 4, 3).tolist() == [[420, 196, 389, 396], [455, 98, 166, 381], [138, 276, 237, 312]]
    with pytest.raises(capi.RsacError):
        capi.index_table(1, 3, 4, 1)          # k > n


def test_ransac_setup_equals_oracle(built_lib, oracle):
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(1, 3000))
        prob = float(rng.choice([0.9, 0.99, 0.999]))
        mi = int(rng.integers(1, 60))
        its = int(rng.integers(1, 400))
        ms = int(rng.choice([4, 6]))
        eps = float(rng.uniform(0.05, 0.9))
        a = capi.pnp_ransac_setup(n, capi.ransac_params(prob, mi, its, ms, eps, 5.991))
        b = oracle.ransac_setup_pnp(n, oracle.params(prob, mi, its, ms, eps, 5.991))
        assert a == b
        if n >= mi:
            assert capi.sim3_ransac_setup(n, capi.Sim3Params(prob, mi, its, 1)) == oracle.ransac_setup_sim3(n, prob, mi, its)


def test_shard_ranges_cover_exactly(built_lib):
    for Cn in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                f, c = capi.shard_range(Cn, r, world)
                assert (f, c) == shard.block_range(Cn, r, world)
                seen += list(range(f, f + c))
            assert seen == list(range(Cn))


def test_host_compiled_core_is_bit_identical_to_oracle(built_lib, oracle):
    """the solver source the kernels are built from, compiled for the host, vs the oracle"""
    L = capi.lib()
    p = synth.pnp_problem(1000, 500, 0.5)
    pb = oracle.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
    tab = oracle.index_table(1000, 500, 4, 120)
    K = (C.c_double * 4)(*p["K"])
    for h in range(120):
        idx = tab[h]
        Ro, to, _ = oracle.epnp_pose(pb, idx)
        p3, p2 = np.ascontiguousarray(p["p3d"][idx]), np.ascontiguousarray(p["p2d"][idx])
        R, t = np.empty(9, np.float32), np.empty(3, np.float32)
        L.rsac_debug_host_epnp4(K, capi._p(p3), capi._p(p2), capi._p(R), capi._p(t))
        assert np.array_equal(R.view(np.uint32), Ro.reshape(-1).view(np.uint32)) and np.array_equal(t.view(np.uint32), to.view(np.uint32)), h
        # default device mode: null space of the 4-point system by Householder QR
        Rq, tq, _ = oracle.epnp_pose(pb, idx, oracle.FLAG_EPNP_QR_NULLSPACE)
        L.rsac_debug_host_epnp4_qr(K, capi._p(p3), capi._p(p2), capi._p(R), capi._p(t))
        assert np.array_equal(R.view(np.uint32), Rq.reshape(-1).view(np.uint32)) and np.array_equal(t.view(np.uint32), tq.view(np.uint32)), h
    # 12x12 eigen-solve: 4 smallest eigenpairs
    rng = np.random.default_rng(1)
    for _ in range(20):
        M = rng.normal(size=(8, 12)) * 50
        A = np.ascontiguousarray(M.T @ M)
        w, v = np.empty(4), np.empty((12, 4))
        L.rsac_debug_host_jacobi12(capi._p(A), capi._p(w), capi._p(v))
        wo, vo = oracle.jacobi_lowest(A, 4)
        assert np.array_equal(w.view(np.uint64), wo.view(np.uint64)) and np.array_equal(v.view(np.uint64), vo.view(np.uint64))
    # Horn (f32)
    for it in range(50):
        P1 = (rng.normal(size=(3, 3)) * 3 + [0, 0, 8]).astype(np.float32)
        P2 = (rng.normal(size=(3, 3)) * 3 + [0, 0, 8]).astype(np.float32)
        for fs in (1, 0):
            Ro, to, so = oracle.sim3_compute(P1, P2, bool(fs))
            R, t, s = np.empty(9, np.float32), np.empty(3, np.float32), C.c_float()
            L.rsac_debug_host_sim3(capi._p(P1), capi._p(P2), fs, capi._p(R), capi._p(t), C.byref(s))
            assert np.array_equal(R.view(np.uint32), Ro.reshape(-1).view(np.uint32)) and np.array_equal(t.view(np.uint32), to.view(np.uint32))
            assert np.float32(s.value) == np.float32(so)
    # MLPnP n = 6 (same libm on the host => bit-identical)
    q = synth.pnp_problem(2000, 300, 0.5)
    Kf = np.array(q["K"], np.float32)
    cov = synth.bearing_covariances(q)
    for use_cov in (False, True):
        mb = oracle.mlpnp_problem(q["p3d"], q["p2d"], q["sigma2"], tuple(Kf), cov if use_cov else None)
        tab = oracle.index_table(2000, 300, 6, 40)
        for h in range(40):
            idx = tab[h]
            Ro, to = oracle.mlpnp_pose(mb, idx)
            p3, p2 = np.ascontiguousarray(q["p3d"][idx]), np.ascontiguousarray(q["p2d"][idx])
            cv = np.ascontiguousarray(cov[idx]) if use_cov else None
            R, t = np.empty(9), np.empty(3)
            L.rsac_debug_host_mlpnp6(capi._p(Kf), capi._p(p3), capi._p(p2), capi._p(cv), capi._p(R), capi._p(t))
            assert np.array_equal(R.view(np.uint64), Ro.reshape(-1).view(np.uint64)) and np.array_equal(t.view(np.uint64), to.view(np.uint64))


def test_host_compiled_poseopt_is_bit_identical_to_oracle(built_lib, oracle):
    """csrc/poseopt.cuh compiled for the host with one lane (edges summed in order) against oracle/orc_poseopt.c:
    separate sources, same operation sequence, same libm -> identical poses, flags and step-control trajectory"""
    from ransac_b200 import capi, synth

    for seed, n, outl, sr, pn in [(1, 250, 0.2, 0.0, (0.02, 0.05)), (2, 250, 0.3, 0.5, (0.02, 0.05)), (3, 400, 0.3, 1.0, (0.02, 0.05)),
                                  (4, 8, 0.2, 0.3, (0.02, 0.05)), (5, 2, 0.0, 0.0, (0.02, 0.05)), (6, 1200, 0.2, 0.2, (0.02, 0.05)),
                                  (7, 0, 0.0, 0.0, (0.02, 0.05)), (8, 200, 0.5, 0.0, (0.3, 1.0)), (9, 200, 0.5, 1.0, (0.3, 1.0))]:
        p = synth.poseopt_problem(seed, n, outl, sr, pose_noise=pn)
        d, out = oracle.pose_optimization(oracle.poseopt_problem(p["p3d"], p["obs"], p["inv_sigma2"], p["K"], p["Rcw"], p["tcw"]))
        r, o2 = capi.debug_host_poseopt(p["p3d"], p["obs"], p["inv_sigma2"], p["K"], np.concatenate([p["Rcw"].ravel(), p["tcw"]]))
        for k in ("n_inliers", "n_bad", "rounds", "iterations", "trials"):
            assert int(r[k]) == int(d[k]), (seed, k)
        assert np.array_equal(r["R"], d["R"].ravel()) and np.array_equal(r["t"], d["t"]), seed
        assert np.array_equal(r["Rf"], d["Rf"].ravel()) and np.array_equal(r["tf"], d["tf"]), seed
        assert np.array_equal(out, o2), seed


def test_no_gpu_means_loud_failure_not_fallback(built_lib):
    if _has_gpu():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert capi.lib().rsac_create(0, C.byref(h)) == capi.ERR_NO_DEVICE and not h.value
    with pytest.raises(capi.RsacError):
        capi.Engine(0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/ may reference it"""
    pkg = os.path.join(ROOT, "orb-slam2-optimized_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".hpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle_api" not in txt and "liboracle" not in txt and '"orc.h"' not in txt, os.path.join(root, f)


def test_synthetic_generator_is_seeded_and_shaped():
    a, b = synth.pnp_problem(1000, 500, 0.5), synth.pnp_problem(1000, 500, 0.5)
    assert np.array_equal(a["p3d"], b["p3d"]) and np.array_equal(a["p2d"], b["p2d"])
    assert a["p3d"].dtype == np.float32 and a["p3d"].shape == (500, 3) and a["inlier"].sum() == 250
    assert set(np.unique(a["sigma2"])) <= set(synth.level_sigma2().tolist())
    s = synth.level_sigma2()
    assert s[0] == 1.0 and abs(float(s[1]) - 1.44) < 1e-6
    q = synth.sim3_problem(3000, 200, 0.4, 1.6)
    assert q["x1c"].shape == (200, 3) and q["inlier"].sum() == 120 and q["s"] == 1.6


def test_host_compiled_sim3opt_is_bit_identical_to_oracle(built_lib, oracle):
    """csrc/sim3opt.cuh compiled for the host with one lane against oracle/orc_poseopt.c (OptimizeSim3 half)"""
    from ransac_b200 import capi, synth

    for seed, n, outl, fs, sc in [(1, 100, 0.15, True, 1.0), (2, 200, 0.3, True, 1.0), (3, 60, 0.0, True, 1.0), (4, 12, 0.5, True, 1.0),
                                  (5, 100, 0.15, False, 1.6), (6, 0, 0.0, True, 1.0), (7, 5, 0.0, True, 1.0)]:
        p = synth.sim3opt_problem(seed, n, outl, scale=sc)
        a = (p["x1c"], p["x2c"], p["obs1"], p["obs2"], p["inv_sigma2_1"], p["inv_sigma2_2"], p["K"], p["K"], p["S12"])
        d, rem = oracle.optimize_sim3(oracle.sim3opt_problem(*a, th2=10.0, fix_scale=fs))
        r, rem2 = capi.debug_host_sim3opt(*a, th2=10.0, fix_scale=fs)
        for k in ("n_inliers", "n_bad", "optimized", "iterations", "trials"):
            assert int(r[k]) == int(d[k]), (seed, k)
        assert np.array_equal(r["R"], d["R"].ravel()) and np.array_equal(r["t"], d["t"]) and r["s"] == d["s"], seed
        assert np.array_equal(rem, rem2), seed
