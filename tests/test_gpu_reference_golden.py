"""GPU parity against outputs of THE REFERENCE'S OWN SOURCES: tests/golden/reference_build.npz holds what the
reference's PnPsolver.cpp / Sim3Solver.cpp -- compiled unmodified by `make -C oracle ref` against the stand-in headers of
oracle/shim/ -- return on stored inputs (scripts/make_reference_golden.py; the CPU suite checks the oracle against the
same file and against the live library).  Here the CUDA engine is compared with the file directly, through the C ABI:
PnP in the reference's own structure (RSAC_FLAG_EPNP_EIGEN: 12 x 12 eigen-solve per hypothesis), exhaustive and with
the staged early exit, and Sim3.  Bit for bit: return value, inlier count, iteration at which the reference stopped,
inlier vector, pose."""
import os

import numpy as np
import pytest

from ransac_b200 import capi

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_build.npz"))


@pytest.mark.parametrize("early_exit", [False, True])
def test_pnp_engine_equals_compiled_reference(engine, early_exit):
    C, n = G["pnp_p3d"].shape[:2]
    ls2 = G["level_sigma2"]
    sigma2 = ls2[G["pnp_octave"]]
    offsets = np.arange(C + 1, dtype=np.int32) * n
    pr = G["pnp_params"]
    prm = capi.ransac_params(pr[0], int(pr[1]), int(pr[2]), int(pr[3]), float(pr[4]), float(pr[5]))
    flags = capi.FLAG_EPNP_EIGEN | (capi.FLAG_EARLY_EXIT if early_exit else 0)
    res, masks = engine.pnp_solve(offsets, G["pnp_p3d"], G["pnp_p2d"], sigma2, [G["pnp_K"]], prm, seeds=G["pnp_seeds"], flags=flags)
    ml = engine.split_masks(masks, offsets)
    for c in range(C):
        r = res[c]
        assert (bool(r["ok"]), bool(r["no_more"]), int(r["n_inliers"]), int(r["best_count"]), int(r["n_refines"])) == \
               (bool(G["pnp_ok"][c]), bool(G["pnp_no_more"][c]), int(G["pnp_n_inliers"][c]), int(G["pnp_best_inliers"][c]), int(G["pnp_n_refines"][c])), c
        assert int(r["n_hyp"]) == int(G["pnp_iterations"][c]), (c, int(r["n_hyp"]), int(G["pnp_iterations"][c]))
        T = G["pnp_T"][c]
        assert (r["R"].reshape(3, 3).view(np.uint32) == T[:3, :3].view(np.uint32)).all(), c
        assert (r["t"].view(np.uint32) == T[:3, 3].view(np.uint32)).all(), c
        assert (ml[c] == G["pnp_inliers"][c]).all(), c


def test_sim3_engine_equals_compiled_reference(engine):
    C, n = G["sim3_x1c"].shape[:2]
    ls2 = G["level_sigma2"]
    offsets = np.arange(C + 1, dtype=np.int32) * n
    sp = G["sim3_params"]
    K = np.array([G["sim3_K"]], np.float32)
    res, masks = engine.sim3_solve(offsets, G["sim3_x1c"], G["sim3_x2c"], ls2[G["sim3_oct1"]], ls2[G["sim3_oct2"]], K, K,
                                   capi.Sim3Params(sp[0], int(sp[1]), int(sp[2]), 1), seeds=G["sim3_seeds"])
    ml = engine.split_masks(masks, offsets)
    for c in range(C):
        r = res[c]
        assert (bool(r["ok"]), int(r["n_inliers"])) == (bool(G["sim3_ok"][c]), int(G["sim3_n_inliers"][c])), c
        assert int(r["n_hyp"]) == int(G["sim3_iterations"][c]), (c, int(r["n_hyp"]), int(G["sim3_iterations"][c]))
        if r["ok"]:
            assert (r["R"].reshape(3, 3).view(np.uint32) == G["sim3_R"][c].view(np.uint32)).all(), c
            assert (r["t"].view(np.uint32) == G["sim3_t"][c].view(np.uint32)).all(), c
            assert (ml[c] == G["sim3_inliers"][c]).all(), c


def test_mlpnp_engine_equals_compiled_reference(engine):
    """MLPnPsolver as shipped (Refine() discards its pose, quirk Q6 => RSAC_FLAG_MLPNP_DISCARD_REFINE); the reference's
    MLPnPsolver.cpp calls libm and Eigen SVDs, so: return value, inlier count, stopping iteration exact; pose 1e-6"""
    C, n = G["mlpnp_p3d"].shape[:2]
    ls2 = G["level_sigma2"]
    offsets = np.arange(C + 1, dtype=np.int32) * n
    pr = G["mlpnp_params"]
    prm = capi.ransac_params(pr[0], int(pr[1]), int(pr[2]), int(pr[3]), float(pr[4]), float(pr[5]))
    for flags in (capi.FLAG_MLPNP_DISCARD_REFINE, capi.FLAG_MLPNP_DISCARD_REFINE | capi.FLAG_EARLY_EXIT):
        res, masks = engine.mlpnp_solve(offsets, G["mlpnp_p3d"], G["mlpnp_p2d"], ls2[G["mlpnp_octave"]], np.array([G["pnp_K"]], np.float32), prm,
                                        seeds=G["mlpnp_seeds"], flags=flags)
        ml = engine.split_masks(masks, offsets)
        for c in range(C):
            r = res[c]
            assert (bool(r["ok"]), int(r["n_inliers"]), int(r["best_count"]), int(r["n_hyp"])) == \
                   (bool(G["mlpnp_ok"][c]), int(G["mlpnp_n_inliers"][c]), int(G["mlpnp_best_inliers"][c]), int(G["mlpnp_iterations"][c])), (c, flags)
            T = G["mlpnp_T"][c]
            assert np.abs(r["R"].reshape(3, 3) - T[:3, :3]).max() < 1e-6 and np.abs(r["t"] - T[:3, 3]).max() < 1e-6 * max(1.0, np.abs(T[:3, 3]).max()), c
            assert (ml[c] != G["mlpnp_inliers"][c]).sum() <= 1, c


# ---- retrieval and matching: tests/golden/reference_build_matching.npz holds the compiled reference's
# outputs for seeded generator cases (scripts/make_reference_golden.py::matching_cases) and a SHA-1 of the generated inputs
def _matching():
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(root, "scripts", "make_reference_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(root, "tests", "golden", "reference_build_matching.npz"))
    cases = mod.matching_cases()
    dig = mod.matching_digests(cases)
    for k, v in dig.items():
        assert str(g["sha1_" + k]) == v, "the seeded generator no longer produces the inputs the golden outputs belong to: " + k
    return g, cases


def test_matching_and_retrieval_engine_equals_compiled_reference(engine):
    """ORBmatcher::SearchByBoW (both overloads), SearchBySim3, SearchByProjection(Frame, KeyFrame) and
    KeyFrameDatabase::DetectRelocalizationCandidates: the engine against outputs of the reference's own ORBmatcher.cpp /
    KeyFrameDatabase.cpp, element for element"""
    g, cases = _matching()
    F, kfs = cases["bow0"]()
    m, n = engine.bow_match([F] + kfs, list(range(1, 7)), [0] * 6, 0.75, True, 0)
    for i in range(6):
        assert n[i] == g["bow0_n"][i] and (m[i] == g["bow0_match"][i]).all(), ("bow0", i)
    cur, kfs1 = cases["bow1"]()
    m, n = engine.bow_match([cur] + kfs1, [0] * 4, list(range(1, 5)), 0.75, True, 1)
    for i in range(4):
        assert n[i] == g["bow1_n"][i] and (m[i] == g["bow1_match"][i]).all(), ("bow1", i)
    prs = cases["pairs"]()
    views = [v for p in prs for v in (p["kf1"], p["kf2"])]
    got, nf = engine.sim3_search(views, [0, 2, 4], [1, 3, 5], [p["K"] for p in prs], [p["R12"] for p in prs], [p["t12"] for p in prs], 7.5,
                                 [p["matched12_in"] for p in prs])
    for i in range(3):
        assert nf[i] == int(g["sim3s_n_%d" % i]) and got[i].tolist() == g["sim3s_match_%d" % i].tolist(), ("sim3 search", i)
    pj = cases["proj"]()
    views = [v for c in pj for v in (c["frame"], c["kf"])]
    got, nm, fell, rounds = engine.proj_search(views, [0, 2, 4], [1, 3, 5], [c["K"] for c in pj], [c["Rcw"] for c in pj], [c["tcw"] for c in pj],
                                               10.0, 100, True, [c["occupied"] for c in pj], [c["already_found"] for c in pj])
    for i in range(3):
        assert nm[i] == int(g["proj_n_%d" % i]) and got[i].tolist() == g["proj_match_%d" % i].tolist(), ("projection search", i)
    db, qs = cases["kfdb"]()
    engine.kfdb_upload(db)
    got = engine.kfdb_detect(qs, mode=0)
    for q in range(len(qs)):
        assert got[q].tolist() == g["kfdb_cand_%d" % q].tolist(), ("retrieval", q)
    assert (engine.kfdb_state().view(np.uint32) == g["kfdb_state"].view(np.uint32)).all()
