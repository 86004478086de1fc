"""GPU: the C++ class API (include/ransac_b200/solvers.hpp: PnPsolver / MLPnPsolver / Sim3Solver with the
reference's method names) driven like Tracking::Relocalization and LoopClosing::ComputeSim3 drive the
reference, checked call by call against a replay built from oracle primitives."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from ransac_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "cpp", "class_driver")


@pytest.fixture(scope="module")
def driver(built_lib):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)
    return DRIVER


def _frame_bytes(p, n_slots, slot_of, rng):
    """a Frame with n_slots keypoints; correspondence i of the problem sits in slot slot_of[i]; the other
    slots have no (or a bad) MapPoint"""
    xy = rng.uniform(0, 700, size=(n_slots, 2)).astype(np.float32)
    octave = rng.integers(0, 8, size=n_slots).astype(np.int32)
    valid = np.zeros(n_slots, np.uint8)
    world = rng.normal(size=(n_slots, 3)).astype(np.float32)
    s2 = synth.level_sigma2()
    xy[slot_of] = p["p2d"]
    octave[slot_of] = p["octave"]
    valid[slot_of] = 1
    world[slot_of] = p["p3d"]
    K = np.array(p["K"], np.float32)
    return struct.pack("<i", n_slots) + K.tobytes() + xy.tobytes() + octave.tobytes() + valid.tobytes() + world.tobytes() + s2.tobytes()


def _run(driver, mode, blob, tmp_path):
    f = tmp_path / (mode + ".bin")
    f.write_bytes(blob)
    out = subprocess.run([driver, mode, str(f)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    return [json.loads(l) for l in out.stdout.splitlines() if l.strip()]


def _pnp_replay(oracle, p, prm, seed):
    """PnPsolver::iterate called repeatedly (PnPsolver.cpp:102-191), rebuilt from oracle primitives"""
    Kd = tuple(float(np.float32(k)) for k in p["K"])        # Frame::fx.. are floats, widened by the solver
    pb = oracle.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], Kd)
    oprm = oracle.params(*prm)
    n = len(p["p3d"])
    minInl, H = oracle.ransac_setup_pnp(n, oprm)
    tab = oracle.index_table(seed, n, 4, H)
    ex = oracle.pnp_ransac(pb, oprm, tab, oracle.FLAG_EXHAUSTIVE | oracle.FLAG_EPNP_QR_NULLSPACE, per_hyp=True)
    counts, poses = ex["hyp_counts"], ex["hyp_pose"]
    thr = (p["sigma2"] * np.float32(prm[5])).astype(np.float32)
    calls, best, bestmask, bestpose, h = [], 0, None, None, 0
    while True:
        ret = None
        while h < H:
            cur = h
            h += 1
            if counts[cur] >= minInl:
                if counts[cur] > best:
                    best = counts[cur]
                    bestpose = poses[cur]
                    _, bestmask, _ = oracle.pnp_check_inliers(pb, thr, poses[cur][:9], poses[cur][9:])
                R, t, _ = oracle.epnp_pose(pb, np.flatnonzero(bestmask))
                cnt, mask, _ = oracle.pnp_check_inliers(pb, thr, R, t)
                if cnt > minInl:
                    ret = dict(ok=1, noMore=0, n=cnt, R=R, t=t, mask=mask)
                    break
        if ret is None:
            if best >= minInl:
                ret = dict(ok=1, noMore=1, n=best, R=bestpose[:9].reshape(3, 3), t=bestpose[9:], mask=bestmask)
            else:
                ret = dict(ok=0, noMore=1, n=0, R=None, t=None, mask=None)
        calls.append(ret)
        if ret["noMore"] or len(calls) > 400:
            return calls, (minInl, H)


def test_pnpsolver_iterate_sequence(driver, oracle, tmp_path):
    rng = np.random.default_rng(0)
    for seed, n, outl, prm in ((1000, 500, 0.5, (0.99, 10, 300, 4, 0.2, 5.991)),        # cfg1
                               (12003, 200, 0.3, (0.99, 10, 300, 4, 0.5, 5.991))):      # Tracking.cpp:1226
        p = synth.pnp_problem(seed, n, outl)
        n_slots = n + 300
        slot_of = np.sort(rng.permutation(n_slots)[:n])
        blob = _frame_bytes(p, n_slots, slot_of, rng) + struct.pack("<diiiffIi", prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], seed, 5)
        got = _run(driver, "pnp", blob, tmp_path)
        ref, (minInl, H) = _pnp_replay(oracle, p, prm, seed)
        assert got[0] == {"H": H, "minInl": minInl, "N": n}
        got = got[1:]
        past = got.pop()                                  # the call after bNoMore: best-so-far fallback (PnPsolver.cpp:173-188)
        assert past["call"] == -1 and past["noMore"] == 1
        assert (past["ok"], past["nInliers"]) == (ref[-1]["ok"], ref[-1]["n"])      # ref[-1] is the exhausted call: the same fallback
        assert len(got) == len(ref)
        for g, r in zip(got, ref):
            assert (g["ok"], g["noMore"], g["nInliers"]) == (r["ok"], r["noMore"], r["n"])
            if r["ok"]:
                T = np.array(g["T"], np.float32).reshape(4, 4)
                assert np.allclose(T[:3, :3], r["R"], rtol=1e-4, atol=1e-6) and np.allclose(T[:3, 3], r["t"], rtol=1e-4, atol=1e-6)
                assert (T[3] == [0, 0, 0, 1]).all()
                assert g["inliers"] == slot_of[np.flatnonzero(r["mask"])].tolist()       # scattered to keypoint indices
        assert len(ref) >= 2                                                             # the resumed path was exercised


def test_pnpsolver_batch_entry_point(driver, oracle, tmp_path):
    rng = np.random.default_rng(1)
    C, n = 5, 300
    blob = struct.pack("<i", C)
    ps, slots = [], []
    for c in range(C):
        p = synth.pnp_problem(4000 + c, n, 0.5)
        slot_of = np.sort(rng.permutation(n + 50)[:n])
        blob += _frame_bytes(p, n + 50, slot_of, rng) + struct.pack("<I", 4000 + c)
        ps.append(p); slots.append(slot_of)
    got = _run(driver, "pnp_batch", blob, tmp_path)
    for c in range(C):
        ref, _ = _pnp_replay(oracle, ps[c], (0.99, 10, 300, 4, 0.2, 5.991), 4000 + c)
        g, r = got[c], ref[0]
        assert (g["ok"], g["nInliers"]) == (r["ok"], r["n"])
        if r["ok"]:
            assert g["inliers"] == slots[c][np.flatnonzero(r["mask"])].tolist()


def test_mlpnpsolver_iterate(driver, oracle, tmp_path):
    rng = np.random.default_rng(2)
    seed, n = 2000, 400
    p = synth.pnp_problem(seed, n, 0.5)
    slot_of = np.sort(rng.permutation(n + 100)[:n])
    prm = (0.99, 10, 300, 6, 0.2, 5.991)
    blob = _frame_bytes(p, n + 100, slot_of, rng) + struct.pack("<diiiffIi", *prm, seed, 5)
    got = _run(driver, "mlpnp", blob, tmp_path)
    Kf = tuple(np.float32(k) for k in p["K"])
    pb = oracle.mlpnp_problem(p["p3d"], p["p2d"], p["sigma2"], Kf)
    _, H = oracle.ransac_setup_pnp(n, oracle.params(*prm))
    o = oracle.mlpnp_ransac(pb, oracle.params(*prm), oracle.index_table(seed, n, 6, H))
    g = got[0]
    assert (g["ok"], g["noMore"], g["nInliers"]) == (o["ok"], o["no_more"], o["n_inliers"])
    T = np.array(g["T"], np.float32).reshape(4, 4)
    assert np.allclose(T, o["T"], rtol=1e-4, atol=1e-6)
    assert g["inliers"] == slot_of[np.flatnonzero(o["mask"])].tolist()


@pytest.mark.parametrize("scale", [1.0, 1.6])
def test_sim3solver_iterate_five_at_a_time(driver, oracle, tmp_path, scale):
    """LoopClosing.cpp:286 calls iterate(5, ...) until it returns true or bNoMore"""
    rng = np.random.default_rng(3)
    seed, n = 3000, 200
    q = synth.sim3_problem(seed, n, 0.4, scale)
    n_slots = n + 40
    slot_of = np.sort(rng.permutation(n_slots)[:n])
    # synthesise keyframes: identity poses, so camera-frame points == world points for both
    R1 = np.eye(3, dtype=np.float32).reshape(-1); t0 = np.zeros(3, np.float32)
    w1 = rng.normal(size=(n_slots, 3)).astype(np.float32); w2 = w1.copy()
    v1 = np.zeros(n_slots, np.uint8); v2 = np.zeros(n_slots, np.uint8)
    o1 = rng.integers(0, 8, n_slots).astype(np.int32); o2 = o1.copy()
    w1[slot_of], w2[slot_of] = q["x1c"], q["x2c"]
    v1[slot_of] = 1; v2[slot_of] = 1
    s2 = synth.level_sigma2()
    lv1 = np.array([int(np.flatnonzero(s2 == x)[0]) for x in q["sigma2_1"]], np.int32)
    lv2 = np.array([int(np.flatnonzero(s2 == x)[0]) for x in q["sigma2_2"]], np.int32)
    o1[slot_of], o2[slot_of] = lv1, lv2
    i1 = np.arange(n_slots, dtype=np.int32); i2 = np.arange(n_slots, dtype=np.int32)
    K = np.array(q["K"], np.float32)
    blob = struct.pack("<iiIdiii", n_slots, 1 if scale == 1.0 else 0, seed, 0.99, 20, 300, 5)
    blob += R1.tobytes() + t0.tobytes() + R1.tobytes() + t0.tobytes() + K.tobytes()
    blob += w1.tobytes() + w2.tobytes() + o1.tobytes() + o2.tobytes() + i1.tobytes() + i2.tobytes() + v1.tobytes() + v2.tobytes() + s2.tobytes()
    got = _run(driver, "sim3", blob, tmp_path)
    pb = oracle.sim3_problem(q["x1c"], q["x2c"], q["sigma2_1"], q["sigma2_2"], q["K"], q["K"], fix_scale=(scale == 1.0))
    H = oracle.ransac_setup_sim3(n, 0.99, 20, 300)
    o = oracle.sim3_ransac(pb, 0.99, 20, 300, oracle.index_table(seed, n, 3, H))
    assert len(got) == (o["n_hyp"] + 4) // 5                      # iterate(5) calls made
    assert all(g["ok"] == 0 and g["noMore"] == 0 for g in got[:-1])
    g = got[-1]
    assert (g["ok"], g["nInliers"]) == (o["ok"], o["n_inliers"])
    assert np.array_equal(np.array(g["R"], np.float32), o["T"][:3, :3].reshape(-1)) and np.array_equal(np.array(g["t"], np.float32), o["T"][:3, 3])
    assert np.float32(g["s"]) == np.float32(o["scale"])
    assert g["inliers"] == slot_of[np.flatnonzero(o["mask"])].tolist()


def test_optimizer_pose_optimization_batch_and_single(driver, oracle, tmp_path):
    """Optimizer::PoseOptimization(Frame*) mirror: frames with empty keypoint slots, stale mvbOutlier flags,
    monocular and stereo keypoints; one batched call for all candidates, then the per-frame form again on
    candidate 0 from its optimised pose (Tracking.cpp:1284 then :1300)."""
    rng = np.random.default_rng(77)
    C = 6
    blob = struct.pack("<i", C)
    frames = []
    for c in range(C):
        n = [250, 300, 120, 8, 2, 400][c]
        p = synth.poseopt_problem(9800 + c, n, 0.25, [0.0, 1.0, 0.5, 0.0, 0.0, 0.3][c])
        n_slots = n + 60
        slot_of = np.sort(rng.permutation(n_slots)[:n])
        octave_of = np.array([int(np.argmin(np.abs(synth.level_sigma2() - 1.0 / s))) for s in p["inv_sigma2"]], np.int32)
        xy = rng.uniform(0, 700, size=(n_slots, 2)).astype(np.float32)
        ur = np.full(n_slots, -1.0, np.float32)
        octv = rng.integers(0, 8, size=n_slots).astype(np.int32)
        has = np.zeros(n_slots, np.uint8)
        world = rng.normal(size=(n_slots, 3)).astype(np.float32)
        xy[slot_of] = p["obs"][:, :2]
        ur[slot_of] = p["obs"][:, 2]
        octv[slot_of] = octave_of
        has[slot_of] = 1
        world[slot_of] = p["p3d"]
        isig = (np.float32(1.0) / synth.level_sigma2()).astype(np.float32)
        assert np.array_equal(isig[octave_of], p["inv_sigma2"])
        T = np.concatenate([p["Rcw"].ravel(), p["tcw"]]).astype(np.float32)
        blob += struct.pack("<i", n_slots) + p["K"].tobytes() + T.tobytes() + xy.tobytes() + ur.tobytes() + octv.tobytes() \
            + has.tobytes() + world.tobytes() + isig.tobytes()
        frames.append((p, slot_of, n_slots))
    lines = _run(driver, "poseopt", blob, tmp_path)
    assert len(lines) == C + 1
    for c, (p, slot_of, n_slots) in enumerate(frames):
        o, oout = oracle.pose_optimization(oracle.poseopt_problem(p["p3d"], p["obs"], p["inv_sigma2"], p["K"], p["Rcw"], p["tcw"]))
        n = len(p["p3d"])
        if c == 0:      # optimised twice: replay the second call from the first call's float pose
            o2, oout = oracle.pose_optimization(oracle.poseopt_problem(p["p3d"], p["obs"], p["inv_sigma2"], p["K"], o["Rf"], o["tf"]))
            assert abs(lines[C]["single"] - o2["n_inliers"]) <= 1
            o = o2
        else:
            assert abs(lines[c]["nInliers"] - o["n_inliers"]) <= 1, c
        T = np.array(lines[c]["T"], np.float32).reshape(4, 4)
        if n >= 3:
            assert np.abs(T[:3, :3] - o["Rf"]).max() < 2e-6 and np.abs(T[:3, 3] - o["tf"]).max() < 2e-5, c
        else:           # returned before the pose was touched
            assert np.array_equal(T[:3, :3], p["Rcw"]) and np.array_equal(T[:3, 3], p["tcw"]), c
        assert np.array_equal(T[3], np.array([0, 0, 0, 1], np.float32))
        flags = np.zeros(n_slots, bool)
        flags[lines[c]["outliers"]] = True
        empty = np.ones(n_slots, bool)
        empty[slot_of] = False
        assert flags[empty].all()                                   # slots without a MapPoint keep their stale flag
        assert (flags[slot_of] != oout.astype(bool)).sum() <= 1, c  # at most one edge on the chi2 boundary


def test_optimizer_optimize_sim3_batch(driver, oracle, tmp_path):
    """Optimizer::OptimizeSim3 mirror (Sim3Optimizer::OptimizeSim3Batch): keyframe poses + world points in, camera-frame
    points computed in float as Optimizer.cpp:1114,1122 do, invalid / unobserved matches skipped, rejected matches nulled."""
    rng = np.random.default_rng(91)
    C = 4
    blob = struct.pack("<i", C)
    cases = []
    isig = (np.float32(1.0) / synth.level_sigma2()).astype(np.float32)
    f32 = np.float32
    for c in range(C):
        n_ok = [100, 150, 12, 60][c]
        p = synth.sim3opt_problem(8700 + c, n_ok, [0.15, 0.3, 0.5, 0.0][c])
        n = n_ok + 10                                        # 10 extra slots that must be skipped
        nk2 = n + 7
        slot = np.sort(rng.permutation(n)[:n_ok])
        i2 = np.full(n, -1, np.int32)
        i2[slot] = rng.permutation(nk2)[:n_ok].astype(np.int32)
        v1 = np.zeros(n, np.uint8); v2 = np.zeros(n, np.uint8)
        v1[slot] = 1; v2[slot] = 1
        skipped = np.setdiff1d(np.arange(n), slot)
        v2[skipped[:4]] = 1; v1[skipped[:4]] = 0             # pMP1 missing
        v1[skipped[4:7]] = 1; v2[skipped[4:7]] = 1           # both there but not observed in KF2 (i2 < 0)
        R1, t1 = synth.random_pose(rng); R2, t2 = synth.random_pose(rng)
        R1, t1, R2, t2 = R1.astype(f32), t1.astype(f32), R2.astype(f32), t2.astype(f32)
        w1 = rng.normal(size=(n, 3)).astype(f32); w2 = rng.normal(size=(n, 3)).astype(f32)
        w1[slot] = ((p["x1c"].astype(np.float64) - t1) @ R1.astype(np.float64)).astype(f32)
        w2[slot] = ((p["x2c"].astype(np.float64) - t2) @ R2.astype(np.float64)).astype(f32)
        cam = lambda R, t, X: np.stack([((R[r, 0] * X[:, 0] + R[r, 1] * X[:, 1]) + R[r, 2] * X[:, 2]) + t[r] for r in range(3)], axis=1).astype(f32)
        x1c, x2c = cam(R1, t1, w1[slot]), cam(R2, t2, w2[slot])
        oct1 = rng.integers(0, 8, size=n).astype(np.int32); oct2 = rng.integers(0, 8, size=nk2).astype(np.int32)
        lv = lambda s: np.array([int(np.argmin(np.abs(isig - v))) for v in s], np.int32)
        oct1[slot] = lv(p["inv_sigma2_1"]); oct2[i2[slot]] = lv(p["inv_sigma2_2"])
        xy1 = rng.uniform(0, 700, size=(n, 2)).astype(f32); xy2 = rng.uniform(0, 700, size=(nk2, 2)).astype(f32)
        xy1[slot] = p["obs1"]; xy2[i2[slot]] = p["obs2"]
        blob += struct.pack("<ii", n, nk2) + R1.tobytes() + t1.tobytes() + R2.tobytes() + t2.tobytes() + p["K"].tobytes() + p["S12"].tobytes() \
            + xy1.tobytes() + xy2.tobytes() + w1.tobytes() + w2.tobytes() + oct1.tobytes() + oct2.tobytes() + i2.tobytes() + v1.tobytes() \
            + v2.tobytes() + isig.tobytes()
        cases.append((p, slot, n, x1c, x2c))
    lines = _run(driver, "sim3opt", blob, tmp_path)
    assert len(lines) == C
    for c, (p, slot, n, x1c, x2c) in enumerate(cases):
        o, orem = oracle.optimize_sim3(oracle.sim3opt_problem(x1c, x2c, p["obs1"], p["obs2"], p["inv_sigma2_1"], p["inv_sigma2_2"],
                                                              p["K"], p["K"], p["S12"], th2=10.0, fix_scale=True))
        alive = np.zeros(n, bool)
        alive[lines[c]["alive"]] = True
        skipped = np.setdiff1d(np.arange(n), slot)
        assert alive[skipped].all()                           # untouched: the reference `continue`s over them
        if (alive[slot] != ~orem.astype(bool)).sum() > 0:
            continue                                          # a match on the chi2 boundary (see test_gpu_sim3opt.py)
        assert lines[c]["nIn"] == o["n_inliers"], c
        assert np.abs(np.array(lines[c]["R"]).reshape(3, 3) - o["R"]).max() < 1e-6 and np.abs(np.array(lines[c]["t"]) - o["t"]).max() < 1e-6, c
        assert lines[c]["s"] == 1.0
