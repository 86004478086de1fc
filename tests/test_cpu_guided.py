"""CPU: the oracle's SearchBySim3 (oracle/orc_guided.c) against a literal numpy-float32 transcription of
ORBmatcher::SearchBySim3 (src/ORBmatcher.cpp:948-1171) with its helpers (KeyFrame::GetFeaturesInArea, IsInImage,
MapPoint::PredictScale, DescriptorDistance).  Match arrays must be equal element for element."""
import math

import numpy as np

from ransac_b200 import synth

F = np.float32


def _features_in_area(kf, x, y, r):
    """KeyFrame::GetFeaturesInArea (src/KeyFrame.cpp:560-599) on the mGrid[ix][iy] vectors"""
    out = []
    mnMinX, mnMinY = F(kf["bounds"][0]), F(kf["bounds"][2])
    w_inv, h_inv = F(kf["grid_w_inv"]), F(kf["grid_h_inv"])
    cols, rows = kf["grid_cols"], kf["grid_rows"]
    nMinCellX = max(0, int(math.floor(F(F(x - mnMinX) - r) * w_inv)))
    if nMinCellX >= cols:
        return out
    nMaxCellX = min(cols - 1, int(math.ceil(F(F(x - mnMinX) + r) * w_inv)))
    if nMaxCellX < 0:
        return out
    nMinCellY = max(0, int(math.floor(F(F(y - mnMinY) - r) * h_inv)))
    if nMinCellY >= rows:
        return out
    nMaxCellY = min(rows - 1, int(math.ceil(F(F(y - mnMinY) + r) * h_inv)))
    if nMaxCellY < 0:
        return out
    for ix in range(nMinCellX, nMaxCellX + 1):
        for iy in range(nMinCellY, nMaxCellY + 1):
            c = ix * rows + iy
            for j in range(kf["grid_off"][c], kf["grid_off"][c + 1]):
                idx = int(kf["grid_idx"][j])
                if abs(F(kf["kp_xy"][idx, 0] - x)) < r and abs(F(kf["kp_xy"][idx, 1] - y)) < r:
                    out.append(idx)
    return out


def _dist(a, b):
    return int(sum(bin(int(x) ^ int(y)).count("1") for x, y in zip(a, b)))


def _mv(R, p, t):
    return np.array([F(F(F(R[i, 0] * p[0]) + F(R[i, 1] * p[1])) + F(R[i, 2] * p[2])) + t[i] for i in range(3)], F)


def _one_way(src, dst, K, Rds, tds, th, already):
    match = [-1] * src["n_feat"]
    fx, fy, cx, cy = (F(k) for k in K)
    Rs, ts = src["Rcw"].astype(F), src["tcw"].astype(F)
    for i in range(src["n_feat"]):
        if not src["mp_valid"][i] or already[i]:
            continue
        pcs = _mv(Rs, src["mp_xyz"][i].astype(F), ts)
        pc = _mv(Rds, pcs, tds)
        if pc[2] < 0.0:
            continue
        invz = F(1.0 / float(pc[2]))
        x, y = F(pc[0] * invz), F(pc[1] * invz)
        u, v = F(F(fx * x) + cx), F(F(fy * y) + cy)
        b = dst["bounds"]
        if not (u >= b[0] and u < b[1] and v >= b[2] and v < b[3]):
            continue
        maxD, minD = F(F(1.2) * src["mp_maxdist"][i]), F(F(0.8) * src["mp_mindist"][i])
        d3 = F(np.sqrt(F(F(F(pc[0] * pc[0]) + F(pc[1] * pc[1])) + F(pc[2] * pc[2]))))
        if d3 < minD or d3 > maxD:
            continue
        ratio = F(src["mp_maxdist"][i] / d3)
        lvl = int(math.ceil(math.log(float(ratio)) / float(F(dst["log_scale_factor"]))))
        lvl = 0 if lvl < 0 else min(lvl, dst["n_levels"] - 1)
        radius = F(F(th) * dst["scale_factors"][lvl])
        best, best_idx = 2 ** 31 - 1, -1
        for idx in _features_in_area(dst, u, v, radius):
            if dst["kp_octave"][idx] < lvl - 1 or dst["kp_octave"][idx] > lvl:
                continue
            d = _dist(src["mp_desc"][i], dst["desc"][idx])
            if d < best:
                best, best_idx = d, idx
        if best <= 100:
            match[i] = best_idx
    return match


def _search_by_sim3(p, th, matched_in):
    k1, k2 = p["kf1"], p["kf2"]
    R12, t12 = p["R12"].astype(F).reshape(3, 3), p["t12"].astype(F)
    R21 = R12.T.copy()
    t21 = np.array([-(F(F(F(R21[i, 0] * t12[0]) + F(R21[i, 1] * t12[1])) + F(R21[i, 2] * t12[2]))) for i in range(3)], F)
    am1, am2 = [False] * k1["n_feat"], [False] * k2["n_feat"]
    for i in range(k1["n_feat"]):
        j = -1 if matched_in is None else int(matched_in[i])
        if j != -1:
            am1[i] = True
            if 0 <= j < k2["n_feat"]:
                am2[j] = True
    m1 = _one_way(k1, k2, p["K"], R21, t21, th, am1)
    m2 = _one_way(k2, k1, p["K"], R12, t12, th, am2)
    return [m1[i] if (m1[i] >= 0 and m2[m1[i]] == i) else -1 for i in range(k1["n_feat"])]


def test_features_in_area_and_predict_scale(oracle):
    p = synth.kf_view_pair(3, n_points=500, n_extra=200)
    kf = p["kf1"]
    okf = oracle.kf_view(kf)
    rng = np.random.default_rng(0)
    for _ in range(200):
        x, y, r = F(rng.uniform(-30, 790)), F(rng.uniform(-30, 510)), F(rng.uniform(1, 40))
        assert oracle.features_in_area(okf, x, y, r).tolist() == _features_in_area(kf, x, y, r)
    for _ in range(200):
        md, cd = F(rng.uniform(1, 40)), F(rng.uniform(0.5, 40))
        lvl = int(math.ceil(math.log(float(F(md / cd))) / float(F(kf["log_scale_factor"]))))
        assert oracle.predict_scale(md, cd, kf["log_scale_factor"], 8) == (0 if lvl < 0 else min(lvl, 7))


def test_search_by_sim3_equals_transcription(oracle):
    for seed, n_pts, pre in ((1, 500, 0.3), (2, 400, 0.0), (4, 300, 0.6)):
        p = synth.kf_view_pair(seed, n_points=n_pts, n_extra=150, prematched=pre)
        k1, k2 = oracle.kf_view(p["kf1"]), oracle.kf_view(p["kf2"])
        for mi in (p["matched12_in"], None):
            got, n = oracle.search_by_sim3(k1, k2, p["K"], p["R12"], p["t12"], 7.5, mi)
            want = _search_by_sim3(p, 7.5, mi)
            assert got.tolist() == want
            assert n == sum(1 for w in want if w >= 0)
        # the matches are the right ones: both features observe the same map point
        good = [p["kf1"]["mp_id"][i] == p["kf2"]["mp_id"][j] for i, j in enumerate(got) if j >= 0]
        assert len(good) > 50 and np.mean(good) > 0.97


def test_search_by_sim3_wrong_transform_finds_little(oracle):
    p = synth.kf_view_pair(5, n_points=400, n_extra=150, prematched=0.0)
    k1, k2 = oracle.kf_view(p["kf1"]), oracle.kf_view(p["kf2"])
    t_bad = p["t12"] + np.float32([1.5, -1.0, 0.8])
    got, n = oracle.search_by_sim3(k1, k2, p["K"], p["R12"], t_bad, 7.5, None)
    assert got.tolist() == _search_by_sim3(dict(p, t12=t_bad), 7.5, None)
    assert n < 40


# ------------------------------------------------------------------ SearchByProjection(Frame, KeyFrame, sAlreadyFound, th, ORBdist)
def _frame_features_in_area(fr, x, y, r, min_level, max_level):
    """Frame::GetFeaturesInArea (src/Frame.cpp:393-446)"""
    out = []
    mnMinX, mnMinY = F(fr["bounds"][0]), F(fr["bounds"][2])
    w_inv, h_inv = F(fr["grid_w_inv"]), F(fr["grid_h_inv"])
    cols, rows = fr["grid_cols"], fr["grid_rows"]
    nMinCellX = max(0, int(math.floor(F(F(x - mnMinX) - r) * w_inv)))
    if nMinCellX >= cols:
        return out
    nMaxCellX = min(cols - 1, int(math.ceil(F(F(x - mnMinX) + r) * w_inv)))
    if nMaxCellX < 0:
        return out
    nMinCellY = max(0, int(math.floor(F(F(y - mnMinY) - r) * h_inv)))
    if nMinCellY >= rows:
        return out
    nMaxCellY = min(rows - 1, int(math.ceil(F(F(y - mnMinY) + r) * h_inv)))
    if nMaxCellY < 0:
        return out
    check = (min_level > 0) or (max_level >= 0)
    for ix in range(nMinCellX, nMaxCellX + 1):
        for iy in range(nMinCellY, nMaxCellY + 1):
            c = ix * rows + iy
            for j in range(fr["grid_off"][c], fr["grid_off"][c + 1]):
                idx = int(fr["grid_idx"][j])
                if check:
                    if fr["kp_octave"][idx] < min_level:
                        continue
                    if max_level >= 0 and fr["kp_octave"][idx] > max_level:
                        continue
                if abs(F(fr["kp_xy"][idx, 0] - x)) < r and abs(F(fr["kp_xy"][idx, 1] - y)) < r:
                    out.append(idx)
    return out


def _three_maxima(histo):
    max1 = max2 = max3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(histo):
        if s > max1:
            max3, max2, max1 = max2, max1, s
            i3, i2, i1 = i2, i1, i
        elif s > max2:
            max3, max2 = max2, s
            i3, i2 = i2, i
        elif s > max3:
            max3, i3 = s, i
    if max2 < F(0.1) * F(max1):
        i2 = i3 = -1
    elif max3 < F(0.1) * F(max1):
        i3 = -1
    return i1, i2, i3


def _search_by_projection(c, th, orb_dist, check_orientation):
    fr, kf = c["frame"], c["kf"]
    Rcw, tcw = c["Rcw"].astype(F).reshape(3, 3), c["tcw"].astype(F)
    fx, fy, cx, cy = (F(k) for k in c["K"])
    Ow = np.array([F(F(F(-Rcw[0, i] * tcw[0]) + F(-Rcw[1, i] * tcw[1])) + F(-Rcw[2, i] * tcw[2])) for i in range(3)], F)
    frame_mp = [bool(o) for o in c["occupied"]]                   # CurrentFrame.mvpMapPoints[i2] != nullptr
    match = [-1] * fr["n_feat"]
    rot_hist = [[] for _ in range(30)]
    nmatches = 0
    for i in range(kf["n_feat"]):
        if not kf["mp_valid"][i] or c["already_found"][i]:
            continue
        xw = kf["mp_xyz"][i].astype(F)
        xc = _mv(Rcw, xw, tcw)
        invzc = F(1.0 / float(xc[2]))
        u = F(F(F(fx * xc[0]) * invzc) + cx)
        v = F(F(F(fy * xc[1]) * invzc) + cy)
        b = fr["bounds"]
        if u < b[0] or u > b[1] or v < b[2] or v > b[3]:
            continue
        PO = (xw - Ow).astype(F)
        d3 = F(np.sqrt(F(F(F(PO[0] * PO[0]) + F(PO[1] * PO[1])) + F(PO[2] * PO[2]))))
        if d3 < F(F(0.8) * kf["mp_mindist"][i]) or d3 > F(F(1.2) * kf["mp_maxdist"][i]):
            continue
        lvl = int(math.ceil(math.log(float(F(kf["mp_maxdist"][i] / d3))) / float(F(fr["log_scale_factor"]))))
        lvl = 0 if lvl < 0 else min(lvl, fr["n_levels"] - 1)
        radius = F(F(th) * fr["scale_factors"][lvl])
        best, best_idx = 256, -1
        for i2 in _frame_features_in_area(fr, u, v, radius, lvl - 1, lvl + 1):
            if frame_mp[i2]:
                continue
            d = _dist(kf["mp_desc"][i], fr["desc"][i2])
            if d < best:
                best, best_idx = d, i2
        if best <= orb_dist:
            frame_mp[best_idx] = True
            match[best_idx] = i
            nmatches += 1
            if check_orientation:
                rot = F(kf["kp_angle"][i] - fr["kp_angle"][best_idx])
                if rot < 0.0:
                    rot = F(rot + F(360.0))
                x = float(F(rot * F(F(1.0) / F(30))))
                bn = int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))     # roundf: half away from zero
                if bn == 30:
                    bn = 0
                rot_hist[bn].append(best_idx)
    if check_orientation:
        i1, i2_, i3 = _three_maxima([len(h) for h in rot_hist])
        for bn in range(30):
            if bn not in (i1, i2_, i3):
                for idx in rot_hist[bn]:
                    match[idx] = -1
                    nmatches -= 1
    return match, nmatches


def test_search_by_projection_equals_transcription(oracle):
    for seed, th, od, co in ((1, 10.0, 100, True), (2, 3.0, 64, True), (3, 10.0, 100, False), (4, 15.0, 100, True)):
        c = synth.proj_search_case(seed, n_points=450, n_extra=150)
        f, k = oracle.kf_view(c["frame"]), oracle.kf_view(c["kf"])
        got, n = oracle.search_by_projection(f, k, c["K"], c["Rcw"], c["tcw"], th, od, co, c["occupied"], c["already_found"])
        want, wn = _search_by_projection(c, th, od, co)
        assert got.tolist() == want and n == wn, seed
        assert n > 80
        assert not (got[c["occupied"] > 0] >= 0).any()            # occupied keypoints are never reassigned
