"""Oracle pins for ORBmatcher::SearchByBoW (oracle/orc_bow.c).

The reference ships no test or fixture for this function, so the restatement is pinned against an independent,
literal Python transcription of the two overloads (src/ORBmatcher.cpp:110-239, :354-487) on small cases, against
numpy popcounts for DescriptorDistance (:1492-1508) and against brute-force properties of ComputeThreeMaxima (:1445-1488)."""
import numpy as np

from ransac_b200 import synth


def _py_distance(a, b):
    return int(sum(bin(int(x) ^ int(y)).count("1") for x, y in zip(a, b)))


def _py_three_maxima(h):
    max1 = max2 = max3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(h):
        if s > max1:
            max3, max2, max1 = max2, max1, s
            i3, i2, i1 = i2, i1, i
        elif s > max2:
            max3, max2 = max2, s
            i3, i2 = i2, i
        elif s > max3:
            max3 = s
            i3 = i
    if max2 < np.float32(0.1) * np.float32(max1):
        i2 = i3 = -1
    elif max3 < np.float32(0.1) * np.float32(max1):
        i3 = -1
    return i1, i2, i3


def _py_search_by_bow(q, t, nn_ratio, check_orientation, mode):
    """ORBmatcher.cpp:110-239 (mode 0) / :354-487 (mode 1), transcribed with dicts for the two std::maps"""
    fq = {int(n): [int(x) for x in q["node_feat"][q["node_off"][k]:q["node_off"][k + 1]]] for k, n in enumerate(q["node_ids"])}
    ft = {int(n): [int(x) for x in t["node_feat"][t["node_off"][k]:t["node_off"][k + 1]]] for k, n in enumerate(t["node_ids"])}
    n_out = t["desc"].shape[0] if mode == 0 else q["desc"].shape[0]
    out = [-1] * n_out
    taken = [False] * t["desc"].shape[0]
    rot_hist = [[] for _ in range(30)]
    nmatches = 0
    for node in sorted(set(fq) & set(ft)):
        for iq in fq[node]:
            if q["valid"] is not None and not q["valid"][iq]:
                continue
            best1, best2, best_idx = 256, 256, -1
            for it in ft[node]:
                if taken[it]:
                    continue
                if mode == 1 and t["valid"] is not None and not t["valid"][it]:
                    continue
                d = _py_distance(q["desc"][iq], t["desc"][it])
                if d < best1:
                    best2, best1, best_idx = best1, d, it
                elif d < best2:
                    best2 = d
            ok = best1 <= 50 if mode == 0 else best1 < 50
            if ok and np.float32(best1) < np.float32(nn_ratio) * np.float32(best2):
                taken[best_idx] = True
                oi = best_idx if mode == 0 else iq
                out[oi] = iq if mode == 0 else best_idx
                if check_orientation:
                    rot = np.float32(q["angle"][iq]) - np.float32(t["angle"][best_idx])
                    if rot < 0.0:
                        rot = np.float32(rot + np.float32(360.0))
                    x = float(np.float32(rot * np.float32(1.0 / 30)))
                    b = int(np.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)      # round(): halves away from zero
                    if b == 30:
                        b = 0
                    rot_hist[b].append(oi)
                nmatches += 1
    if check_orientation:
        i1, i2, i3 = _py_three_maxima([len(h) for h in rot_hist])
        for i in range(30):
            if i in (i1, i2, i3):
                continue
            for oi in rot_hist[i]:
                out[oi] = -1
                nmatches -= 1
    return np.array(out, np.int32), nmatches


def test_descriptor_distance_is_popcount(oracle):
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = rng.integers(0, 2 ** 32, 8, dtype=np.uint64).astype(np.uint32)
        b = rng.integers(0, 2 ** 32, 8, dtype=np.uint64).astype(np.uint32)
        assert oracle.descriptor_distance(a, b) == _py_distance(a, b) == int(np.unpackbits((a ^ b).view(np.uint8)).sum())
    z = np.zeros(8, np.uint32)
    assert oracle.descriptor_distance(z, z) == 0 and oracle.descriptor_distance(z, ~z) == 256


def test_three_maxima(oracle):
    rng = np.random.default_rng(1)
    cases = [np.zeros(30, int), np.full(30, 7), np.arange(30), np.arange(30)[::-1].copy()]
    cases += [rng.integers(0, 40, 30) for _ in range(200)] + [rng.integers(0, 3, 30) for _ in range(100)]
    one = np.zeros(30, int); one[17] = 100; one[3] = 9; one[20] = 10
    cases.append(one)
    for h in cases:
        assert oracle.three_maxima(h) == _py_three_maxima(list(h))


def test_search_by_bow_matches_the_literal_transcription(oracle):
    for seed, mode, orient in ((1, 0, True), (2, 0, False), (3, 1, True), (4, 1, False), (5, 0, True)):
        F = synth.bow_frame(100 + seed, 260, 12)
        KF = synth.bow_keyframe(200 + seed, F, 240, shared=0.5, flip_bits=45 if seed == 5 else 25)
        if mode == 1:
            F = dict(F, valid=(np.random.default_rng(seed).random(260) < 0.8).astype(np.uint8))
        got, n = oracle.search_by_bow(oracle.bow_features(KF), oracle.bow_features(F), 0.75, orient, mode)
        want, nw = _py_search_by_bow(KF, F, 0.75, orient, mode)
        assert n == nw and (got == want).all(), (seed, mode, orient)
        assert n == int((got >= 0).sum()) and n > 20
        if mode == 0:
            # sanity: the matches are mostly the planted ones
            dst, src = KF["truth"]
            planted = dict(zip(src.tolist(), dst.tolist()))
            hit = sum(1 for i, qv in enumerate(got) if qv >= 0 and planted.get(i) == qv)
            assert hit >= 0.9 * n


def test_search_by_bow_edge_cases(oracle):
    F = synth.bow_frame(7, 64, 4)
    KF = synth.bow_keyframe(8, F, 64)
    # no usable MapPoint at all
    KF0 = dict(KF, valid=np.zeros(64, np.uint8))
    got, n = oracle.search_by_bow(oracle.bow_features(KF0), oracle.bow_features(F), 0.75, True, 0)
    assert n == 0 and (got == -1).all()
    # disjoint vocabularies
    F2 = dict(F, node_ids=(F["node_ids"] + np.uint32(5_000_000)))
    got, n = oracle.search_by_bow(oracle.bow_features(KF), oracle.bow_features(F2), 0.75, True, 0)
    assert n == 0
    # identical descriptors everywhere: best == second best, the ratio test rejects everything
    same = dict(F, desc=np.tile(F["desc"][:1], (64, 1)))
    got, n = oracle.search_by_bow(oracle.bow_features(dict(same, valid=np.ones(64, np.uint8))), oracle.bow_features(same), 0.75, False, 0)
    assert n == 0
