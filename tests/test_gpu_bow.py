"""GPU parity: batched ORBmatcher::SearchByBoW (CUDA, through the C ABI) vs the CPU oracle -- integer work, bit-exact."""
import numpy as np
import pytest

from ransac_b200 import synth

pytestmark = pytest.mark.gpu


def _check(engine, oracle, sets, qs, ts, mode, orient, ratio=0.75):
    matches, nm = engine.bow_match(sets, qs, ts, ratio, orient, mode)
    keeps = [oracle.bow_features(s) for s in sets]
    total = 0
    for p, (q, t) in enumerate(zip(qs, ts)):
        want, n = oracle.search_by_bow(keeps[q], keeps[t], ratio, orient, mode)
        assert nm[p] == n, (p, nm[p], n)
        assert (matches[p] == want).all(), (p, np.argwhere(matches[p] != want).ravel()[:8])
        total += n
    return total


@pytest.mark.parametrize("orient", [True, False])
def test_relocalisation_shape_one_frame_many_keyframes(engine, oracle, orient):
    """Tracking::Relocalization (Tracking.cpp:1207-1232): every candidate keyframe against the current frame"""
    F = synth.bow_frame(11, 1500, 100)
    kfs = [synth.bow_keyframe(1000 + i, F, 1200 + 37 * (i % 5), shared=0.1 + 0.05 * (i % 7), rot=10.0 * i) for i in range(24)]
    sets = [F] + kfs
    total = _check(engine, oracle, sets, list(range(1, 25)), [0] * 24, 0, orient)
    assert total > 24 * 30


def test_loop_closure_shape_keyframe_vs_keyframes(engine, oracle):
    """LoopClosing::ComputeSim3 (LoopClosing.cpp:251): the current keyframe (outer loop) against every candidate"""
    cur = synth.bow_frame(21, 1300, 90)
    cur["valid"] = (np.random.default_rng(5).random(1300) < 0.75).astype(np.uint8)
    kfs = [synth.bow_keyframe(2000 + i, cur, 1100, shared=0.3, rot=5.0 * i) for i in range(10)]
    # mode 1: query = pKF1 = current keyframe, target = candidate; note the roles: the planted keyframes are the targets
    total = _check(engine, oracle, [cur] + kfs, [0] * 10, list(range(1, 11)), 1, True)
    assert total > 10 * 50


def test_edge_cases_empty_ragged_and_big_nodes(engine, oracle):
    F = synth.bow_frame(31, 400, 3)                      # ~133 features per node: more than one warp pass per node
    kf = synth.bow_keyframe(32, F, 500, shared=0.6)
    empty = dict(desc=np.zeros((0, 8), np.uint32), angle=np.zeros(0, np.float32), valid=None, node_ids=np.zeros(0, np.uint32),
                 node_off=np.zeros(1, np.int32), node_feat=np.zeros(0, np.uint32))
    novalid = dict(kf, valid=np.zeros(500, np.uint8))
    dup = dict(F, desc=np.tile(F["desc"][:1], (400, 1)), valid=np.ones(400, np.uint8))    # every distance 0: ties everywhere
    sets = [F, kf, empty, novalid, dup]
    _check(engine, oracle, sets, [1, 2, 1, 3, 4, 4], [0, 0, 2, 0, 0, 4], 0, True)
    _check(engine, oracle, sets, [1, 4], [4, 1], 1, True)
    _check(engine, oracle, sets, [1], [0], 0, True, ratio=0.95)
