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


def test_pnp_batch_built_on_device_from_search_by_bow(engine):
    """Tracking::Relocalization (Tracking.cpp:1207-1232): SearchByBoW per candidate, a PnPsolver for every candidate with at least
    15 matches.  rsac_pnp_upload_from_bow builds that PnP batch on the device from the match arrays of the last rsac_bow_run
    -- only the match counts visit the host.  It must equal the batch a host would build from the downloaded matches
    (rsac_pnp_upload_indexed): same records and masks; and the candidates that really saw the place recover the true pose."""
    from ransac_b200 import capi
    w = synth.reloc_world(7, C=24)
    C = len(w["kfs"])
    sets = [w["frame"]] + w["kfs"]
    engine.bow_upload(sets, list(range(1, C + 1)), [0] * C, 0.75, True, 0)
    engine.bow_run()
    matches, nm = engine.bow_download()
    prm = capi.ransac_params(0.99, 10, 300, 4, 0.5, 5.991)              # Tracking.cpp:1226
    # (a) on the device
    offsets = engine.pnp_upload_from_bow(nm, w["K"], prm, w["seeds"], min_matches=15, kp_uv=w["kp_uv"], kp_sigma2=w["kp_sigma2"], mp_xyz=w["mp_xyz"])
    engine.pnp_run(capi.FLAG_EARLY_EXIT)
    res_a, m_a = engine.pnp_download()
    engine.poseopt_from_pnp()
    engine.poseopt_run()
    po_a, fl_a = engine.poseopt_download()
    # (b) the same batch from the downloaded matches
    kp_idx, mp_idx, off_b = [], [], [0]
    for c in range(C):
        js = np.flatnonzero(matches[c] >= 0) if nm[c] >= 15 else np.zeros(0, np.int64)
        kp_idx.append(js.astype(np.uint16)); mp_idx.append(w["kfs"][c]["mp_index"][matches[c][js]].astype(np.uint32))
        off_b.append(off_b[-1] + len(js))
    assert offsets.tolist() == off_b
    engine.pnp_upload_indexed(np.array(off_b, np.int32), np.concatenate(kp_idx), np.concatenate(mp_idx), w["K"], prm, seeds=w["seeds"])
    engine.pnp_run(capi.FLAG_EARLY_EXIT)
    res_b, m_b = engine.pnp_download()
    assert res_a.tobytes() == res_b.tobytes() and (m_a == m_b).all()
    engine.poseopt_from_pnp()
    engine.poseopt_run()
    po_b, fl_b = engine.poseopt_download()
    assert po_a.tobytes() == po_b.tobytes() and (fl_a == fl_b).all()
    discarded = np.array([nm[c] < 15 for c in range(C)])
    assert discarded.sum() >= 3 and not res_a["ok"][discarded].any() and (res_a["no_more"][discarded] == 1).all()
    ok = res_a["ok"] == 1
    assert ok.sum() >= C // 2
    for c in np.flatnonzero(ok):
        assert np.abs(res_a[c]["R"].reshape(3, 3) - w["R"]).max() < 0.03, c
        assert np.abs(po_a[c]["Rf"].reshape(3, 3) - w["R"]).max() < 0.02, c
