"""CPU: the oracle's candidate retrieval (oracle/orc_kfdb.c) against a literal Python transcription of
KeyFrameDatabase::DetectRelocalizationCandidates / DetectLoopCandidates (src/KeyFrameDatabase.cpp:174-284, 51-172) built on
dicts and lists exactly as the reference builds them on std::map / std::list, with float32 arithmetic where the reference
uses float.  Candidate lists must be equal element for element (order included)."""
import numpy as np

from ransac_b200 import synth

F32 = np.float32


def _l1_score(qw, qv, kw, kv):
    """L1Scoring::score (Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-66), sequential double sum in word order"""
    i = j = 0
    score = 0.0
    while i < len(qw) and j < len(kw):
        if qw[i] == kw[j]:
            score += abs(qv[i] - kv[j]) - abs(qv[i]) - abs(kv[j])
            i += 1; j += 1
        elif qw[i] < kw[j]:
            i = int(np.searchsorted(qw, kw[j], side="left"))
        else:
            j = int(np.searchsorted(kw, qw[i], side="left"))
    return -score / 2.0


class _KF:
    def __init__(self, idx, words, vals, covis):
        self.idx, self.words, self.vals, self.covis = idx, words, vals, covis
        self.query = None
        self.nwords = 0
        self.score = F32(0)


def _transcription(db, mode, qw, qv, query_id, kfs, conn=(), min_score=0.0):
    inverted = {}
    for kf in kfs:                                                   # KeyFrameDatabase::add, insertion order
        for w in kf.words:
            inverted.setdefault(int(w), []).append(kf)
    connected = set(int(c) for c in conn)
    share = []
    for w in qw:                                                     # std::map order: ascending
        for kf in inverted.get(int(w), []):
            if kf.query != query_id:
                kf.nwords = 0
                if mode == 0 or kf.idx not in connected:
                    kf.query = query_id
                    share.append(kf)
            kf.nwords += 1
    if not share:
        return []
    max_common = max(kf.nwords for kf in share)
    min_common = int(F32(max_common) * F32(0.8))
    score_and_match = []
    for kf in share:
        if kf.nwords > min_common:
            si = F32(_l1_score(qw, qv, kf.words, kf.vals))
            kf.score = si
            if mode == 0 or si >= F32(min_score):
                score_and_match.append((si, kf))
    if not score_and_match:
        return []
    acc_list = []
    best_acc = F32(0) if mode == 0 else F32(min_score)
    for si, kf in score_and_match:
        best_score, acc, best = si, si, kf
        for j in kf.covis:
            if j < 0:
                break
            k2 = kfs[j]
            if mode == 0:
                if k2.query != query_id:
                    continue
            elif not (k2.query == query_id and k2.nwords > min_common):
                continue
            acc = F32(acc + k2.score)
            if k2.score > best_score:
                best, best_score = k2, k2.score
        acc_list.append((acc, best))
        if acc > best_acc:
            best_acc = acc
    retain = F32(F32(0.75) * best_acc)
    out, added = [], set()
    for acc, kf in acc_list:
        if acc > retain and kf.idx not in added:
            out.append(kf.idx)
            added.add(kf.idx)
    return out


def _kfs(db):
    return [_KF(k, db["bow_word"][db["bow_off"][k]:db["bow_off"][k + 1]], db["bow_val"][db["bow_off"][k]:db["bow_off"][k + 1]], db["covis"][k])
            for k in range(db["K"])]


def test_l1_score_matches_transcription_and_numpy(oracle):
    rng = np.random.default_rng(3)
    for _ in range(20):
        w1, v1 = synth._bow_vector(rng, rng.integers(0, 5000, 800))
        w2, v2 = synth._bow_vector(rng, rng.integers(0, 5000, 700))
        s = oracle.bow_l1_score(w1, v1, w2, v2)
        assert s == _l1_score(w1, v1, w2, v2)
        d1 = dict(zip(w1.tolist(), v1.tolist())); d2 = dict(zip(w2.tolist(), v2.tolist()))
        l1 = sum(abs(d1.get(k, 0.0) - d2.get(k, 0.0)) for k in set(d1) | set(d2))
        assert abs(s - (1.0 - 0.5 * l1)) < 1e-12                      # Nister's identity for L1-normalised vectors


def test_relocalization_candidates_equal_transcription_with_state(oracle):
    """a sequence of queries against one database: mRelocScore is carried from query to query (quirk Q11)"""
    db = synth.kf_database(1, K=300, n_places=30)
    odb = oracle.kfdb(db)
    kfs = _kfs(db)
    state = np.zeros(db["K"], np.float32)
    n_total = 0
    for q in range(12):
        qw, qv = synth.kf_query(100 + q, db, place=(q * 7) % 30)
        want = _transcription(db, 0, qw, qv, ("reloc", q), kfs)
        got = oracle.detect_candidates(odb, qw, qv, mode=0, score_state=state)
        assert got.tolist() == want, q
        assert (state == np.array([kf.score for kf in kfs], np.float32)).all()
        n_total += len(want)
    assert n_total >= 12                                             # every query finds its place


def test_loop_candidates_equal_transcription(oracle):
    db = synth.kf_database(2, K=300, n_places=30)
    odb = oracle.kfdb(db)
    found = 0
    for q in range(270, 300, 3):                                     # the revisiting stretch of the trajectory
        kfs = _kfs(db)
        qw = db["bow_word"][db["bow_off"][q]:db["bow_off"][q + 1]]
        qv = db["bow_val"][db["bow_off"][q]:db["bow_off"][q + 1]]
        conn = [int(c) for c in db["covis"][q] if c >= 0] + [q]       # GetConnectedKeyFrames (+ itself: it is in the database)
        # minScore as LoopClosing::DetectLoop computes it: the lowest score to a connected keyframe (LoopClosing.cpp:121-133)
        min_score = min([1.0] + [float(oracle.bow_l1_score(qw, qv, kfs[c].words, kfs[c].vals)) for c in conn if c != q])
        for ms in (min_score, 0.0, 0.9):
            kfs = _kfs(db)
            want = _transcription(db, 1, qw, qv, ("loop", q), kfs, conn=conn, min_score=ms)
            got = oracle.detect_candidates(odb, qw, qv, mode=1, conn=conn, min_score=ms)
            assert got.tolist() == want, (q, ms)
            found += len(want)
            assert not set(want) & set(conn)
    assert found > 0


def test_retrieval_edge_cases(oracle):
    db = synth.kf_database(3, K=50, n_places=5)
    odb = oracle.kfdb(db)
    # a query that shares no word with anybody
    assert oracle.detect_candidates(odb, np.array([db["vocab"] + 5], np.uint32), np.array([1.0])).tolist() == []
    assert oracle.detect_candidates(odb, np.zeros(0, np.uint32), np.zeros(0)).tolist() == []
    # every keyframe connected: nothing left for the loop detector
    qw, qv = synth.kf_query(9, db, 2)
    assert oracle.detect_candidates(odb, qw, qv, mode=1, conn=list(range(50)), min_score=0.0).tolist() == []
