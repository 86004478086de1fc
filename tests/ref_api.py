"""ctypes binding of oracle/_ref/libref_solvers.so -- TEST INFRASTRUCTURE ONLY.

The library holds the REFERENCE's own PnPsolver.cpp, Sim3Solver.cpp, MLPnPsolver.cpp, KeyFrameDatabase.cpp, ORBmatcher.cpp,
DUtils/Random.cpp and the vendored DBoW2 BowVector / FeatureVector / ScoringObject, compiled unmodified from
/root/reference by `make -C oracle ref` against the stand-in headers under oracle/shim/ (Eigen and OpenCV are not
installed in this image; what the stand-ins pin and what they cannot is stated in oracle/shim/Eigen/Dense).
It exists to check the oracle; nothing in the product, bench.py or the GPU tests loads it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
PATH = os.environ.get("REF_SO", os.path.join(ORACLE_DIR, "_ref", "libref_solvers.so"))
_LIB = None


def available() -> bool:
    """built here (needs /root/reference) or shipped prebuilt"""
    if not os.path.exists(PATH) and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "orc_linalg.o", "orc_guided.o", "orc_bow.o"], check=False)
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=False)
    return os.path.exists(PATH)


def _load(path):
    L = C.CDLL(path)
    for f in ("ref_pnp_create", "ref_sim3_create", "ref_mlpnp_create", "ref_kfdb_create"):
        getattr(L, f).restype = C.c_void_p
    L.ref_pnp_compute_pose.restype = C.c_double
    L.ref_bow_l1_score.restype = C.c_double
    return L


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _load(PATH)
    return _LIB


LAPACK_PATH = os.path.join(ORACLE_DIR, "_ref", "libref_solvers_lapack.so")


def lapack_available() -> bool:
    """the variant whose dense solves go to LAPACK (`make -C oracle ref-lapack`)"""
    if not os.path.exists(LAPACK_PATH) and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "orc_linalg.o", "orc_guided.o", "orc_bow.o"], check=False)
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref-lapack"], check=False)
    return os.path.exists(LAPACK_PATH)


class use:
    """with ref_api.use(ref_api.LAPACK_PATH): ... -- objects created inside bind to that library"""

    def __init__(self, path):
        self.path = path

    def __enter__(self):
        global _LIB
        self.prev = _LIB
        _LIB = _load(self.path)
        return _LIB

    def __exit__(self, *exc):
        global _LIB
        _LIB = self.prev
        return False


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def random_ints(seed, lo, hi, count):
    L = lib()
    L.ref_seed(C.c_int(seed))
    return [L.ref_random_int(C.c_int(lo), C.c_int(hi)) for _ in range(count)]


def seed(s):
    lib().ref_seed(C.c_int(s))


class PnP:
    """The reference's PnPsolver over one frame: keypoint i carries map point i unless state[i] == 0 (none) / 2 (bad)."""

    def __init__(self, kp_xy, octave, level_sigma2, mp_xyz, K, state=None):
        self.L = lib_kfdb()
        kp_xy = np.ascontiguousarray(kp_xy, np.float32)
        mp_xyz = np.ascontiguousarray(mp_xyz, np.float32)
        octave = np.ascontiguousarray(octave, np.int32)
        ls2 = np.ascontiguousarray(level_sigma2, np.float32)
        self.n_kp = kp_xy.shape[0]
        st = np.ones(self.n_kp, np.uint8) if state is None else np.ascontiguousarray(state, np.uint8)
        self.h = C.c_void_p(self.L.ref_pnp_create(C.c_int(self.n_kp), _p(kp_xy), _p(octave), _p(ls2), C.c_int(len(ls2)), _p(mp_xyz), _p(st),
                                                 C.c_float(K[0]), C.c_float(K[1]), C.c_float(K[2]), C.c_float(K[3])))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_pnp_destroy(self.h)
            self.h = None

    def set_params(self, prob=0.99, min_inliers=8, max_its=300, min_set=4, eps=0.4, th2=5.991):
        self.L.ref_pnp_set_params(self.h, C.c_double(prob), C.c_int(min_inliers), C.c_int(max_its), C.c_int(min_set), C.c_float(eps), C.c_float(th2))

    def params(self):
        N, mi, its = C.c_int(), C.c_int(), C.c_int()
        eps = C.c_float()
        self.L.ref_pnp_get_params(self.h, C.byref(N), C.byref(mi), C.byref(its), C.byref(eps), None, None)
        max_err = np.zeros(max(N.value, 1), np.float32)
        kpi = np.zeros(max(N.value, 1), np.int32)
        self.L.ref_pnp_get_params(self.h, C.byref(N), C.byref(mi), C.byref(its), C.byref(eps), _p(max_err), _p(kpi))
        return dict(N=N.value, min_inliers=mi.value, max_its=its.value, eps=eps.value, max_err=max_err[:N.value], kp_index=kpi[:N.value])

    def iterate(self, n_iterations):
        no_more, n_inl = C.c_int(), C.c_int()
        inl = np.zeros(max(self.n_kp, 1), np.uint8)
        T = np.zeros(16, np.float32)
        ok = self.L.ref_pnp_iterate(self.h, C.c_int(n_iterations), C.byref(no_more), _p(inl), C.byref(n_inl), _p(T))
        return dict(ok=bool(ok), no_more=bool(no_more.value), n_inliers=n_inl.value, inliers=inl[:self.n_kp].astype(bool), T=T.reshape(4, 4))

    def state(self, N):
        it, best, refd = C.c_int(), C.c_int(), C.c_int()
        T = np.zeros(16, np.float32)
        mask = np.zeros(max(N, 1), np.uint8)
        self.L.ref_pnp_state(self.h, C.byref(it), C.byref(best), C.byref(refd), _p(T), _p(mask))
        return dict(iterations=it.value, best_inliers=best.value, refined_inliers=refd.value, best_T=T.reshape(4, 4), best_mask=mask[:N].astype(bool))

    def compute_pose(self, idx):
        idx = np.ascontiguousarray(idx, np.int32)
        R, t = np.zeros(9, np.float32), np.zeros(3, np.float32)
        err = self.L.ref_pnp_compute_pose(self.h, _p(idx), C.c_int(len(idx)), _p(R), _p(t))
        return R.reshape(3, 3), t, err

    def check_inliers(self, R, t, N):
        R = np.ascontiguousarray(R, np.float32).reshape(-1)
        t = np.ascontiguousarray(t, np.float32)
        mask = np.zeros(max(N, 1), np.uint8)
        cnt = self.L.ref_pnp_check_inliers(self.h, _p(R), _p(t), _p(mask))
        return cnt, mask[:N].astype(bool)


class Sim3:
    """The reference's Sim3Solver over two keyframes at the origin (camera-frame points given directly)."""

    def __init__(self, x1c, x2c, octave1, octave2, level_sigma2, K1, K2, state=None):
        self.L = lib_kfdb()
        x1c = np.ascontiguousarray(x1c, np.float32)
        x2c = np.ascontiguousarray(x2c, np.float32)
        o1 = np.ascontiguousarray(octave1, np.int32)
        o2 = np.ascontiguousarray(octave2, np.int32)
        ls2 = np.ascontiguousarray(level_sigma2, np.float32)
        self.n = x1c.shape[0]
        st = np.ones(self.n, np.uint8) if state is None else np.ascontiguousarray(state, np.uint8)
        k1 = np.ascontiguousarray(K1, np.float32)
        k2 = np.ascontiguousarray(K2, np.float32)
        self.h = C.c_void_p(self.L.ref_sim3_create(C.c_int(self.n), _p(x1c), _p(x2c), _p(o1), _p(o2), _p(ls2), C.c_int(len(ls2)), _p(st), _p(k1), _p(k2)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_sim3_destroy(self.h)
            self.h = None

    def set_params(self, prob=0.99, min_inliers=6, max_its=300):
        self.L.ref_sim3_set_params(self.h, C.c_double(prob), C.c_int(min_inliers), C.c_int(max_its))

    def params(self):
        N, its = C.c_int(), C.c_int()
        self.L.ref_sim3_get_params(self.h, C.byref(N), C.byref(its), None, None)
        e1 = np.zeros(max(N.value, 1), np.uint64)
        e2 = np.zeros(max(N.value, 1), np.uint64)
        self.L.ref_sim3_get_params(self.h, C.byref(N), C.byref(its), _p(e1), _p(e2))
        return dict(N=N.value, max_its=its.value, max_err1=e1[:N.value], max_err2=e2[:N.value])

    def iterate(self, n_iterations):
        no_more, n_inl = C.c_int(), C.c_int()
        inl = np.zeros(max(self.n, 1), np.uint8)
        ok = self.L.ref_sim3_iterate(self.h, C.c_int(n_iterations), C.byref(no_more), _p(inl), C.byref(n_inl))
        return dict(ok=bool(ok), no_more=bool(no_more.value), n_inliers=n_inl.value, inliers=inl[:self.n].astype(bool))

    def state(self, N):
        it, best = C.c_int(), C.c_int()
        R, t = np.zeros(9, np.float32), np.zeros(3, np.float32)
        mask = np.zeros(max(N, 1), np.uint8)
        self.L.ref_sim3_state(self.h, C.byref(it), C.byref(best), _p(R), _p(t), _p(mask))
        return dict(iterations=it.value, best_inliers=best.value, R=R.reshape(3, 3), t=t, best_mask=mask[:N].astype(bool))

    def compute_and_check(self, idx3, N):
        idx3 = np.ascontiguousarray(idx3, np.int32)
        R, t = np.zeros(9, np.float32), np.zeros(3, np.float32)
        mask = np.zeros(max(N, 1), np.uint8)
        cnt = self.L.ref_sim3_compute_and_check(self.h, _p(idx3), _p(R), _p(t), _p(mask))
        return R.reshape(3, 3), t, cnt, mask[:N].astype(bool)


def eig_record(on=True):
    lib().ref_eig_record(C.c_int(1 if on else 0))


def eig_calls():
    """12 x 12 eigen-problems solved since eig_record(): one per compute_pose call (PnPsolver.cpp:380)"""
    return lib().ref_eig_calls()


def eig_take(max_calls=4096):
    out = np.zeros((max_calls, 12, 4), np.float64)
    n = lib().ref_eig_take(_p(out), C.c_int(max_calls))
    return out[:n]


class KfDb:
    """The reference's KeyFrameDatabase (KeyFrameDatabase.cpp) over its own DBoW2 BowVector / L1Scoring; the keyframes
    of `db` (synth.kf_database) are add()-ed in index order and keep their query marks between calls, as in a running
    system."""

    def __init__(self, db):
        self.L = lib_kfdb()
        self.K = int(db["K"])
        off = np.ascontiguousarray(db["bow_off"], np.int64)
        w = np.ascontiguousarray(db["bow_word"], np.uint32)
        v = np.ascontiguousarray(db["bow_val"], np.float64)
        cov = np.ascontiguousarray(db["covis"], np.int32)
        assert cov.shape == (self.K, 10)
        self.h = C.c_void_p(self.L.ref_kfdb_create(C.c_int(self.K), _p(off), _p(w), _p(v), _p(cov), C.c_uint(int(db["vocab"]) + 16)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_kfdb_destroy(self.h)
            self.h = None

    def reloc(self, qword, qval, frame_id):
        qw = np.ascontiguousarray(qword, np.uint32)
        qv = np.ascontiguousarray(qval, np.float64)
        out = np.empty(max(self.K, 1), np.int32)
        n = self.L.ref_kfdb_reloc(self.h, C.c_int(len(qw)), _p(qw), _p(qv), C.c_ulong(frame_id), _p(out), C.c_int(len(out)))
        return out[:n].tolist()

    def loop(self, q, query_id, conn, min_score):
        cn = np.ascontiguousarray(conn, np.int32)
        out = np.empty(max(self.K, 1), np.int32)
        n = self.L.ref_kfdb_loop(self.h, C.c_int(q), C.c_ulong(query_id), C.c_int(len(cn)), _p(cn), C.c_float(min_score), _p(out), C.c_int(len(out)))
        return out[:n].tolist()

    def reloc_scores(self):
        s = np.zeros(max(self.K, 1), np.float32)
        self.L.ref_kfdb_reloc_scores(self.h, _p(s))
        return s[:self.K]


def lib_kfdb():
    return lib()


def bow_l1_score(w1, v1, w2, v2):
    w1, w2 = np.ascontiguousarray(w1, np.uint32), np.ascontiguousarray(w2, np.uint32)
    v1, v2 = np.ascontiguousarray(v1, np.float64), np.ascontiguousarray(v2, np.float64)
    return lib_kfdb().ref_bow_l1_score(C.c_int(len(w1)), _p(w1), _p(v1), C.c_int(len(w2)), _p(w2), _p(v2))


class MLPnP:
    """The reference's MLPnPsolver (left out of its own build, CMakeLists.txt:75) over one frame."""

    def __init__(self, kp_xy, octave, level_sigma2, mp_xyz, K, state=None):
        self.L = lib_kfdb()
        kp_xy = np.ascontiguousarray(kp_xy, np.float32)
        mp_xyz = np.ascontiguousarray(mp_xyz, np.float32)
        octave = np.ascontiguousarray(octave, np.int32)
        ls2 = np.ascontiguousarray(level_sigma2, np.float32)
        self.n_kp = kp_xy.shape[0]
        st = np.ones(self.n_kp, np.uint8) if state is None else np.ascontiguousarray(state, np.uint8)
        self.h = C.c_void_p(self.L.ref_mlpnp_create(C.c_int(self.n_kp), _p(kp_xy), _p(octave), _p(ls2), C.c_int(len(ls2)), _p(mp_xyz), _p(st),
                                               C.c_float(K[0]), C.c_float(K[1]), C.c_float(K[2]), C.c_float(K[3])))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_mlpnp_destroy(self.h)
            self.h = None

    def set_params(self, prob=0.99, min_inliers=8, max_its=300, min_set=6, eps=0.4, th2=5.991):
        self.L.ref_mlpnp_set_params(self.h, C.c_double(prob), C.c_int(min_inliers), C.c_int(max_its), C.c_int(min_set), C.c_float(eps), C.c_float(th2))

    def params(self):
        N, mi, its = C.c_int(), C.c_int(), C.c_int()
        self.L.ref_mlpnp_get_params(self.h, C.byref(N), C.byref(mi), C.byref(its), None, None)
        max_err = np.zeros(max(N.value, 1), np.float32)
        kpi = np.zeros(max(N.value, 1), np.int32)
        self.L.ref_mlpnp_get_params(self.h, C.byref(N), C.byref(mi), C.byref(its), _p(max_err), _p(kpi))
        return dict(N=N.value, min_inliers=mi.value, max_its=its.value, max_err=max_err[:N.value], kp_index=kpi[:N.value])

    def iterate(self, n_iterations):
        no_more, n_inl = C.c_int(), C.c_int()
        inl = np.zeros(max(self.n_kp, 1), np.uint8)
        T = np.zeros(16, np.float32)
        ok = self.L.ref_mlpnp_iterate(self.h, C.c_int(n_iterations), C.byref(no_more), _p(inl), C.byref(n_inl), _p(T))
        return dict(ok=bool(ok), no_more=bool(no_more.value), n_inliers=n_inl.value, inliers=inl[:self.n_kp].astype(bool), T=T.reshape(4, 4))

    def state(self):
        it, best, refd = C.c_int(), C.c_int(), C.c_int()
        self.L.ref_mlpnp_state(self.h, C.byref(it), C.byref(best), C.byref(refd))
        return dict(iterations=it.value, best_inliers=best.value, refined_inliers=refd.value)

    def compute_pose(self, idx, cov=None):
        idx = np.ascontiguousarray(idx, np.int32)
        covc = None if cov is None else np.ascontiguousarray(cov, np.float64)
        R, t = np.zeros(9), np.zeros(3)
        self.L.ref_mlpnp_compute_pose(self.h, _p(idx), C.c_int(len(idx)), None if covc is None else _p(covc), _p(R), _p(t))
        return R.reshape(3, 3), t

    def check_inliers(self, R, t, N):
        R = np.ascontiguousarray(R, np.float64).reshape(-1)
        t = np.ascontiguousarray(t, np.float64)
        mask = np.zeros(max(N, 1), np.uint8)
        cnt = self.L.ref_mlpnp_check_inliers(self.h, _p(R), _p(t), _p(mask))
        return cnt, mask[:N].astype(bool)

    def res_jac(self, pt, nr, ns, w, t):
        a = [np.ascontiguousarray(x, np.float64) for x in (pt, nr, ns, w, t)]
        r, J = np.zeros(2), np.zeros(12)
        self.L.ref_mlpnp_res_jac(self.h, *[_p(x) for x in a], _p(r), _p(J))
        return r, J.reshape(2, 6)

    def rodrigues(self, w):
        w = np.ascontiguousarray(w, np.float64)
        R, wb = np.zeros(9), np.zeros(3)
        self.L.ref_mlpnp_rodrigues(self.h, _p(w), _p(R), _p(wb))
        return R.reshape(3, 3), wb


# ---- ORBmatcher (takes the oracle binding's struct wrappers: oracle_api.bow_features / kf_view) ----
def search_by_bow(q, t, nn_ratio=0.75, check_orientation=True, mode=0):
    n_out = t.st.n_feat if mode == 0 else q.st.n_feat
    out = np.empty(max(n_out, 1), np.int32)
    n = lib().ref_search_by_bow(C.byref(q.st), C.byref(t.st), C.c_float(nn_ratio), C.c_int(int(check_orientation)), C.c_int(mode), _p(out))
    return out[:n_out], n


def descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint32), np.ascontiguousarray(b, np.uint32)
    return lib().ref_descriptor_distance(_p(a), _p(b))


def search_by_sim3(kf1, kf2, K, R12, t12, th=7.5, matched12_in=None):
    K = np.ascontiguousarray(K, np.float32); R12 = np.ascontiguousarray(R12, np.float32).reshape(9); t12 = np.ascontiguousarray(t12, np.float32)
    mi = None if matched12_in is None else np.ascontiguousarray(matched12_in, np.int32)
    out = np.empty(max(kf1.st.n_feat, 1), np.int32)
    n = lib().ref_search_by_sim3(C.byref(kf1.st), C.byref(kf2.st), _p(K), _p(R12), _p(t12), C.c_float(th), None if mi is None else _p(mi), _p(out))
    return out[:kf1.st.n_feat].copy(), n


def search_by_projection(frame, kf, K, Rcw, tcw, th=10.0, orb_dist=100, check_orientation=True, occupied=None, already_found=None):
    K = np.ascontiguousarray(K, np.float32); R = np.ascontiguousarray(Rcw, np.float32).reshape(9); t = np.ascontiguousarray(tcw, np.float32)
    oc = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    af = None if already_found is None else np.ascontiguousarray(already_found, np.uint8)
    out = np.empty(max(frame.st.n_feat, 1), np.int32)
    n = lib().ref_search_by_projection(C.byref(frame.st), C.byref(kf.st), _p(K), _p(R), _p(t), C.c_float(th), C.c_int(orb_dist), C.c_int(int(check_orientation)),
                                       None if oc is None else _p(oc), None if af is None else _p(af), _p(out))
    return out[:frame.st.n_feat].copy(), n
