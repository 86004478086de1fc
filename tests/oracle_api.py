"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the checker; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.  The product path
(orb-slam2-optimized_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

FLAG_STALE_ROWS = 1
FLAG_EXHAUSTIVE = 2
FLAG_MLPNP_DISCARD_REFINE = 4
FLAG_EPNP_QR_NULLSPACE = 8


class RansacParams(C.Structure):
    _fields_ = [("prob", C.c_double), ("min_inliers", C.c_int), ("max_its", C.c_int),
                ("min_set", C.c_int), ("eps", C.c_float), ("th2", C.c_float)]


class PnPProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("p3d", C.c_void_p), ("p2d", C.c_void_p), ("sigma2", C.c_void_p),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double)]


class Sim3Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("x1c", C.c_void_p), ("x2c", C.c_void_p), ("sigma2_1", C.c_void_p),
                ("sigma2_2", C.c_void_p), ("K1", C.c_float * 4), ("K2", C.c_float * 4), ("fix_scale", C.c_int)]


class MLPnPProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("p3d", C.c_void_p), ("p2d", C.c_void_p), ("sigma2", C.c_void_p),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("cov", C.c_void_p)]


class Result(C.Structure):
    _fields_ = [("ok", C.c_int), ("no_more", C.c_int), ("n_inliers", C.c_int), ("best_hyp", C.c_int),
                ("refined", C.c_int), ("n_hyp", C.c_int), ("n_refines", C.c_int), ("n_failed_refines", C.c_int),
                ("best_count", C.c_int), ("T", C.c_float * 16), ("scale", C.c_float)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "T"}
        d["T"] = np.array(self.T, np.float32).reshape(4, 4)
        return d


def _host_has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            return " fma " in f.read().replace("\n", " ")
    except OSError:
        return False


def build():
    """make; a library built with -mfma on another host is rebuilt when this host has no FMA3"""
    flag_file = os.path.join(ORACLE_DIR, ".fmaflag")
    if os.path.exists(flag_file) and "-mfma" in open(flag_file).read() and not _host_has_fma():
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "clean"], check=True)
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        flag_file = os.path.join(ORACLE_DIR, ".fmaflag")
        if not os.path.exists(path) or (os.path.exists(flag_file) and "-mfma" in open(flag_file).read() and not _host_has_fma()):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_epnp_pose.restype = C.c_double
        _LIB.orc_epnp_flops.restype = C.c_double
        _LIB.orc_epnp_pose_mode.restype = C.c_double
        _LIB.orc_epnp_flops_mode.restype = C.c_double
        for f in ("orc_pnp_batch", "orc_sim3_batch", "orc_mlpnp_batch", "orc_pnp_score_timed"):
            getattr(_LIB, f).restype = C.c_double
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def index_table(seed, n, k, H):
    out = np.empty((H, k), np.uint32)
    lib().orc_index_table(C.c_uint(seed), C.c_int(n), C.c_int(k), C.c_int(H), _p(out))
    return out


def random_ints(seed, lo, hi, count):
    L = lib()
    L.orc_rng_seed(C.c_uint(seed))
    return [L.orc_random_int(C.c_int(lo), C.c_int(hi)) for _ in range(count)]


def jacobi_eig(a, dtype=np.float64):
    a = np.array(a, dtype, order="C", copy=True)
    n = a.shape[0]
    w = np.empty(n, dtype)
    v = np.empty((n, n), dtype)
    fn = lib().orc_jacobi_eig_d if dtype == np.float64 else lib().orc_jacobi_eig_f
    fn(C.c_int(n), _p(a), _p(w), _p(v))
    return w, v


def jacobi_lowest(a, nv):
    a = np.array(a, np.float64, order="C", copy=True)
    n = a.shape[0]
    w = np.empty(nv)
    v = np.empty((n, nv))
    lib().orc_jacobi_lowest_d(C.c_int(n), C.c_int(nv), _p(a), _p(w), _p(v))
    return w, v


def svd_lstsq(L, b):
    L = np.ascontiguousarray(L, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.empty(L.shape[1])
    lib().orc_svd_lstsq_d(C.c_int(L.shape[0]), C.c_int(L.shape[1]), _p(L), _p(b), _p(x))
    return x


def inv3(m):
    m = np.ascontiguousarray(m, np.float64)
    o = np.empty((3, 3))
    lib().orc_inv3_d(_p(m), _p(o))
    return o


def polar3(m):
    m = np.ascontiguousarray(m, np.float64)
    o = np.empty((3, 3))
    lib().orc_polar3_d(_p(m), _p(o))
    return o


def rank3(m):
    m = np.ascontiguousarray(m, np.float64)
    return lib().orc_rank3_fullpiv_d(_p(m))


def ldlt6_solve(a, g):
    a = np.ascontiguousarray(a, np.float64)
    g = np.ascontiguousarray(g, np.float64)
    x = np.empty(6)
    lib().orc_ldlt6_solve_d(_p(a), _p(g), _p(x))
    return x


def ransac_setup_pnp(n, prm: RansacParams):
    mi, it = C.c_int(), C.c_int()
    lib().orc_pnp_ransac_setup(C.c_int(n), C.byref(prm), C.byref(mi), C.byref(it))
    return mi.value, it.value


def ransac_setup_sim3(n, prob, min_inl, max_its):
    it = C.c_int()
    lib().orc_sim3_ransac_setup(C.c_int(n), C.c_double(prob), C.c_int(min_inl), C.c_int(max_its), C.byref(it))
    return it.value


class _Keep:
    """a ctypes struct plus the numpy arrays it points into"""

    def __init__(self, st, *arrs):
        self.st = st
        self.arrs = arrs


def pnp_problem(p3d, p2d, sigma2, K):
    p3d, p2d, sigma2 = _f32(p3d), _f32(p2d), _f32(sigma2)
    st = PnPProblem(p3d.shape[0], _p(p3d), _p(p2d), _p(sigma2), K[0], K[1], K[2], K[3])
    return _Keep(st, p3d, p2d, sigma2)


def mlpnp_problem(p3d, p2d, sigma2, K, cov=None):
    p3d, p2d, sigma2 = _f32(p3d), _f32(p2d), _f32(sigma2)
    covc = None if cov is None else np.ascontiguousarray(cov, np.float64)
    st = MLPnPProblem(p3d.shape[0], _p(p3d), _p(p2d), _p(sigma2), K[0], K[1], K[2], K[3],
                      None if covc is None else _p(covc))
    return _Keep(st, p3d, p2d, sigma2, covc)


def sim3_problem(x1c, x2c, s1, s2, K1, K2, fix_scale=True):
    x1c, x2c, s1, s2 = _f32(x1c), _f32(x2c), _f32(s1), _f32(s2)
    st = Sim3Problem(x1c.shape[0], _p(x1c), _p(x2c), _p(s1), _p(s2), (C.c_float * 4)(*K1), (C.c_float * 4)(*K2),
                     1 if fix_scale else 0)
    return _Keep(st, x1c, x2c, s1, s2)


def params(prob=0.99, min_inliers=8, max_its=300, min_set=4, eps=0.4, th2=5.991):
    return RansacParams(prob, min_inliers, max_its, min_set, eps, th2)


def epnp_pose(pb: _Keep, idx, flags=0):
    """flags: FLAG_EPNP_QR_NULLSPACE selects the QR null space for 4-point sets"""
    idx = np.ascontiguousarray(idx, np.uint32)
    R = np.empty(9, np.float32)
    t = np.empty(3, np.float32)
    err = lib().orc_epnp_pose_mode(C.byref(pb.st), _p(idx), C.c_int(len(idx)), C.c_int(flags), _p(R), _p(t))
    return R.reshape(3, 3), t, err


def epnp_mtm(pb: _Keep, idx):
    """M^T M (12 x 12) of the EPnP system on the subset idx"""
    idx = np.ascontiguousarray(idx, np.uint32)
    out = np.zeros((12, 12), np.float64)
    lib().orc_epnp_mtm(C.byref(pb.st), _p(idx), C.c_int(len(idx)), _p(out))
    return out


def epnp_pose_basis(pb: _Keep, idx, U):
    """compute_pose downstream of a caller-supplied null-space basis U (12 x 4)"""
    idx = np.ascontiguousarray(idx, np.uint32)
    U = np.ascontiguousarray(U, np.float64)
    R, t = np.empty(9, np.float32), np.empty(3, np.float32)
    lib().orc_epnp_pose_basis.restype = C.c_double
    err = lib().orc_epnp_pose_basis(C.byref(pb.st), _p(idx), C.c_int(len(idx)), _p(U), _p(R), _p(t))
    return R.reshape(3, 3), t, err


def epnp_flops(pb: _Keep, table, flags=0):
    table = np.ascontiguousarray(table, np.uint32)
    return lib().orc_epnp_flops_mode(C.byref(pb.st), _p(table), C.c_int(table.shape[0]), C.c_int(table.shape[1]), C.c_int(flags))


def pnp_check_inliers(pb: _Keep, max_err, R, t):
    max_err, R, t = _f32(max_err), _f32(R).reshape(-1), _f32(t)
    n = pb.st.n
    mask = np.empty(n, np.uint8)
    err2 = np.empty(n, np.float32)
    cnt = lib().orc_pnp_check_inliers(C.byref(pb.st), _p(max_err), _p(R), _p(t), _p(mask), _p(err2))
    return cnt, mask.astype(bool), err2


def pnp_score(pb: _Keep, max_err, poses, want_masks=True):
    max_err, poses = _f32(max_err), _f32(poses)
    H = poses.shape[0]
    masks = np.empty((H, pb.st.n), np.uint8) if want_masks else None
    counts = np.empty(H, np.int32)
    lib().orc_pnp_score(C.byref(pb.st), _p(max_err), C.c_int(H), _p(poses), None if masks is None else _p(masks), _p(counts))
    return counts, masks


def pnp_ransac(pb: _Keep, prm: RansacParams, table, flags=0, per_hyp=False):
    table = np.ascontiguousarray(table, np.uint32)
    n = pb.st.n
    res = Result()
    mask = np.zeros(max(n, 1), np.uint8)
    _, H = ransac_setup_pnp(n, prm)
    hc = np.full(H, -1, np.int32) if per_hyp else None
    hp = np.full((H, 12), np.nan, np.float32) if per_hyp else None
    lib().orc_pnp_ransac(C.byref(pb.st), C.byref(prm), _p(table), C.c_int(flags), C.byref(res), _p(mask),
                         None if hc is None else _p(hc), None if hp is None else _p(hp))
    out = res.as_dict()
    out["mask"] = mask[:n].astype(bool)
    if per_hyp:
        out["hyp_counts"], out["hyp_pose"] = hc, hp
    return out


def mlpnp_pose(pb: _Keep, idx):
    idx = np.ascontiguousarray(idx, np.uint32)
    R = np.empty(9)
    t = np.empty(3)
    lib().orc_mlpnp_pose(C.byref(pb.st), _p(idx), C.c_int(len(idx)), _p(R), _p(t))
    return R.reshape(3, 3), t


def mlpnp_check_inliers(pb: _Keep, max_err, R, t):
    max_err = _f32(max_err)
    R = np.ascontiguousarray(R, np.float64).reshape(-1)
    t = np.ascontiguousarray(t, np.float64)
    n = pb.st.n
    mask = np.empty(n, np.uint8)
    err2 = np.empty(n, np.float32)
    cnt = lib().orc_mlpnp_check_inliers(C.byref(pb.st), _p(max_err), _p(R), _p(t), _p(mask), _p(err2))
    return cnt, mask.astype(bool), err2


def mlpnp_ransac(pb: _Keep, prm: RansacParams, table, flags=0, per_hyp=False):
    table = np.ascontiguousarray(table, np.uint32)
    n = pb.st.n
    res = Result()
    mask = np.zeros(max(n, 1), np.uint8)
    _, H = ransac_setup_pnp(n, prm)
    hc = np.full(H, -1, np.int32) if per_hyp else None
    hp = np.full((H, 12), np.nan, np.float64) if per_hyp else None
    lib().orc_mlpnp_ransac(C.byref(pb.st), C.byref(prm), _p(table), C.c_int(flags), C.byref(res), _p(mask),
                           None if hc is None else _p(hc), None if hp is None else _p(hp))
    out = res.as_dict()
    out["mask"] = mask[:n].astype(bool)
    if per_hyp:
        out["hyp_counts"], out["hyp_pose"] = hc, hp
    return out


def mlpnp_res_jac(pt, nr, ns, w, t):
    a = [np.ascontiguousarray(x, np.float64) for x in (pt, nr, ns, w, t)]
    r = np.empty(2)
    J = np.empty((2, 6))
    lib().orc_mlpnp_res_jac(*[_p(x) for x in a], _p(r), _p(J))
    return r, J


def rodrigues2rot(w):
    w = np.ascontiguousarray(w, np.float64)
    R = np.empty((3, 3))
    lib().orc_rodrigues2rot(_p(w), _p(R))
    return R


def rot2rodrigues(R):
    R = np.ascontiguousarray(R, np.float64)
    w = np.empty(3)
    lib().orc_rot2rodrigues(_p(R), _p(w))
    return w


def sim3_compute(P1, P2, fix_scale=True):
    P1, P2 = _f32(P1), _f32(P2)
    R = np.empty(9, np.float32)
    t = np.empty(3, np.float32)
    s = C.c_float()
    lib().orc_sim3_compute(_p(P1), _p(P2), C.c_int(1 if fix_scale else 0), _p(R), _p(t), C.byref(s))
    return R.reshape(3, 3), t, s.value


def sim3_check_inliers(pb: _Keep, R, t, s=1.0):
    R, t = _f32(R).reshape(-1), _f32(t)
    n = pb.st.n
    mask = np.empty(n, np.uint8)
    err = np.empty((n, 2), np.float32)
    cnt = lib().orc_sim3_check_inliers(C.byref(pb.st), _p(R), _p(t), C.c_float(s), _p(mask), _p(err))
    return cnt, mask.astype(bool), err


def sim3_ransac(pb: _Keep, prob, min_inliers, max_its, table, flags=0, per_hyp=False):
    table = np.ascontiguousarray(table, np.uint32)
    n = pb.st.n
    res = Result()
    mask = np.zeros(max(n, 1), np.uint8)
    H = ransac_setup_sim3(n, prob, min_inliers, max_its) if n > 0 else 1
    hc = np.full(H, -1, np.int32) if per_hyp else None
    hp = np.full((H, 13), np.nan, np.float32) if per_hyp else None
    lib().orc_sim3_ransac(C.byref(pb.st), C.c_double(prob), C.c_int(min_inliers), C.c_int(max_its), _p(table),
                          C.c_int(flags), C.byref(res), _p(mask), None if hc is None else _p(hc),
                          None if hp is None else _p(hp))
    out = res.as_dict()
    out["mask"] = mask[:n].astype(bool)
    if per_hyp:
        out["hyp_counts"], out["hyp_pose"] = hc, hp
    return out


def _tables_ptr(tables):
    arrs = [np.ascontiguousarray(t, np.uint32) for t in tables]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    return arrs, ptrs


def pnp_batch(pbs, prm, tables, flags=0, nthreads=1):
    """returns (seconds, evals_done, list of result dicts)"""
    n = len(pbs)
    arr = (PnPProblem * n)(*[p.st for p in pbs])
    keep, ptrs = _tables_ptr(tables)
    res = (Result * n)()
    ev = C.c_longlong()
    dt = lib().orc_pnp_batch(C.c_int(n), arr, C.byref(prm), ptrs, C.c_int(flags), C.c_int(nthreads), res, C.byref(ev))
    return dt, ev.value, [r.as_dict() for r in res]


def _batch_masks(fn_name, arr_t, pbs, prm, tables, flags, nthreads):
    """threaded batch driver that also returns the inlier masks: (list of result dicts, list of bool masks)"""
    n = len(pbs)
    arr = (arr_t * n)(*[p.st for p in pbs])
    keep, ptrs = _tables_ptr(tables)
    res = (Result * n)()
    masks = [np.zeros(max(int(p.st.n), 1), np.uint8) for p in pbs]
    mptr = (C.c_void_p * n)(*[m.ctypes.data for m in masks])
    ev = C.c_longlong()
    fn = getattr(lib(), fn_name)
    fn.restype = C.c_double
    fn(C.c_int(n), arr, C.byref(prm), ptrs, C.c_int(flags), C.c_int(nthreads), res, mptr, C.byref(ev))
    return [r.as_dict() for r in res], [m[:int(p.st.n)].astype(bool) for m, p in zip(masks, pbs)]


def pnp_batch_masks(pbs, prm, tables, flags=0, nthreads=1):
    return _batch_masks("orc_pnp_batch_masks", PnPProblem, pbs, prm, tables, flags, nthreads)


def mlpnp_batch_masks(pbs, prm, tables, flags=0, nthreads=1):
    return _batch_masks("orc_mlpnp_batch_masks", MLPnPProblem, pbs, prm, tables, flags, nthreads)


def sim3_batch(pbs, prob, min_inliers, max_its, tables, flags=0, nthreads=1):
    n = len(pbs)
    arr = (Sim3Problem * n)(*[p.st for p in pbs])
    keep, ptrs = _tables_ptr(tables)
    res = (Result * n)()
    ev = C.c_longlong()
    dt = lib().orc_sim3_batch(C.c_int(n), arr, C.c_double(prob), C.c_int(min_inliers), C.c_int(max_its), ptrs,
                              C.c_int(flags), C.c_int(nthreads), res, C.byref(ev))
    return dt, ev.value, [r.as_dict() for r in res]


def mlpnp_batch(pbs, prm, tables, flags=0, nthreads=1):
    n = len(pbs)
    arr = (MLPnPProblem * n)(*[p.st for p in pbs])
    keep, ptrs = _tables_ptr(tables)
    res = (Result * n)()
    ev = C.c_longlong()
    dt = lib().orc_mlpnp_batch(C.c_int(n), arr, C.byref(prm), ptrs, C.c_int(flags), C.c_int(nthreads), res, C.byref(ev))
    return dt, ev.value, [r.as_dict() for r in res]


def pnp_score_timed(pb: _Keep, max_err, poses, nthreads=1):
    max_err, poses = _f32(max_err), _f32(poses)
    counts = np.empty(poses.shape[0], np.int32)
    dt = lib().orc_pnp_score_timed(C.byref(pb.st), _p(max_err), C.c_int(poses.shape[0]), _p(poses), C.c_int(nthreads), _p(counts))
    return dt, counts


class BowFeatures(C.Structure):
    _fields_ = [("n_feat", C.c_int), ("desc", C.c_void_p), ("angle", C.c_void_p), ("valid", C.c_void_p),
                ("n_nodes", C.c_int), ("node_ids", C.c_void_p), ("node_off", C.c_void_p), ("node_feat", C.c_void_p)]


def bow_features(f):
    desc = np.ascontiguousarray(f["desc"], np.uint32).reshape(-1, 8)
    ang = np.ascontiguousarray(f["angle"], np.float32)
    val = None if f.get("valid") is None else np.ascontiguousarray(f["valid"], np.uint8)
    nid = np.ascontiguousarray(f["node_ids"], np.uint32)
    noff = np.ascontiguousarray(f["node_off"], np.int32)
    nfe = np.ascontiguousarray(f["node_feat"], np.uint32)
    st = BowFeatures(desc.shape[0], _p(desc), _p(ang), None if val is None else _p(val), len(nid), _p(nid), _p(noff), _p(nfe))
    return _Keep(st, desc, ang, val, nid, noff, nfe)


def descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint32), np.ascontiguousarray(b, np.uint32)
    return lib().orc_descriptor_distance(_p(a), _p(b))


def three_maxima(histo):
    h = np.ascontiguousarray(histo, np.int32)
    i1, i2, i3 = C.c_int(), C.c_int(), C.c_int()
    lib().orc_three_maxima(_p(h), C.c_int(len(h)), C.byref(i1), C.byref(i2), C.byref(i3))
    return i1.value, i2.value, i3.value


def search_by_bow(q: _Keep, t: _Keep, nn_ratio=0.75, check_orientation=True, mode=0):
    """(match array, nmatches): mode 0 indexed by target (frame) feature, mode 1 by query (KF1) feature"""
    n_out = t.st.n_feat if mode == 0 else q.st.n_feat
    out = np.empty(max(n_out, 1), np.int32)
    n = lib().orc_search_by_bow(C.byref(q.st), C.byref(t.st), C.c_float(nn_ratio), C.c_int(int(check_orientation)), C.c_int(mode), _p(out))
    return out[:n_out], n


class KfDb(C.Structure):
    _fields_ = [("K", C.c_int), ("bow_off", C.c_void_p), ("bow_word", C.c_void_p), ("bow_val", C.c_void_p), ("covis", C.c_void_p)]


def kfdb(db):
    """db: dict(bow_off int64 [K+1], bow_word uint32, bow_val float64, covis int32 [K,10])"""
    off = np.ascontiguousarray(db["bow_off"], np.int64)
    w = np.ascontiguousarray(db["bow_word"], np.uint32)
    v = np.ascontiguousarray(db["bow_val"], np.float64)
    cv = np.ascontiguousarray(db["covis"], np.int32).reshape(-1, 10)
    st = KfDb(len(off) - 1, _p(off), _p(w), _p(v), _p(cv))
    return _Keep(st, off, w, v, cv)


def bow_l1_score(w1, v1, w2, v2):
    w1, w2 = np.ascontiguousarray(w1, np.uint32), np.ascontiguousarray(w2, np.uint32)
    v1, v2 = np.ascontiguousarray(v1, np.float64), np.ascontiguousarray(v2, np.float64)
    f = lib().orc_bow_l1_score
    f.restype = C.c_double
    return f(C.c_int(len(w1)), _p(w1), _p(v1), C.c_int(len(w2)), _p(w2), _p(v2))


def kfdb_last_query_seconds():
    f = lib().orc_kfdb_last_query_seconds
    f.restype = C.c_double
    return f()


def detect_candidates(db: _Keep, qword, qval, mode=0, conn=None, min_score=0.0, score_state=None):
    """KeyFrameDatabase::DetectRelocalizationCandidates (mode 0) / DetectLoopCandidates (mode 1): candidate keyframe
    indices in the reference's order.  score_state (float32 [K]) is updated in place (mode 0)."""
    qw = np.ascontiguousarray(qword, np.uint32)
    qv = np.ascontiguousarray(qval, np.float64)
    cn = np.ascontiguousarray(conn if conn is not None else [], np.int32)
    out = np.empty(max(db.st.K, 1), np.int32)
    if score_state is not None:
        assert score_state.dtype == np.float32 and score_state.flags.c_contiguous
    n = lib().orc_detect_candidates(C.byref(db.st), C.c_int(mode), C.c_int(len(qw)), _p(qw), _p(qv), C.c_int(len(cn)), _p(cn) if len(cn) else None,
                                    C.c_float(min_score), None if score_state is None else _p(score_state), _p(out), C.c_int(len(out)))
    assert n >= 0
    return out[:n].copy()


class KfView(C.Structure):
    _fields_ = [("n_feat", C.c_int), ("kp_xy", C.c_void_p), ("kp_octave", C.c_void_p), ("kp_angle", C.c_void_p), ("desc", C.c_void_p), ("mp_valid", C.c_void_p),
                ("mp_xyz", C.c_void_p), ("mp_desc", C.c_void_p), ("mp_maxdist", C.c_void_p), ("mp_mindist", C.c_void_p),
                ("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("bounds", C.c_float * 4), ("grid_cols", C.c_int), ("grid_rows", C.c_int),
                ("grid_w_inv", C.c_float), ("grid_h_inv", C.c_float), ("grid_off", C.c_void_p), ("grid_idx", C.c_void_p),
                ("n_levels", C.c_int), ("scale_factors", C.c_void_p), ("log_scale_factor", C.c_float)]


def kf_view(v):
    a = dict(kp_xy=np.ascontiguousarray(v["kp_xy"], np.float32), kp_octave=np.ascontiguousarray(v["kp_octave"], np.int32),
             kp_angle=np.ascontiguousarray(v["kp_angle"], np.float32),
             desc=np.ascontiguousarray(v["desc"], np.uint32), mp_valid=np.ascontiguousarray(v["mp_valid"], np.uint8),
             mp_xyz=np.ascontiguousarray(v["mp_xyz"], np.float32), mp_desc=np.ascontiguousarray(v["mp_desc"], np.uint32),
             mp_maxdist=np.ascontiguousarray(v["mp_maxdist"], np.float32), mp_mindist=np.ascontiguousarray(v["mp_mindist"], np.float32),
             grid_off=np.ascontiguousarray(v["grid_off"], np.int32), grid_idx=np.ascontiguousarray(v["grid_idx"], np.int32),
             scale_factors=np.ascontiguousarray(v["scale_factors"], np.float32))
    st = KfView(int(v["n_feat"]), _p(a["kp_xy"]), _p(a["kp_octave"]), _p(a["kp_angle"]), _p(a["desc"]), _p(a["mp_valid"]), _p(a["mp_xyz"]), _p(a["mp_desc"]),
                _p(a["mp_maxdist"]), _p(a["mp_mindist"]), (C.c_float * 9)(*np.asarray(v["Rcw"], np.float32).reshape(-1)),
                (C.c_float * 3)(*np.asarray(v["tcw"], np.float32).reshape(-1)), (C.c_float * 4)(*np.asarray(v["bounds"], np.float32)),
                int(v["grid_cols"]), int(v["grid_rows"]), float(v["grid_w_inv"]), float(v["grid_h_inv"]), _p(a["grid_off"]), _p(a["grid_idx"]),
                int(v["n_levels"]), _p(a["scale_factors"]), float(v["log_scale_factor"]))
    return _Keep(st, *a.values())


def features_in_area(kf: _Keep, x, y, r):
    out = np.empty(max(kf.st.n_feat, 1), np.int32)
    n = lib().orc_features_in_area(C.byref(kf.st), C.c_float(x), C.c_float(y), C.c_float(r), _p(out))
    return out[:n].copy()


def predict_scale(max_distance, current_dist, log_scale_factor, n_levels):
    return lib().orc_predict_scale(C.c_float(max_distance), C.c_float(current_dist), C.c_float(log_scale_factor), C.c_int(n_levels))


def search_by_projection(frame: _Keep, kf: _Keep, K, Rcw, tcw, th=10.0, orb_dist=100, check_orientation=True, occupied=None, already_found=None):
    """ORBmatcher::SearchByProjection(Frame, KeyFrame, sAlreadyFound, th, ORBdist): (frame_match [N_frame], nmatches)"""
    K = np.ascontiguousarray(K, np.float32); R = np.ascontiguousarray(Rcw, np.float32).reshape(9); t = np.ascontiguousarray(tcw, np.float32)
    oc = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    af = None if already_found is None else np.ascontiguousarray(already_found, np.uint8)
    out = np.empty(max(frame.st.n_feat, 1), np.int32)
    n = lib().orc_search_by_projection(C.byref(frame.st), C.byref(kf.st), _p(K), _p(R), _p(t), C.c_float(th), C.c_int(orb_dist),
                                       C.c_int(int(check_orientation)), None if oc is None else _p(oc), None if af is None else _p(af), _p(out))
    assert n >= 0
    return out[:frame.st.n_feat].copy(), n


def search_by_sim3(kf1: _Keep, kf2: _Keep, K, R12, t12, th=7.5, matched12_in=None, s12=1.0):
    """ORBmatcher::SearchBySim3: (match12 [N1] = KF2 feature newly matched to each KF1 feature or -1, nFound)"""
    K = np.ascontiguousarray(K, np.float32); R12 = np.ascontiguousarray(R12, np.float32).reshape(9); t12 = np.ascontiguousarray(t12, np.float32)
    mi = None if matched12_in is None else np.ascontiguousarray(matched12_in, np.int32)
    out = np.empty(max(kf1.st.n_feat, 1), np.int32)
    n = lib().orc_search_by_sim3(C.byref(kf1.st), C.byref(kf2.st), _p(K), _p(R12), _p(t12), C.c_float(s12), C.c_float(th), None if mi is None else _p(mi), _p(out))
    assert n >= 0
    return out[:kf1.st.n_feat].copy(), n


class PoseOptProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("p3d", C.c_void_p), ("obs", C.c_void_p), ("inv_sigma2", C.c_void_p),
                ("K", C.c_float * 5), ("Rcw", C.c_float * 9), ("tcw", C.c_float * 3)]


class PoseOptResult(C.Structure):
    _fields_ = [("n_inliers", C.c_int32), ("n_bad", C.c_int32), ("rounds", C.c_int32), ("iterations", C.c_int32),
                ("trials", C.c_int32), ("reserved", C.c_int32), ("R", C.c_double * 9), ("t", C.c_double * 3),
                ("Rf", C.c_float * 9), ("tf", C.c_float * 3)]


def poseopt_problem(p3d, obs, inv_sigma2, K, Rcw, tcw):
    p3d, obs, inv_sigma2 = _f32(p3d), _f32(obs), _f32(inv_sigma2)
    st = PoseOptProblem(int(p3d.shape[0]), _p(p3d), _p(obs), _p(inv_sigma2), (C.c_float * 5)(*[float(k) for k in K]),
                        (C.c_float * 9)(*[float(x) for x in np.asarray(Rcw, np.float32).ravel()]),
                        (C.c_float * 3)(*[float(x) for x in np.asarray(tcw, np.float32).ravel()]))
    return _Keep(st, p3d, obs, inv_sigma2)


def pose_optimization(pb: _Keep):
    """Optimizer::PoseOptimization on one frame -> (dict, outlier flags uint8 [n])"""
    res = PoseOptResult()
    out = np.zeros(max(pb.st.n, 1), np.uint8)
    lib().orc_pose_optimization(C.byref(pb.st), C.byref(res), _p(out))
    d = {k: getattr(res, k) for k in ("n_inliers", "n_bad", "rounds", "iterations", "trials")}
    d["R"] = np.array(res.R).reshape(3, 3)
    d["t"] = np.array(res.t)
    d["Rf"] = np.array(res.Rf, np.float32).reshape(3, 3)
    d["tf"] = np.array(res.tf, np.float32)
    return d, out[:pb.st.n]


class Sim3OptProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("x1c", C.c_void_p), ("x2c", C.c_void_p), ("obs1", C.c_void_p), ("obs2", C.c_void_p),
                ("inv_sigma2_1", C.c_void_p), ("inv_sigma2_2", C.c_void_p), ("K1", C.c_float * 4), ("K2", C.c_float * 4),
                ("S12", C.c_float * 13), ("th2", C.c_float), ("fix_scale", C.c_int)]


class Sim3OptResult(C.Structure):
    _fields_ = [("n_inliers", C.c_int32), ("n_bad", C.c_int32), ("optimized", C.c_int32), ("iterations", C.c_int32),
                ("trials", C.c_int32), ("reserved", C.c_int32), ("R", C.c_double * 9), ("t", C.c_double * 3), ("s", C.c_double),
                ("q", C.c_double * 4)]


def sim3opt_problem(x1c, x2c, obs1, obs2, is1, is2, K1, K2, S12, th2=10.0, fix_scale=True):
    arrs = [_f32(a) for a in (x1c, x2c, obs1, obs2, is1, is2)]
    st = Sim3OptProblem(int(arrs[0].shape[0]), *[_p(a) for a in arrs], (C.c_float * 4)(*[float(k) for k in K1]),
                        (C.c_float * 4)(*[float(k) for k in K2]), (C.c_float * 13)(*[float(x) for x in np.asarray(S12, np.float32).ravel()]),
                        C.c_float(th2), 1 if fix_scale else 0)
    return _Keep(st, *arrs)


def optimize_sim3(pb: _Keep):
    """Optimizer::OptimizeSim3 on one keyframe pair -> (dict, removed flags uint8 [n])"""
    res = Sim3OptResult()
    rem = np.zeros(max(pb.st.n, 1), np.uint8)
    lib().orc_optimize_sim3(C.byref(pb.st), C.byref(res), _p(rem))
    d = {k: getattr(res, k) for k in ("n_inliers", "n_bad", "optimized", "iterations", "trials", "s")}
    d["R"] = np.array(res.R).reshape(3, 3)
    d["t"] = np.array(res.t)
    d["q"] = np.array(res.q)
    return d, rem[:pb.st.n]
