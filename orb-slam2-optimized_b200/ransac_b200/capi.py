"""ctypes binding of libransac_b200.so (the C ABI in include/ransac_b200.h).

This is plumbing for bench.py / tests; the product is the shared library.  There is no
fallback: if the library is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("RSAC_LIB", os.path.join(_PKG, "libransac_b200.so"))   # RSAC_LIB: tuning variants only
_LIB = None

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_STATE, ERR_ALLOC = range(6)
FLAG_KEEP_MASKS = 1
FLAG_MLPNP_DISCARD_REFINE = 4
FLAG_EPNP_EIGEN = 8          # 4-point EPnP null space from the 12x12 eigen-solve (default: Householder QR)
FLAG_EARLY_EXIT = 16         # PnP: hypotheses in phases, nothing behind the reference's stopping point is computed
STAGE_PACK, STAGE_RNG, STAGE_SOLVE, STAGE_SCORE, STAGE_SELECT = range(5)
STAGE_NAMES = ["pack", "rng", "solve", "score", "select"]


class RansacParams(C.Structure):
    _fields_ = [("prob", C.c_double), ("min_inliers", C.c_int32), ("max_its", C.c_int32),
                ("min_set", C.c_int32), ("eps", C.c_float), ("th2", C.c_float)]


class Sim3Params(C.Structure):
    _fields_ = [("prob", C.c_double), ("min_inliers", C.c_int32), ("max_its", C.c_int32), ("fix_scale", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("ok", C.c_int32), ("no_more", C.c_int32), ("n_inliers", C.c_int32), ("best_hyp", C.c_int32),
                ("refined", C.c_int32), ("n_refines", C.c_int32), ("best_count", C.c_int32), ("n_hyp", C.c_int32),
                ("R", C.c_float * 9), ("t", C.c_float * 3), ("s", C.c_float), ("problem", C.c_int32),
                ("reserved", C.c_int32 * 2)]


RESULT_DTYPE = np.dtype([("ok", "<i4"), ("no_more", "<i4"), ("n_inliers", "<i4"), ("best_hyp", "<i4"),
                         ("refined", "<i4"), ("n_refines", "<i4"), ("best_count", "<i4"), ("n_hyp", "<i4"),
                         ("R", "<f4", (9,)), ("t", "<f4", (3,)), ("s", "<f4"), ("problem", "<i4"),
                         ("reserved", "<i4", (2,))])
assert RESULT_DTYPE.itemsize == 96 == C.sizeof(Result)


class DeviceInfo(C.Structure):
    _fields_ = [("sm_count", C.c_int32), ("sm_clock_khz", C.c_int32), ("mem_clock_khz", C.c_int32),
                ("cc_major", C.c_int32), ("cc_minor", C.c_int32), ("total_mem", C.c_uint64), ("name", C.c_char * 64)]


class PnPBatch(C.Structure):
    _fields_ = [("C", C.c_int32), ("offsets", C.c_void_p), ("p3d", C.c_void_p), ("p2d", C.c_void_p),
                ("sigma2", C.c_void_p), ("K", C.c_void_p), ("params", C.c_void_p), ("n_params", C.c_int32),
                ("seeds", C.c_void_p), ("tables", C.c_void_p), ("table_offsets", C.c_void_p)]


class PnPIndexedBatch(C.Structure):
    _fields_ = [("n_keypoints", C.c_int32), ("kp_uv", C.c_void_p), ("kp_sigma2", C.c_void_p), ("n_mappoints", C.c_int32),
                ("mp_xyz", C.c_void_p), ("C", C.c_int32), ("offsets", C.c_void_p), ("kp_idx", C.c_void_p), ("mp_idx", C.c_void_p),
                ("K", C.c_void_p), ("params", C.c_void_p), ("n_params", C.c_int32), ("seeds", C.c_void_p), ("tables", C.c_void_p),
                ("table_offsets", C.c_void_p)]


class Sim3Batch(C.Structure):
    _fields_ = [("C", C.c_int32), ("offsets", C.c_void_p), ("x1c", C.c_void_p), ("x2c", C.c_void_p),
                ("sigma2_1", C.c_void_p), ("sigma2_2", C.c_void_p), ("K1", C.c_void_p), ("K2", C.c_void_p),
                ("params", C.c_void_p), ("n_params", C.c_int32), ("seeds", C.c_void_p), ("tables", C.c_void_p),
                ("table_offsets", C.c_void_p)]


class MLPnPBatch(C.Structure):
    _fields_ = [("C", C.c_int32), ("offsets", C.c_void_p), ("p3d", C.c_void_p), ("p2d", C.c_void_p),
                ("sigma2", C.c_void_p), ("K", C.c_void_p), ("cov", C.c_void_p), ("params", C.c_void_p),
                ("n_params", C.c_int32), ("seeds", C.c_void_p), ("tables", C.c_void_p), ("table_offsets", C.c_void_p)]


class PoseOptBatch(C.Structure):
    _fields_ = [("C", C.c_int32), ("offsets", C.c_void_p), ("p3d", C.c_void_p), ("obs", C.c_void_p),
                ("inv_sigma2", C.c_void_p), ("K", C.c_void_p), ("Tcw", C.c_void_p)]


# rsac_poseopt_result (include/ransac_b200.h)
POSEOPT_DTYPE = np.dtype([("n_inliers", np.int32), ("n_bad", np.int32), ("rounds", np.int32), ("iterations", np.int32),
                          ("trials", np.int32), ("problem", np.int32), ("R", np.float64, (9,)), ("t", np.float64, (3,)),
                          ("Rf", np.float32, (9,)), ("tf", np.float32, (3,))], align=True)


class Sim3OptBatch(C.Structure):
    _fields_ = [("C", C.c_int32), ("offsets", C.c_void_p), ("x1c", C.c_void_p), ("x2c", C.c_void_p), ("obs1", C.c_void_p),
                ("obs2", C.c_void_p), ("inv_sigma2_1", C.c_void_p), ("inv_sigma2_2", C.c_void_p), ("K1", C.c_void_p),
                ("K2", C.c_void_p), ("S12", C.c_void_p), ("th2", C.c_void_p), ("fix_scale", C.c_void_p)]


# rsac_sim3opt_result (include/ransac_b200.h)
SIM3OPT_DTYPE = np.dtype([("n_inliers", np.int32), ("n_bad", np.int32), ("optimized", np.int32), ("iterations", np.int32),
                          ("trials", np.int32), ("problem", np.int32), ("R", np.float64, (9,)), ("t", np.float64, (3,)),
                          ("s", np.float64), ("q", np.float64, (4,))], align=True)


class BowFeatures(C.Structure):
    _fields_ = [("n_feat", C.c_int32), ("desc", C.c_void_p), ("angle", C.c_void_p), ("valid", C.c_void_p),
                ("n_nodes", C.c_int32), ("node_ids", C.c_void_p), ("node_off", C.c_void_p), ("node_feat", C.c_void_p),
                ("mp_index", C.c_void_p)]


class PnPFromBow(C.Structure):
    _fields_ = [("n_matches", C.c_void_p), ("min_matches", C.c_int32), ("n_keypoints", C.c_int32), ("kp_uv", C.c_void_p),
                ("kp_sigma2", C.c_void_p), ("n_mappoints", C.c_int32), ("mp_xyz", C.c_void_p), ("K", C.c_void_p), ("params", C.c_void_p),
                ("seeds", C.c_void_p)]


class BowBatch(C.Structure):
    _fields_ = [("n_sets", C.c_int32), ("sets", C.c_void_p), ("C", C.c_int32), ("query_set", C.c_void_p), ("target_set", C.c_void_p),
                ("nn_ratio", C.c_float), ("check_orientation", C.c_int32), ("mode", C.c_int32)]


class KfDb(C.Structure):
    _fields_ = [("n_keyframes", C.c_int32), ("bow_off", C.c_void_p), ("bow_word", C.c_void_p), ("bow_val", C.c_void_p),
                ("covis", C.c_void_p), ("score_state", C.c_void_p)]


class KfDbQueries(C.Structure):
    _fields_ = [("Q", C.c_int32), ("bow_off", C.c_void_p), ("bow_word", C.c_void_p), ("bow_val", C.c_void_p), ("mode", C.c_int32),
                ("min_score", C.c_void_p), ("conn_off", C.c_void_p), ("conn", C.c_void_p)]


class KfView(C.Structure):
    _fields_ = [("n_feat", C.c_int32), ("kp_xy", C.c_void_p), ("kp_octave", C.c_void_p), ("kp_angle", C.c_void_p), ("desc", C.c_void_p), ("mp_valid", C.c_void_p),
                ("mp_xyz", C.c_void_p), ("mp_desc", C.c_void_p), ("mp_maxdist", C.c_void_p), ("mp_mindist", C.c_void_p),
                ("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("bounds", C.c_float * 4), ("grid_cols", C.c_int32), ("grid_rows", C.c_int32),
                ("grid_w_inv", C.c_float), ("grid_h_inv", C.c_float), ("grid_off", C.c_void_p), ("grid_idx", C.c_void_p),
                ("n_levels", C.c_int32), ("scale_factors", C.c_void_p), ("log_scale_factor", C.c_float)]


class Sim3SearchBatch(C.Structure):
    _fields_ = [("n_views", C.c_int32), ("views", C.c_void_p), ("C", C.c_int32), ("kf1", C.c_void_p), ("kf2", C.c_void_p), ("K", C.c_void_p),
                ("R12", C.c_void_p), ("t12", C.c_void_p), ("s12", C.c_void_p), ("th", C.c_float), ("matched12_in", C.c_void_p)]


class ProjSearchBatch(C.Structure):
    _fields_ = [("n_views", C.c_int32), ("views", C.c_void_p), ("C", C.c_int32), ("frame", C.c_void_p), ("kf", C.c_void_p), ("K", C.c_void_p),
                ("Rcw", C.c_void_p), ("tcw", C.c_void_p), ("th", C.c_float), ("orb_dist", C.c_int32), ("check_orientation", C.c_int32),
                ("occupied", C.c_void_p), ("already_found", C.c_void_p)]


class Sim3FromViews(C.Structure):
    _fields_ = [("C", C.c_int32), ("kf1", C.c_void_p), ("kf2", C.c_void_p), ("matches12", C.c_void_p), ("K1", C.c_void_p), ("K2", C.c_void_p),
                ("params", C.c_void_p), ("n_params", C.c_int32), ("seeds", C.c_void_p)]


class RsacError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"ransac_b200 error {code}: {msg}")
        self.code = code


def lib():
    """Loads the library; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run __graft_entry__.build() "
                                    "(python orb-slam2-optimized_b200/build.py); there is no CPU fallback")
        L = C.CDLL(os.environ.get("RSAC_LIB", LIB_PATH))     # RSAC_LIB: a tuning variant built by build.py (RSAC_LIB_OUT)
        L.rsac_last_error.restype = C.c_char_p
        L.rsac_launch_count.restype = C.c_int64
        for f in ("rsac_pnp_total_hypotheses", "rsac_sim3_total_hypotheses", "rsac_mlpnp_total_hypotheses",
                  "rsac_score_exact_evals"):
            if hasattr(L, f):
                getattr(L, f).restype = C.c_int64
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pnp_ransac_setup(n, prm: RansacParams):
    mi, it = C.c_int(), C.c_int()
    lib().rsac_pnp_ransac_setup(C.c_int(n), C.byref(prm), C.byref(mi), C.byref(it))
    return mi.value, it.value


def sim3_ransac_setup(n, prm: Sim3Params):
    it = C.c_int()
    lib().rsac_sim3_ransac_setup(C.c_int(n), C.byref(prm), C.byref(it))
    return it.value


def index_table(seed, n, k, H):
    out = np.empty((H, k), np.uint32)
    rc = lib().rsac_index_table(C.c_uint32(seed), C.c_int(n), C.c_int(k), C.c_int(H), _p(out))
    if rc:
        raise RsacError(rc, "rsac_index_table")
    return out


def rand_stream(seed, count):
    out = np.empty(count, np.int32)
    lib().rsac_rand_stream(C.c_uint32(seed), C.c_int(count), _p(out))
    return out


def shard_range(Cn, rank, world):
    f, c = C.c_int(), C.c_int()
    rc = lib().rsac_shard_range(C.c_int(Cn), C.c_int(rank), C.c_int(world), C.byref(f), C.byref(c))
    if rc:
        raise RsacError(rc, "rsac_shard_range")
    return f.value, c.value


def ransac_params(prob=0.99, min_inliers=8, max_its=300, min_set=4, eps=0.4, th2=5.991):
    return RansacParams(prob, min_inliers, max_its, min_set, eps, th2)


class Engine:
    """Owns one rsac_engine handle."""

    def __init__(self, device: int = 0):
        self.L = lib()
        self.h = C.c_void_p()
        rc = self.L.rsac_create(C.c_int(device), C.byref(self.h))
        if rc:
            raise RsacError(rc, "rsac_create: no usable CUDA device" if rc == ERR_NO_DEVICE else "rsac_create")
        self._keep = []

    def close(self):
        if self.h:
            self.L.rsac_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise RsacError(rc, f"{what}: {self.L.rsac_last_error(self.h).decode()}")

    # -- plumbing
    def set_stream(self, cuda_stream: int | None):
        self._ck(self.L.rsac_set_stream(self.h, C.c_void_p(cuda_stream or 0)), "set_stream")

    def sync(self):
        self._ck(self.L.rsac_sync(self.h), "sync")

    def set_problem_base(self, base: int):
        self._ck(self.L.rsac_set_problem_base(self.h, C.c_int(base)), "set_problem_base")

    def set_problem_ids(self, ids):
        """global problem index of every problem of the next batches (None / empty: base + local index)"""
        if ids is None or len(ids) == 0:
            self._ck(self.L.rsac_set_problem_ids(self.h, None, 0), "rsac_set_problem_ids")
            return
        a = np.ascontiguousarray(ids, np.int32)
        self._ck(self.L.rsac_set_problem_ids(self.h, _p(a), C.c_int(len(a))), "rsac_set_problem_ids")

    def set_graphs(self, on=True):
        self._ck(self.L.rsac_set_graphs(self.h, C.c_int(1 if on else 0)), "rsac_set_graphs")

    # ---- native NCCL exchange of the per-candidate records (include/ransac_b200.h, "multi-GPU")
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = (C.c_char * 128)()
        rc = lib().rsac_nccl_get_unique_id(buf)
        if rc:
            raise RsacError(rc, "rsac_nccl_get_unique_id")
        return bytes(buf)

    def nccl_init(self, uid: bytes, rank: int, world: int):
        buf = (C.c_char * 128).from_buffer_copy(uid)
        self._ck(self.L.rsac_nccl_init(self.h, buf, C.c_int(rank), C.c_int(world)), "rsac_nccl_init")

    def nccl_allgather_results(self, d_send: int, count_per_rank: int, d_gathered: int):
        self._ck(self.L.rsac_nccl_allgather_results(self.h, C.c_void_p(d_send), C.c_int(count_per_rank), C.c_void_p(d_gathered)),
                 "rsac_nccl_allgather_results")

    def nccl_destroy(self):
        self._ck(self.L.rsac_nccl_destroy(self.h), "rsac_nccl_destroy")

    def set_first_phase(self, hypotheses: int):
        self._ck(self.L.rsac_set_first_phase(self.h, C.c_int(hypotheses)), "set_first_phase")

    def set_phases(self, first: int, second: int):
        self._ck(self.L.rsac_set_phases(self.h, C.c_int(first), C.c_int(second)), "set_phases")

    def set_stages(self, bounds):
        b = np.ascontiguousarray(bounds, np.int32)
        self._ck(self.L.rsac_set_stages(self.h, C.c_int(len(b)), _p(b) if len(b) else None), "set_stages")

    def pnp_phase_stats(self):
        """(first_phase, problems in phase B, problems in phase C, hypotheses solved) of the last early-exit run"""
        out = (C.c_int64 * 4)()
        self._ck(self.L.rsac_pnp_phase_stats(self.h, out), "pnp_phase_stats")
        return tuple(int(v) for v in out)

    def pnp_rerun(self, resume_from, flags=0, d_results_out: int | None = None):
        r = np.ascontiguousarray(resume_from, np.int32)
        self._ck(self.L.rsac_pnp_rerun(self.h, C.c_int(flags), _p(r), C.c_void_p(d_results_out or 0)), "pnp_rerun")

    def device_info(self):
        info = DeviceInfo()
        self._ck(self.L.rsac_get_device_info(self.h, C.byref(info)), "device_info")
        return dict(sm_count=info.sm_count, sm_clock_khz=info.sm_clock_khz, mem_clock_khz=info.mem_clock_khz,
                    cc=(info.cc_major, info.cc_minor), total_mem=info.total_mem, name=info.name.decode())

    def timer_begin(self):
        self._ck(self.L.rsac_timer_begin(self.h), "timer_begin")

    def timer_end(self) -> float:
        ms = C.c_float()
        self._ck(self.L.rsac_timer_end(self.h, C.byref(ms)), "timer_end")
        return ms.value

    def profile_enable(self, on=True):
        self._ck(self.L.rsac_profile_enable(self.h, C.c_int(1 if on else 0)), "profile_enable")

    def profile_reset(self):
        self._ck(self.L.rsac_profile_reset(self.h), "profile_reset")

    def profile(self):
        out = {}
        for i, nm in enumerate(STAGE_NAMES):
            ms, n = C.c_double(), C.c_int64()
            self._ck(self.L.rsac_profile_get(self.h, C.c_int(i), C.byref(ms), C.byref(n)), "profile_get")
            out[nm] = (ms.value, n.value)
        return out

    def profile_trace(self, max_entries=4096):
        st = np.zeros(max_entries, np.int32)
        ms = np.zeros(max_entries, np.float32)
        n = self.L.rsac_profile_trace(self.h, C.c_int(max_entries), _p(st), _p(ms))
        if n < 0:
            raise RsacError(3, "profile_trace")
        return [(STAGE_NAMES[int(s)], float(m)) for s, m in zip(st[:n], ms[:n])]

    def launch_count(self) -> int:
        return self.L.rsac_launch_count(self.h)

    def measure_peaks(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.rsac_measure_peaks(self.h, C.byref(a), C.byref(b)), "measure_peaks")
        return a.value, b.value

    # -- PnP
    def _pnp_desc(self, offsets, p3d, p2d, sigma2, K, params, seeds=None, tables=None, table_offsets=None):
        offsets = np.ascontiguousarray(offsets, np.int32)
        Cn = len(offsets) - 1
        p3d = np.ascontiguousarray(p3d, np.float32).reshape(-1, 3)
        p2d = np.ascontiguousarray(p2d, np.float32).reshape(-1, 2)
        sigma2 = np.ascontiguousarray(sigma2, np.float32).reshape(-1)
        K = np.ascontiguousarray(K, np.float64).reshape(-1, 4)
        if K.shape[0] == 1 and Cn != 1:
            K = np.ascontiguousarray(np.repeat(K, max(Cn, 1), axis=0))
        if isinstance(params, RansacParams):
            parr = (RansacParams * 1)(params)
            npar = 1
        else:
            parr = (RansacParams * len(params))(*params)
            npar = len(params)
        seeds = None if seeds is None else np.ascontiguousarray(seeds, np.uint32)
        tables = None if tables is None else np.ascontiguousarray(tables, np.uint32).reshape(-1)
        table_offsets = None if table_offsets is None else np.ascontiguousarray(table_offsets, np.int64)
        desc = PnPBatch(Cn, _p(offsets), _p(p3d), _p(p2d), _p(sigma2), _p(K), C.cast(parr, C.c_void_p), npar,
                        _p(seeds), _p(tables), _p(table_offsets))
        self._keep = [offsets, p3d, p2d, sigma2, K, parr, seeds, tables, table_offsets]
        return desc, Cn, offsets

    def pnp_upload(self, offsets, p3d, p2d, sigma2, K, params, seeds=None, tables=None, table_offsets=None):
        desc, Cn, offsets = self._pnp_desc(offsets, p3d, p2d, sigma2, K, params, seeds, tables, table_offsets)
        self._ck(self.L.rsac_pnp_upload(self.h, C.byref(desc)), "pnp_upload")
        self._pnp_C = Cn
        self._pnp_total = int(offsets[-1])
        self._pnp_words = ((np.diff(offsets) + 31) // 32).astype(np.int64)
        return Cn

    def pnp_upload_indexed(self, offsets, kp_idx, mp_idx, K, params, seeds=None, kp_uv=None, kp_sigma2=None, mp_xyz=None):
        """indexed wire format: (keypoint index u16, map-point index u32) pairs over resident tables; kp_uv / mp_xyz None keeps
        the tables of the previous upload"""
        offsets = np.ascontiguousarray(offsets, np.int32)
        Cn = len(offsets) - 1
        kp_idx = np.ascontiguousarray(kp_idx, np.uint16).reshape(-1)
        mp_idx = np.ascontiguousarray(mp_idx, np.uint32).reshape(-1)
        K = np.ascontiguousarray(K, np.float64).reshape(4)
        parr = (RansacParams * 1)(params)
        seeds = None if seeds is None else np.ascontiguousarray(seeds, np.uint32)
        kp_uv = None if kp_uv is None else np.ascontiguousarray(kp_uv, np.float32).reshape(-1, 2)
        kp_sigma2 = None if kp_sigma2 is None else np.ascontiguousarray(kp_sigma2, np.float32).reshape(-1)
        mp_xyz = None if mp_xyz is None else np.ascontiguousarray(mp_xyz, np.float32).reshape(-1, 3)
        desc = PnPIndexedBatch(0 if kp_uv is None else kp_uv.shape[0], _p(kp_uv), _p(kp_sigma2), 0 if mp_xyz is None else mp_xyz.shape[0],
                               _p(mp_xyz), Cn, _p(offsets), _p(kp_idx), _p(mp_idx), _p(K), C.cast(parr, C.c_void_p), 1, _p(seeds), None, None)
        self._keep = [offsets, kp_idx, mp_idx, K, parr, seeds, kp_uv, kp_sigma2, mp_xyz]
        self._ck(self.L.rsac_pnp_upload_indexed(self.h, C.byref(desc)), "pnp_upload_indexed")
        self._pnp_C = Cn
        self._pnp_total = int(offsets[-1])
        self._pnp_words = ((np.diff(offsets) + 31) // 32).astype(np.int64)
        return Cn

    def pnp_upload_from_bow(self, n_matches, K, params, seeds, min_matches=15, kp_uv=None, kp_sigma2=None, mp_xyz=None):
        """the PnP batch built on the device from the last bow_run (mode 0, one frame): only the match COUNTS come from the host"""
        nm = np.ascontiguousarray(n_matches, np.int32)
        K = np.ascontiguousarray(K, np.float64).reshape(4)
        parr = (RansacParams * 1)(params)
        seeds = np.ascontiguousarray(seeds, np.uint32)
        kp_uv = None if kp_uv is None else np.ascontiguousarray(kp_uv, np.float32).reshape(-1, 2)
        kp_sigma2 = None if kp_sigma2 is None else np.ascontiguousarray(kp_sigma2, np.float32).reshape(-1)
        mp_xyz = None if mp_xyz is None else np.ascontiguousarray(mp_xyz, np.float32).reshape(-1, 3)
        d = PnPFromBow(_p(nm), int(min_matches), 0 if kp_uv is None else kp_uv.shape[0], _p(kp_uv), _p(kp_sigma2),
                       0 if mp_xyz is None else mp_xyz.shape[0], _p(mp_xyz), _p(K), C.cast(parr, C.c_void_p), _p(seeds))
        self._ck(self.L.rsac_pnp_upload_from_bow(self.h, C.byref(d)), "rsac_pnp_upload_from_bow")
        counts = np.where(nm >= min_matches, nm, 0).astype(np.int64)
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self._pnp_C = len(nm)
        self._pnp_total = int(offsets[-1])
        self._pnp_words = ((np.diff(offsets) + 31) // 32).astype(np.int64)
        return offsets

    def pnp_run(self, flags=0, d_results_out: int | None = None):
        self._ck(self.L.rsac_pnp_run(self.h, C.c_int(flags), C.c_void_p(d_results_out or 0)), "pnp_run")

    def pnp_download(self, want_masks=True):
        res = np.zeros(self._pnp_C, RESULT_DTYPE)
        masks = np.zeros(int(self._pnp_words.sum()), np.uint32) if want_masks else None
        self._ck(self.L.rsac_pnp_download(self.h, _p(res), _p(masks)), "pnp_download")
        return res, masks

    def pnp_download_async(self, results_ptr: int | None, masks_ptr: int | None):
        """D2H into caller-owned PINNED host memory (raw addresses), no synchronisation"""
        self._ck(self.L.rsac_pnp_download_async(self.h, C.c_void_p(results_ptr or 0), C.c_void_p(masks_ptr or 0)),
                 "pnp_download_async")

    def pnp_solve(self, offsets, p3d, p2d, sigma2, K, params, seeds=None, tables=None, table_offsets=None, flags=0):
        self.pnp_upload(offsets, p3d, p2d, sigma2, K, params, seeds, tables, table_offsets)
        self.pnp_run(flags)
        return self.pnp_download()

    def pnp_hypotheses(self):
        n = self.L.rsac_pnp_total_hypotheses(self.h)
        poses = np.zeros((n, 12), np.float32)
        counts = np.zeros(n, np.int32)
        self._ck(self.L.rsac_pnp_get_hypotheses(self.h, _p(poses), _p(counts)), "pnp_get_hypotheses")
        return poses, counts

    def split_masks(self, masks, offsets):
        """concatenated words -> list of bool arrays (compact correspondence index)"""
        out = []
        w0 = 0
        for c in range(len(offsets) - 1):
            n = int(offsets[c + 1] - offsets[c])
            nw = (n + 31) // 32
            out.append(unpack_mask(masks[w0:w0 + nw], n))
            w0 += nw
        return out

    # -- MLPnP
    def mlpnp_upload(self, offsets, p3d, p2d, sigma2, K, params, cov=None, seeds=None, tables=None, table_offsets=None):
        offsets = np.ascontiguousarray(offsets, np.int32)
        Cn = len(offsets) - 1
        p3d = np.ascontiguousarray(p3d, np.float32).reshape(-1, 3)
        p2d = np.ascontiguousarray(p2d, np.float32).reshape(-1, 2)
        sigma2 = np.ascontiguousarray(sigma2, np.float32).reshape(-1)
        K = np.ascontiguousarray(K, np.float32).reshape(-1, 4)
        if K.shape[0] == 1 and Cn != 1:
            K = np.ascontiguousarray(np.repeat(K, max(Cn, 1), axis=0))
        cov = None if cov is None else np.ascontiguousarray(cov, np.float64).reshape(-1, 9)
        if isinstance(params, RansacParams):
            parr, npar = (RansacParams * 1)(params), 1
        else:
            parr, npar = (RansacParams * len(params))(*params), len(params)
        seeds = None if seeds is None else np.ascontiguousarray(seeds, np.uint32)
        tables = None if tables is None else np.ascontiguousarray(tables, np.uint32).reshape(-1)
        table_offsets = None if table_offsets is None else np.ascontiguousarray(table_offsets, np.int64)
        desc = MLPnPBatch(Cn, _p(offsets), _p(p3d), _p(p2d), _p(sigma2), _p(K), _p(cov), C.cast(parr, C.c_void_p), npar,
                          _p(seeds), _p(tables), _p(table_offsets))
        self._keep_mlpnp = [offsets, p3d, p2d, sigma2, K, cov, parr, seeds, tables, table_offsets]
        self._ck(self.L.rsac_mlpnp_upload(self.h, C.byref(desc)), "mlpnp_upload")
        self._mlpnp_C = Cn
        self._mlpnp_words = ((np.diff(offsets) + 31) // 32).astype(np.int64)
        return Cn

    def mlpnp_run(self, flags=0, d_results_out: int | None = None):
        self._ck(self.L.rsac_mlpnp_run(self.h, C.c_int(flags), C.c_void_p(d_results_out or 0)), "mlpnp_run")

    def mlpnp_phase_stats(self):
        out = (C.c_int64 * 4)()
        self._ck(self.L.rsac_mlpnp_phase_stats(self.h, out), "mlpnp_phase_stats")
        return tuple(int(v) for v in out)

    def mlpnp_rerun(self, resume_from, flags=0, d_results_out: int | None = None):
        r = np.ascontiguousarray(resume_from, np.int32)
        self._ck(self.L.rsac_mlpnp_rerun(self.h, C.c_int(flags), _p(r), C.c_void_p(d_results_out or 0)), "mlpnp_rerun")

    def mlpnp_download(self, want_masks=True):
        res = np.zeros(self._mlpnp_C, RESULT_DTYPE)
        masks = np.zeros(int(self._mlpnp_words.sum()), np.uint32) if want_masks else None
        self._ck(self.L.rsac_mlpnp_download(self.h, _p(res), _p(masks)), "mlpnp_download")
        return res, masks

    def mlpnp_solve(self, offsets, p3d, p2d, sigma2, K, params, cov=None, seeds=None, tables=None, table_offsets=None, flags=0):
        self.mlpnp_upload(offsets, p3d, p2d, sigma2, K, params, cov, seeds, tables, table_offsets)
        self.mlpnp_run(flags)
        return self.mlpnp_download()

    def mlpnp_hypotheses(self):
        n = self.L.rsac_mlpnp_total_hypotheses(self.h)
        poses = np.zeros((n, 12), np.float64)
        counts = np.zeros(n, np.int32)
        self._ck(self.L.rsac_mlpnp_get_hypotheses(self.h, _p(poses), _p(counts)), "mlpnp_get_hypotheses")
        return poses, counts

    # -- Sim3
    def sim3_upload(self, offsets, x1c, x2c, s1, s2, K1, K2, params, seeds=None, tables=None, table_offsets=None):
        offsets = np.ascontiguousarray(offsets, np.int32)
        Cn = len(offsets) - 1
        x1c = np.ascontiguousarray(x1c, np.float32).reshape(-1, 3)
        x2c = np.ascontiguousarray(x2c, np.float32).reshape(-1, 3)
        s1 = np.ascontiguousarray(s1, np.float32).reshape(-1)
        s2 = np.ascontiguousarray(s2, np.float32).reshape(-1)
        K1 = np.ascontiguousarray(K1, np.float32).reshape(-1, 4)
        K2 = np.ascontiguousarray(K2, np.float32).reshape(-1, 4)
        if K1.shape[0] == 1 and Cn != 1:
            K1 = np.ascontiguousarray(np.repeat(K1, max(Cn, 1), axis=0))
        if K2.shape[0] == 1 and Cn != 1:
            K2 = np.ascontiguousarray(np.repeat(K2, max(Cn, 1), axis=0))
        if isinstance(params, Sim3Params):
            parr, npar = (Sim3Params * 1)(params), 1
        else:
            parr, npar = (Sim3Params * len(params))(*params), len(params)
        seeds = None if seeds is None else np.ascontiguousarray(seeds, np.uint32)
        tables = None if tables is None else np.ascontiguousarray(tables, np.uint32).reshape(-1)
        table_offsets = None if table_offsets is None else np.ascontiguousarray(table_offsets, np.int64)
        desc = Sim3Batch(Cn, _p(offsets), _p(x1c), _p(x2c), _p(s1), _p(s2), _p(K1), _p(K2), C.cast(parr, C.c_void_p), npar,
                         _p(seeds), _p(tables), _p(table_offsets))
        self._keep_sim3 = [offsets, x1c, x2c, s1, s2, K1, K2, parr, seeds, tables, table_offsets]
        self._ck(self.L.rsac_sim3_upload(self.h, C.byref(desc)), "sim3_upload")
        self._sim3_C = Cn
        self._sim3_words = ((np.diff(offsets) + 31) // 32).astype(np.int64)
        return Cn

    def sim3_run(self, flags=0, d_results_out: int | None = None):
        self._ck(self.L.rsac_sim3_run(self.h, C.c_int(flags), C.c_void_p(d_results_out or 0)), "sim3_run")

    def sim3_download(self, want_masks=True):
        res = np.zeros(self._sim3_C, RESULT_DTYPE)
        masks = np.zeros(int(self._sim3_words.sum()), np.uint32) if want_masks else None
        self._ck(self.L.rsac_sim3_download(self.h, _p(res), _p(masks)), "sim3_download")
        return res, masks

    def sim3_solve(self, offsets, x1c, x2c, s1, s2, K1, K2, params, seeds=None, tables=None, table_offsets=None, flags=0):
        self.sim3_upload(offsets, x1c, x2c, s1, s2, K1, K2, params, seeds, tables, table_offsets)
        self.sim3_run(flags)
        return self.sim3_download()

    def sim3_hypotheses(self, hyp_words: int):
        """poses [sumH,13], counts [sumH], masks (flat words; hyp_words = total per-hypothesis words)"""
        n = self.L.rsac_sim3_total_hypotheses(self.h)
        poses = np.zeros((n, 13), np.float32)
        counts = np.zeros(n, np.int32)
        masks = np.zeros(max(hyp_words, 1), np.uint32)
        self._ck(self.L.rsac_sim3_get_hypotheses(self.h, _p(poses), _p(counts), _p(masks)), "sim3_get_hypotheses")
        return poses, counts, masks

    # -- Optimizer::PoseOptimization (batched)
    def poseopt_upload(self, offsets, p3d, obs, inv_sigma2, K, Tcw):
        offsets = np.ascontiguousarray(offsets, np.int32)
        Cn = len(offsets) - 1
        p3d = np.ascontiguousarray(p3d, np.float32).reshape(-1, 3)
        obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 3)
        inv_sigma2 = np.ascontiguousarray(inv_sigma2, np.float32).reshape(-1)
        K = np.ascontiguousarray(K, np.float32).reshape(-1, 5)
        if K.shape[0] == 1 and Cn != 1:
            K = np.ascontiguousarray(np.repeat(K, max(Cn, 1), axis=0))
        Tcw = np.ascontiguousarray(Tcw, np.float32).reshape(-1, 12)
        desc = PoseOptBatch(Cn, _p(offsets), _p(p3d), _p(obs), _p(inv_sigma2), _p(K), _p(Tcw))
        self._keep_poseopt = [offsets, p3d, obs, inv_sigma2, K, Tcw]
        self._ck(self.L.rsac_poseopt_upload(self.h, C.byref(desc)), "poseopt_upload")
        self._poseopt_C, self._poseopt_total = Cn, int(offsets[-1])
        return Cn

    def poseopt_from_pnp(self, bf=0.0):
        """chain PoseOptimization behind the engine's last PnP run on the device (flags come back per PnP correspondence)"""
        self._ck(self.L.rsac_poseopt_from_pnp(self.h, C.c_float(bf)), "poseopt_from_pnp")
        self._poseopt_C, self._poseopt_total = self._pnp_C, self._pnp_total

    def poseopt_run(self):
        self._ck(self.L.rsac_poseopt_run(self.h), "poseopt_run")

    def poseopt_download(self):
        res = np.zeros(self._poseopt_C, POSEOPT_DTYPE)
        outlier = np.zeros(max(self._poseopt_total, 1), np.uint8)
        self._ck(self.L.rsac_poseopt_download(self.h, _p(res), _p(outlier)), "poseopt_download")
        return res, outlier[:self._poseopt_total]

    def poseopt_solve(self, offsets, p3d, obs, inv_sigma2, K, Tcw):
        self.poseopt_upload(offsets, p3d, obs, inv_sigma2, K, Tcw)
        self.poseopt_run()
        return self.poseopt_download()

    # -- Optimizer::OptimizeSim3 (batched)
    def sim3opt_solve(self, offsets, x1c, x2c, obs1, obs2, is1, is2, K1, K2, S12, th2, fix_scale=None):
        offsets = np.ascontiguousarray(offsets, np.int32)
        Cn = len(offsets) - 1
        f = lambda a, w: np.ascontiguousarray(a, np.float32).reshape(-1, w)
        x1c, x2c, obs1, obs2 = f(x1c, 3), f(x2c, 3), f(obs1, 2), f(obs2, 2)
        is1, is2 = np.ascontiguousarray(is1, np.float32).reshape(-1), np.ascontiguousarray(is2, np.float32).reshape(-1)
        K1, K2, S12 = f(K1, 4), f(K2, 4), f(S12, 13)
        th2 = np.ascontiguousarray(np.broadcast_to(np.asarray(th2, np.float32), (max(Cn, 1),)))
        fs = None if fix_scale is None else np.ascontiguousarray(fix_scale, np.int32)
        desc = Sim3OptBatch(Cn, _p(offsets), _p(x1c), _p(x2c), _p(obs1), _p(obs2), _p(is1), _p(is2), _p(K1), _p(K2), _p(S12), _p(th2), _p(fs))
        self._ck(self.L.rsac_sim3opt_upload(self.h, C.byref(desc)), "sim3opt_upload")
        self._ck(self.L.rsac_sim3opt_run(self.h), "sim3opt_run")
        res = np.zeros(Cn, SIM3OPT_DTYPE)
        removed = np.zeros(max(int(offsets[-1]), 1), np.uint8)
        self._ck(self.L.rsac_sim3opt_download(self.h, _p(res), _p(removed)), "sim3opt_download")
        return res, removed[:int(offsets[-1])]

    def sim3opt_run(self):
        self._ck(self.L.rsac_sim3opt_run(self.h), "sim3opt_run")

    def sim3opt_from_search(self, th2=10.0, fix_scale=True, K2=None):
        """OptimizeSim3 on the device behind the last sim3_search_run (LoopClosing.cpp:309-311)"""
        k2 = None if K2 is None else np.ascontiguousarray(K2, np.float32).reshape(-1, 4)
        self._ck(self.L.rsac_sim3opt_from_search(self.h, C.c_float(th2), C.c_int(1 if fix_scale else 0), _p(k2)), "rsac_sim3opt_from_search")

    def sim3opt_download_chained(self):
        """(results [C], list of per-pair KF1-indexed flag arrays (0 kept, 1 removed, 2 no match), n_edges [C])"""
        n1 = self._s3s_n1
        Cn = len(n1)
        res = np.zeros(max(Cn, 1), SIM3OPT_DTYPE)
        flags = np.zeros(max(int(sum(n1)), 1), np.uint8)
        ne = np.zeros(max(Cn, 1), np.int32)
        self._ck(self.L.rsac_sim3opt_download_chained(self.h, _p(res), _p(flags), _p(ne)), "rsac_sim3opt_download_chained")
        offs = np.concatenate([[0], np.cumsum(n1)]).astype(np.int64)
        return res[:Cn], [flags[offs[i]:offs[i + 1]].copy() for i in range(Cn)], ne[:Cn]

    # -- scoring stress
    def score_pnp_upload(self, poses, p3d, p2d, max_err, K):
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 12)
        p3d = np.ascontiguousarray(p3d, np.float32).reshape(-1, 3)
        p2d = np.ascontiguousarray(p2d, np.float32).reshape(-1, 2)
        max_err = np.ascontiguousarray(max_err, np.float32).reshape(-1)
        Kd = (C.c_double * 4)(*[float(k) for k in K])
        self._score_H, self._score_n = poses.shape[0], p3d.shape[0]
        self._ck(self.L.rsac_score_pnp_upload(self.h, C.c_int(poses.shape[0]), _p(poses), C.c_int(p3d.shape[0]), _p(p3d),
                                              _p(p2d), _p(max_err), Kd), "score_pnp_upload")

    def score_pnp_run(self, want_masks=True):
        self._ck(self.L.rsac_score_pnp_run(self.h, C.c_int(1 if want_masks else 0)), "score_pnp_run")

    def score_pnp_download(self, want_masks=True):
        words = (self._score_n + 31) // 32
        counts = np.zeros(self._score_H, np.int32)
        masks = np.zeros((self._score_H, words), np.uint32) if want_masks else None
        self._ck(self.L.rsac_score_pnp_download(self.h, _p(masks), _p(counts)), "score_pnp_download")
        return counts, masks

    def score_pnp(self, poses, p3d, p2d, max_err, K, want_masks=True):
        self.score_pnp_upload(poses, p3d, p2d, max_err, K)
        self.score_pnp_run(want_masks)
        return self.score_pnp_download(want_masks)

    # ---- ORBmatcher::SearchByBoW
    def _bow_desc(self, sets, query_set, target_set, nn_ratio, check_orientation, mode):
        keep, arr = [], (BowFeatures * max(len(sets), 1))()
        for i, f in enumerate(sets):
            desc = np.ascontiguousarray(f["desc"], np.uint32).reshape(-1, 8)
            ang = np.ascontiguousarray(f["angle"], np.float32)
            val = None if f.get("valid") is None else np.ascontiguousarray(f["valid"], np.uint8)
            nid = np.ascontiguousarray(f["node_ids"], np.uint32)
            noff = np.ascontiguousarray(f["node_off"], np.int32)
            nfe = np.ascontiguousarray(f["node_feat"], np.uint32)
            mpi = None if f.get("mp_index") is None else np.ascontiguousarray(f["mp_index"], np.uint32)
            keep += [desc, ang, val, nid, noff, nfe, mpi]
            arr[i] = BowFeatures(desc.shape[0], _p(desc), _p(ang), _p(val), len(nid), _p(nid), _p(noff), _p(nfe), _p(mpi))
        qs, ts = np.ascontiguousarray(query_set, np.int32), np.ascontiguousarray(target_set, np.int32)
        keep += [qs, ts, arr]
        b = BowBatch(len(sets), C.cast(arr, C.c_void_p), len(qs), _p(qs), _p(ts), C.c_float(nn_ratio), int(check_orientation), int(mode))
        n_out = [sets[t if mode == 0 else q]["desc"].shape[0] for q, t in zip(qs, ts)]
        return b, keep, n_out

    def bow_upload(self, sets, query_set, target_set, nn_ratio=0.75, check_orientation=True, mode=0):
        b, keep, n_out = self._bow_desc(sets, query_set, target_set, nn_ratio, check_orientation, mode)
        self._ck(self.L.rsac_bow_upload(self.h, C.byref(b)), "rsac_bow_upload")
        self._bow_n_out = n_out

    def bow_run(self):
        self._ck(self.L.rsac_bow_run(self.h), "rsac_bow_run")

    def bow_download(self):
        """(list of per-pair match arrays, n_matches [C])"""
        n_out = self._bow_n_out
        flat = np.empty(max(int(sum(n_out)), 1), np.int32)
        nm = np.empty(max(len(n_out), 1), np.int32)
        self._ck(self.L.rsac_bow_download(self.h, _p(flat), _p(nm)), "rsac_bow_download")
        offs = np.concatenate([[0], np.cumsum(n_out)]).astype(np.int64)
        return [flat[offs[i]:offs[i + 1]] for i in range(len(n_out))], nm[:len(n_out)]

    def bow_match(self, sets, query_set, target_set, nn_ratio=0.75, check_orientation=True, mode=0):
        self.bow_upload(sets, query_set, target_set, nn_ratio, check_orientation, mode)
        self.bow_run()
        return self.bow_download()

    # -- keyframe database: candidate retrieval
    def kfdb_upload(self, db, score_state=None):
        """db: dict(bow_off int64 [K+1], bow_word uint32, bow_val float64, covis int32 [K,10])"""
        off = np.ascontiguousarray(db["bow_off"], np.int64)
        w = np.ascontiguousarray(db["bow_word"], np.uint32)
        v = np.ascontiguousarray(db["bow_val"], np.float64)
        cv = np.ascontiguousarray(db["covis"], np.int32).reshape(-1, 10)
        st = None if score_state is None else np.ascontiguousarray(score_state, np.float32)
        d = KfDb(len(off) - 1, _p(off), _p(w), _p(v), _p(cv), _p(st))
        self._ck(self.L.rsac_kfdb_upload(self.h, C.byref(d)), "rsac_kfdb_upload")
        self._kfdb_K = len(off) - 1

    def kfdb_query_upload(self, queries, mode=0, min_score=None, conn=None):
        """queries: list of (words uint32 ascending, values float64); conn: list of connected-keyframe index lists (mode 1)"""
        Q = len(queries)
        off = np.concatenate([[0], np.cumsum([len(q[0]) for q in queries])]).astype(np.int64)
        w = np.ascontiguousarray(np.concatenate([q[0] for q in queries]) if Q else [], np.uint32)
        v = np.ascontiguousarray(np.concatenate([q[1] for q in queries]) if Q else [], np.float64)
        ms = co = cn = None
        if mode == 1:
            ms = np.ascontiguousarray(min_score, np.float32).reshape(-1)
            co = np.concatenate([[0], np.cumsum([len(c) for c in conn])]).astype(np.int64)
            cn = np.ascontiguousarray(np.concatenate([np.asarray(c, np.int32) for c in conn]) if Q and co[-1] else [], np.int32)
        d = KfDbQueries(Q, _p(off), _p(w) if len(w) else None, _p(v) if len(v) else None, mode, _p(ms), _p(co), _p(cn) if cn is not None and len(cn) else None)
        self._ck(self.L.rsac_kfdb_query_upload(self.h, C.byref(d)), "rsac_kfdb_query_upload")
        self._kfdb_Q = Q

    def kfdb_run(self):
        self._ck(self.L.rsac_kfdb_run(self.h), "rsac_kfdb_run")

    def kfdb_download(self, cap=None):
        """list of candidate index arrays, one per query, in the reference's order"""
        Q = self._kfdb_Q
        cap = max(1, self._kfdb_K if cap is None else cap)
        counts = np.zeros(max(Q, 1), np.int32)
        cand = np.full((max(Q, 1), cap), -1, np.int32)
        self._ck(self.L.rsac_kfdb_download(self.h, _p(counts), _p(cand), C.c_int32(cap)), "rsac_kfdb_download")
        return [cand[q, :min(int(counts[q]), cap)].copy() for q in range(Q)], counts[:Q]

    def kfdb_detect(self, queries, mode=0, min_score=None, conn=None):
        self.kfdb_query_upload(queries, mode, min_score, conn)
        self.kfdb_run()
        return self.kfdb_download()[0]

    def kfdb_state(self):
        st = np.zeros(max(self._kfdb_K, 1), np.float32)
        self._ck(self.L.rsac_kfdb_get_state(self.h, _p(st)), "rsac_kfdb_get_state")
        return st[:self._kfdb_K]

    # -- resident keyframe views and the Sim3Solver constructor over them
    def views_upload(self, views):
        arr, keep = self._kf_views(views)
        self._ck(self.L.rsac_views_upload(self.h, C.c_int(len(views)), C.cast(arr, C.c_void_p)), "rsac_views_upload")
        self._views_n = [int(v["n_feat"]) for v in views]

    def sim3_upload_from_views(self, kf1, kf2, matches12, K1, K2, params, seeds):
        """matches12: list of per-pair int32 arrays (KF2 feature per KF1 feature, -1 none).  Returns (offsets, idx1, idx2)."""
        k1, k2 = np.ascontiguousarray(kf1, np.int32), np.ascontiguousarray(kf2, np.int32)
        Cn = len(k1)
        m = np.ascontiguousarray(np.concatenate([np.asarray(x, np.int32) for x in matches12]) if Cn else [], np.int32)
        K1 = np.ascontiguousarray(K1, np.float32).reshape(-1, 4); K2 = np.ascontiguousarray(K2, np.float32).reshape(-1, 4)
        parr = (Sim3Params * 1)(params)
        seeds = np.ascontiguousarray(seeds, np.uint32)
        offs = np.zeros(Cn + 1, np.int32)
        idx1 = np.zeros(max(len(m), 1), np.int32); idx2 = np.zeros(max(len(m), 1), np.int32)
        d = Sim3FromViews(Cn, _p(k1), _p(k2), _p(m) if len(m) else None, _p(K1), _p(K2), C.cast(parr, C.c_void_p), 1, _p(seeds))
        self._ck(self.L.rsac_sim3_upload_from_views(self.h, C.byref(d), _p(offs), _p(idx1), _p(idx2)), "rsac_sim3_upload_from_views")
        n = int(offs[-1])
        self._sim3_C = Cn
        self._sim3_total = n
        self._sim3_words = ((np.diff(offs) + 31) // 32).astype(np.int64)
        return offs, idx1[:n].copy(), idx2[:n].copy()

    # -- guided matching: ORBmatcher::SearchBySim3
    def sim3_search_upload(self, views, kf1, kf2, K, R12, t12, th=7.5, matched12_in=None, s12=None):
        """views: list of keyframe-view dicts (synth.kf_view); kf1 / kf2: view index per pair; K [C,4], R12 [C,9], t12 [C,3];
        matched12_in: list of per-pair int32 arrays (or None)"""
        resident = views is None
        arr, keep = (None, []) if resident else self._kf_views(views)
        k1, k2 = np.ascontiguousarray(kf1, np.int32), np.ascontiguousarray(kf2, np.int32)
        Kc = np.ascontiguousarray(K, np.float32).reshape(-1, 4)
        R = np.ascontiguousarray(R12, np.float32).reshape(-1, 9)
        t = np.ascontiguousarray(t12, np.float32).reshape(-1, 3)
        sc = None if s12 is None else np.ascontiguousarray(s12, np.float32)
        mi = None if matched12_in is None else np.ascontiguousarray(np.concatenate([np.asarray(m, np.int32) for m in matched12_in]) if len(k1) else [], np.int32)
        b = Sim3SearchBatch(0 if resident else len(views), None if resident else C.cast(arr, C.c_void_p), len(k1), _p(k1), _p(k2), _p(Kc), _p(R), _p(t),
                            _p(sc), C.c_float(th), _p(mi) if mi is not None and len(mi) else None)
        self._ck(self.L.rsac_sim3_search_upload(self.h, C.byref(b)), "rsac_sim3_search_upload")
        self._s3s_n1 = [self._views_n[i] for i in k1] if resident else [int(views[i]["n_feat"]) for i in k1]

    @staticmethod
    def _kf_views(views):
        keep = []
        arr = (KfView * max(len(views), 1))()
        for i, v in enumerate(views):
            ang = np.ascontiguousarray(v["kp_angle"], np.float32)
            keep.append(ang)
            a = [np.ascontiguousarray(v["kp_xy"], np.float32), np.ascontiguousarray(v["kp_octave"], np.int32), np.ascontiguousarray(v["desc"], np.uint32),
                 np.ascontiguousarray(v["mp_valid"], np.uint8), np.ascontiguousarray(v["mp_xyz"], np.float32), np.ascontiguousarray(v["mp_desc"], np.uint32),
                 np.ascontiguousarray(v["mp_maxdist"], np.float32), np.ascontiguousarray(v["mp_mindist"], np.float32),
                 np.ascontiguousarray(v["grid_off"], np.int32), np.ascontiguousarray(v["grid_idx"], np.int32), np.ascontiguousarray(v["scale_factors"], np.float32)]
            keep += a
            arr[i] = KfView(int(v["n_feat"]), _p(a[0]), _p(a[1]), _p(ang), _p(a[2]), _p(a[3]), _p(a[4]), _p(a[5]), _p(a[6]), _p(a[7]),
                            (C.c_float * 9)(*np.asarray(v["Rcw"], np.float32).reshape(-1)), (C.c_float * 3)(*np.asarray(v["tcw"], np.float32).reshape(-1)),
                            (C.c_float * 4)(*np.asarray(v["bounds"], np.float32)), int(v["grid_cols"]), int(v["grid_rows"]), float(v["grid_w_inv"]),
                            float(v["grid_h_inv"]), _p(a[8]), _p(a[9]), int(v["n_levels"]), _p(a[10]), float(v["log_scale_factor"]))
        return arr, keep

    # -- guided matching: ORBmatcher::SearchByProjection(Frame, KeyFrame, sAlreadyFound, th, ORBdist)
    def proj_search_upload(self, views, frame, kf, K, Rcw, tcw, th=10.0, orb_dist=100, check_orientation=True, occupied=None, already_found=None):
        arr, keep = self._kf_views(views)
        fi, ki = np.ascontiguousarray(frame, np.int32), np.ascontiguousarray(kf, np.int32)
        Kc = np.ascontiguousarray(K, np.float32).reshape(-1, 4)
        R = np.ascontiguousarray(Rcw, np.float32).reshape(-1, 9)
        t = np.ascontiguousarray(tcw, np.float32).reshape(-1, 3)
        oc = None if occupied is None else np.ascontiguousarray(np.concatenate([np.asarray(o, np.uint8) for o in occupied]) if len(fi) else [], np.uint8)
        af = None if already_found is None else np.ascontiguousarray(np.concatenate([np.asarray(o, np.uint8) for o in already_found]) if len(fi) else [], np.uint8)
        b = ProjSearchBatch(len(views), C.cast(arr, C.c_void_p), len(fi), _p(fi), _p(ki), _p(Kc), _p(R), _p(t), C.c_float(th), int(orb_dist),
                            int(check_orientation), _p(oc) if oc is not None and len(oc) else None, _p(af) if af is not None and len(af) else None)
        self._ck(self.L.rsac_proj_search_upload(self.h, C.byref(b)), "rsac_proj_search_upload")
        self._sbp_nf = [int(views[i]["n_feat"]) for i in fi]

    def proj_search_run(self):
        self._ck(self.L.rsac_proj_search_run(self.h), "rsac_proj_search_run")

    def proj_search_download(self):
        """(list of per-pair frame_match arrays, n_matches [C], fell_back [C], rounds [C])"""
        nf = self._sbp_nf
        Cn = len(nf)
        flat = np.empty(max(int(sum(nf)), 1), np.int32)
        nm = np.zeros(max(Cn, 1), np.int32)
        info = np.zeros((2, max(Cn, 1)), np.int32)
        self._ck(self.L.rsac_proj_search_download(self.h, _p(flat), _p(nm), _p(info)), "rsac_proj_search_download")
        offs = np.concatenate([[0], np.cumsum(nf)]).astype(np.int64)
        return [flat[offs[i]:offs[i + 1]].copy() for i in range(Cn)], nm[:Cn], info[0, :Cn], info[1, :Cn]

    def proj_search(self, views, frame, kf, K, Rcw, tcw, th=10.0, orb_dist=100, check_orientation=True, occupied=None, already_found=None):
        self.proj_search_upload(views, frame, kf, K, Rcw, tcw, th, orb_dist, check_orientation, occupied, already_found)
        self.proj_search_run()
        return self.proj_search_download()

    def sim3_search_run(self):
        self._ck(self.L.rsac_sim3_search_run(self.h), "rsac_sim3_search_run")

    def sim3_search_download(self):
        """(list of per-pair match12 arrays, n_found [C])"""
        n1 = self._s3s_n1
        flat = np.empty(max(int(sum(n1)), 1), np.int32)
        nf = np.zeros(max(len(n1), 1), np.int32)
        self._ck(self.L.rsac_sim3_search_download(self.h, _p(flat), _p(nf)), "rsac_sim3_search_download")
        offs = np.concatenate([[0], np.cumsum(n1)]).astype(np.int64)
        return [flat[offs[i]:offs[i + 1]].copy() for i in range(len(n1))], nf[:len(n1)]

    def sim3_search(self, views, kf1, kf2, K, R12, t12, th=7.5, matched12_in=None, s12=None):
        self.sim3_search_upload(views, kf1, kf2, K, R12, t12, th, matched12_in, s12)
        self.sim3_search_run()
        return self.sim3_search_download()

    def score_exact_evals(self) -> int:
        return self.L.rsac_score_exact_evals(self.h)


def debug_host_poseopt(p3d, obs, inv_sigma2, K, Tcw):
    """the device source of PoseOptimization compiled for the host, one frame (test hook, not a fallback)"""
    p3d = np.ascontiguousarray(p3d, np.float32).reshape(-1, 3)
    obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 3)
    inv_sigma2 = np.ascontiguousarray(inv_sigma2, np.float32).reshape(-1)
    K = np.ascontiguousarray(K, np.float32).reshape(5)
    Tcw = np.ascontiguousarray(Tcw, np.float32).reshape(12)
    res = np.zeros(1, POSEOPT_DTYPE)
    outlier = np.zeros(max(p3d.shape[0], 1), np.uint8)
    rc = lib().rsac_debug_host_poseopt(C.c_int(p3d.shape[0]), _p(p3d), _p(obs), _p(inv_sigma2), _p(K), _p(Tcw), _p(res), _p(outlier))
    if rc:
        raise RsacError(rc, "debug_host_poseopt")
    return res[0], outlier[:p3d.shape[0]]


def debug_host_sim3opt(x1c, x2c, obs1, obs2, is1, is2, K1, K2, S12, th2=10.0, fix_scale=True):
    """the device source of OptimizeSim3 compiled for the host, one keyframe pair (test hook, not a fallback)"""
    f = lambda a, w: np.ascontiguousarray(a, np.float32).reshape(-1, w)
    x1c, x2c, obs1, obs2 = f(x1c, 3), f(x2c, 3), f(obs1, 2), f(obs2, 2)
    is1, is2 = np.ascontiguousarray(is1, np.float32).reshape(-1), np.ascontiguousarray(is2, np.float32).reshape(-1)
    K1, K2, S12 = f(K1, 4).ravel(), f(K2, 4).ravel(), f(S12, 13).ravel()
    res = np.zeros(1, SIM3OPT_DTYPE)
    removed = np.zeros(max(x1c.shape[0], 1), np.uint8)
    rc = lib().rsac_debug_host_sim3opt(C.c_int(x1c.shape[0]), _p(x1c), _p(x2c), _p(obs1), _p(obs2), _p(is1), _p(is2), _p(K1), _p(K2),
                                       _p(S12), C.c_float(th2), C.c_int(1 if fix_scale else 0), _p(res), _p(removed))
    if rc:
        raise RsacError(rc, "debug_host_sim3opt")
    return res[0], removed[:x1c.shape[0]]


def unpack_mask(words: np.ndarray, n: int) -> np.ndarray:
    """uint32 words (bit i of word w = item 32w+i) -> bool[n]"""
    words = np.ascontiguousarray(words, np.uint32)
    b = ((words[..., :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)
    return b.reshape(words.shape[:-1] + (-1,))[..., :n]
