#!/usr/bin/env python
"""bench.py -- relocalisation-sweep throughput of the B200 RANSAC engine (BASELINE.json metric).

Workload (config.workload = "cfg4"): relocalisation sweeps of 1024 candidate keyframes x 500 2D-3D matches, 50 %
outliers, PnPsolver EPnP RANSAC with SetRansacParameters(0.99,10,300,4,0.2,5.991) => H = 300 hypotheses per candidate.
The reference stops at the first hypothesis whose Refine() succeeds; the device does the same in stages
(RSAC_FLAG_EARLY_EXIT), then replays the reference's sequential semantics (PnPsolver::iterate / Refine) per candidate
-- records and masks are identical to solving and scoring all 300 (RSAC_BENCH_EXHAUSTIVE=1 times that mode).

A STEP is a fixed bundle of 64 independent sweeps = 65 536 candidates, whatever the number of GPUs (STRONG scaling:
BASELINE cfg4 "1024 candidates sharded at 1/2/4/8 GPUs with NCCL best-pose gather").  With N GPUs every sweep's
candidates are sharded in contiguous blocks of 1024/N; a rank concatenates its shards of N consecutive
sweeps into one device batch of 1024 candidates, and after every batch the per-candidate records (96 B) of the N sweeps
are all-gathered over NCCL, so that every rank holds the N complete sweeps.  A step is 64/N batches per GPU.
  value : candidates/s, inputs resident in HBM
  e2e   : the same through the host-buffer C-ABI call sequence (H2D of every batch's inputs from pinned memory in the
          indexed wire format -- keypoint tables + (keypoint, map point) index pairs over the device-resident map --,
          D2H of records + inlier masks inside the timed region)
  roofline      : dominant kernel of the sweep (EPnP minimal solver, FP64 CUDA cores)
  roofline_score: CheckInliers kernel on cfg5 (4096 poses x 10 000 correspondences, FP32 CUDA cores)
  cpu_baseline  : the CPU oracle (port of the reference, see oracle/) on the host cores
After the timed passes rank 0 checks CONTENT: the end-to-end records and masks equal the resident run's, and (N > 1) the
records gathered from N ranks -- through torch.distributed and through the library's own rsac_nccl_allgather_results --
are byte-identical to a one-rank run of the same sweeps.

--impl reference runs the reference arm: the oracle port of PnPsolver on all host threads, same workload, same metric,
each step a bounded sample of the bundle; beside it (`compiled_reference`) the reference's own PnPsolver.cpp, compiled unmodified
against stand-in Eigen / OpenCV headers (oracle/_ref), one process per core on one sweep.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

C_SWEEP = 1024                # candidates of one relocalisation sweep (BASELINE cfg4)
N_MATCH = 500
PRM = dict(prob=0.99, min_inliers=10, max_its=300, min_set=4, eps=0.2, th2=5.991)
H_HYP = 300
NBLOCK = 8                    # distinct synthetic sweeps ("blocks") that are cycled
BUNDLE = int(os.environ.get("RSAC_BENCH_BUNDLE", "64"))   # sweeps per step (a multiple of 8): 65 536 candidates per step
METRIC = "relocalization candidates/s (PnP EPnP RANSAC sweep; hyp x corr evals/s in extras)"
FLOP_PER_EVAL = 31            # PnP CheckInliers
PIPE = int(os.environ.get("RSAC_BENCH_PIPE", "0"))   # device batches in flight per GPU (0: default, six)
EXHAUSTIVE = os.environ.get("RSAC_BENCH_EXHAUSTIVE", "0") == "1"   # solve and score all 300 hypotheses of every candidate
EIGEN = os.environ.get("RSAC_BENCH_EIGEN", "0") == "1"             # 12x12 eigen-solve null space (the reference's structure)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


N_KP = 2000                   # keypoints of the frame being relocalised
N_MAP = 200000                # map points of one synthetic map (one map per block; the blocks' maps are concatenated)


def make_block(j):
    """synthetic sweep j = ONE relocalisation as Tracking::Relocalization sees it (Tracking.cpp:1196-1232): one frame of 2000
    keypoints, 1024 candidate keyframes, each contributing 500 (keypoint, map point) pairs, half of them wrong map points.
    Carries the indexed form (tables + index pairs) and the equivalent flat arrays p3d [C,n,3], p2d [C,n,2], sigma2 [C,n]."""
    from ransac_b200 import synth
    return synth.reloc_frame(4000 + j, C_SWEEP, n_kp=N_KP, n_match=N_MATCH, outlier_ratio=0.5, n_map=N_MAP)


def bind_to_gpu_numa(local_rank):
    """Pin this rank's threads (and so its first-touch pinned host buffers) to the NUMA node of its GPU: with one process
    per GPU and unbound ranks, all eight ranks' input buffers were served by one socket (round 1: 145 GB/s aggregate, e2e
    scaling efficiency 0.72).  Returns the node, or None when the topology is not exposed (containers often hide it)."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        dom, rest = bus.split(":", 1)
        node = int(open(f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def shard_of(C, rank, world):
    per = (C + world - 1) // world
    first = min(C, rank * per)
    return first, min(C, first + per) - first


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU implementation of the path (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle_api as O
    O.build()
    cores = os.cpu_count() or 1
    prm = O.params(**PRM)
    oflags = 0 if EIGEN else O.FLAG_EPNP_QR_NULLSPACE   # the port's faster mode (same arithmetic as the device path)
    # bounded sample of a step: SAMPLE of the bundle's 64 sweeps (8 distinct blocks are cycled, like the GPU arm)
    sample = max(1, int(os.environ.get("RSAC_BENCH_REF_SWEEPS", "4")))
    blocks, blocks_raw = [], []
    for j in range(min(NBLOCK, sample)):
        bj = make_block(j)
        blocks_raw.append(bj)
        blocks.append(([O.pnp_problem(bj["p3d"][c], bj["p2d"][c], bj["sigma2"][c], bj["K"]) for c in range(C_SWEEP)],
                       [O.index_table(int(sd), N_MATCH, 4, H_HYP) for sd in bj["seeds"]]))
    for _ in range(max(1, min(args.warmup, 1))):
        O.pnp_batch(blocks[0][0][:64], prm, blocks[0][1][:64], oflags, cores)
    t_tot, ev_tot, n_tot = 0.0, 0, 0
    for k in range(args.steps):
        for q in range(sample):
            pbs, tables = blocks[(k * sample + q) % len(blocks)]
            dt, ev, res = O.pnp_batch(pbs, prm, tables, oflags, cores)
            t_tot += dt
            ev_tot += ev
            n_tot += len(pbs)
    val = n_tot / t_tot
    compiled = None
    if os.environ.get("RSAC_BENCH_COMPILED_REF", "1") == "1":
        try:
            compiled = compiled_reference_leg(blocks_raw[0], cores)
        except Exception as err:
            compiled = {"unavailable": repr(err)}
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "candidates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64 solve / f32 score",
            "data": "synthetic", "config": {"workload": "cfg4", "candidates_per_sweep": C_SWEEP, "sweeps_per_step": BUNDLE,
                                            "candidates_per_step": BUNDLE * C_SWEEP, "matches": N_MATCH,
                                            "hypotheses": H_HYP, "mode": "reference semantics (early exit)"},
            "cpu_baseline": {"value": val, "unit": "candidates/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} of the {BUNDLE} sweeps of a step ({sample * C_SWEEP} candidates per step), one solver call "
                                       "per core (BASELINE.md mode B); 4-point null space " +
                                       ("by the 12x12 eigen-solve (the reference's structure)" if EIGEN else
                                        "by Householder QR as on the device (the port's faster mode)"),
                             "evals_per_s": ev_tot / t_tot},
            "e2e": {"value": val, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if compiled is not None:
        line["compiled_reference"] = compiled
    print(json.dumps(line))
    return 0


def _compiled_ref_worker(bj, lo, hi, barrier, out):
    """one host process = one thread of the reference: PnPsolver keeps its state in rand() and in function statics
    (PnPsolver.cpp:625-642), so the compiled reference runs one solver at a time per PROCESS"""
    import ref_api as R
    from ransac_b200 import synth
    s2 = synth.level_sigma2()
    solvers = []
    for c in range(lo, hi):
        octave = np.searchsorted(s2, bj["sigma2"][c]).astype(np.int32)
        sv = R.PnP(bj["p2d"][c], octave, s2, bj["p3d"][c], bj["K"])
        sv.set_params(PRM["prob"], PRM["min_inliers"], PRM["max_its"], PRM["min_set"], PRM["eps"], PRM["th2"])
        solvers.append((sv, int(bj["seeds"][c])))
    barrier.wait()
    t0 = time.perf_counter()
    n_ok = 0
    for sv, sd in solvers:
        R.seed(sd)
        n_ok += int(sv.iterate(5)["ok"])        # Tracking.cpp:1255; the first call runs to the stopping point (PnPsolver.cpp:119)
    out.put((t0, time.perf_counter(), hi - lo, n_ok))


def compiled_reference_leg(bj, cores):
    """The reference's OWN PnPsolver.cpp (compiled unmodified into oracle/_ref/libref_solvers.so; its Eigen calls land in the
    stand-in headers' dense kernels -- Jacobi eigen-solves, one-sided Jacobi SVD -- not in Eigen's own) on a bounded sample of the same
    workload, one process per host core.  Reported beside the port; the port's QR mode stays the headline because it is the faster
    CPU implementation.  None where the library is absent."""
    import multiprocessing as mp
    import ref_api as R
    if not os.path.exists(R.PATH):
        return None
    n = min(C_SWEEP, int(os.environ.get("RSAC_BENCH_COMPILED_REF_N", "1024")))
    ctx = mp.get_context("fork")
    nproc = max(1, min(cores, n))
    barrier, out = ctx.Barrier(nproc), ctx.Queue()
    per = (n + nproc - 1) // nproc
    procs = [ctx.Process(target=_compiled_ref_worker, args=(bj, min(n, w * per), min(n, (w + 1) * per), barrier, out)) for w in range(nproc)]
    for q in procs:
        q.start()
    try:
        got = [out.get(timeout=120) for _ in procs]
    except Exception as err:                                # a worker died: report it, never lose the reference line over this leg
        for q in procs:
            if q.is_alive():
                q.kill()
        return {"unavailable": "compiled-reference worker failed: %r" % (err,)}
    for q in procs:
        q.join()
    wall = max(g[1] for g in got) - min(g[0] for g in got)
    return {"value": n / wall, "unit": "candidates/s", "cores": nproc, "kind": "reference",
            "sample": f"{n} candidates of sweep 0, iterate(5) on each (runs to the reference's stopping point), one process per core; "
                      "wall clock from the first process' start to the last one's end",
            "candidates_ok": int(sum(g[3] for g in got)),
            "note": "src/PnPsolver.cpp compiled unmodified against the stand-in Eigen / OpenCV headers of oracle/shim (12x12 eigen-solve per "
                    "hypothesis as in the reference; dense kernels are the stand-ins', not Eigen's)"}


# --------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip cfg5/cfg1 side measurements and the CPU baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from ransac_b200 import capi, shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    numa_node = bind_to_gpu_numa(local_rank) if world > 1 and os.environ.get("RSAC_BENCH_NUMA", "1") == "1" else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    assert world in (1, 2, 4, 8) and BUNDLE % NBLOCK == 0, "sweeps are sharded over 1, 2, 4 or 8 ranks"

    # Several device batches are in flight per GPU (one engine + stream each): every kernel of a sweep is one
    # latency-bound wave (or a fraction of one), and a rare candidate that needs several Refine() calls, or the
    # clean-up stage, stretches a whole sweep -- overlapping batches fills those holes
    global PIPE
    if PIPE <= 0:
        PIPE = 6
    RUN_FLAGS = (0 if EXHAUSTIVE else capi.FLAG_EARLY_EXIT) | (capi.FLAG_EPNP_EIGEN if EIGEN else 0)
    # strong scaling: rank r owns candidates [first, first + count) of EVERY sweep; its shards of `world` consecutive
    # sweeps (blocks j*world .. j*world + world - 1) are concatenated into device batch j (1024 candidates)
    first, count = shard_of(C_SWEEP, rank, world)
    NB = NBLOCK // world                     # distinct device batches of this rank
    RUNS = BUNDLE // world                   # batches per step and GPU
    C_RUN = count * world                    # = 1024
    prm = capi.ransac_params(**PRM)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    offsets = (np.arange(C_RUN + 1, dtype=np.int64) * N_MATCH).astype(np.int32)
    blocks_full = [make_block(j) for j in range(NBLOCK)]
    batches = []
    for j in range(NB):
        sel = [blocks_full[j * world + i] for i in range(world)]
        cat = lambda k, shp: np.concatenate([b_[k][first:first + count] for b_ in sel]).reshape(shp)
        ids = np.concatenate([(j * world + i) * C_SWEEP + first + np.arange(count) for i in range(world)]).astype(np.int32)
        # indexed wire format of the same batch: the keypoint tables of its `world` frames one after another, map-point
        # indices into the resident map (the blocks' maps concatenated)
        kpi = np.concatenate([b_["kp_idx"][first:first + count].astype(np.int64) + i * N_KP for i, b_ in enumerate(sel)]).astype(np.uint16)
        mpi = np.concatenate([b_["mp_idx"][first:first + count].astype(np.int64) + (j * world + i) * N_MAP
                              for i, b_ in enumerate(sel)]).astype(np.uint32)
        batches.append(dict(p3d=pin(cat("p3d", (-1, 3))).numpy(), p2d=pin(cat("p2d", (-1, 2))).numpy(), s2=pin(cat("sigma2", (-1,))).numpy(),
                            kp_idx=pin(kpi.reshape(-1)), mp_idx=pin(mpi.reshape(-1)),
                            kp_uv=pin(np.concatenate([b_["kp_uv"] for b_ in sel])), kp_s2=pin(np.concatenate([b_["kp_sigma2"] for b_ in sel])),
                            seeds=np.concatenate([b_["seeds"][first:first + count] for b_ in sel]), K=sel[0]["K"], ids=ids))
    the_map = np.ascontiguousarray(np.concatenate([b_["mp_xyz"] for b_ in blocks_full]), np.float32)     # resident on every engine
    b = blocks_full[0]

    NRES = max(NB, PIPE)                 # resident pass: engine i keeps batch (i mod NB) uploaded
    NSLOT = PIPE + 1                     # end-to-end pass: engines 0 .. PIPE are re-uploaded every run
    NENG = max(NRES, NSLOT)
    engines, streams, d_local = [], [], []
    for i in range(NENG):
        e = capi.Engine(local_rank)
        s = torch.cuda.Stream(device=dev)
        e.set_stream(s.cuda_stream)
        engines.append(e)
        streams.append(s)
        d_local.append(torch.full((C_RUN, shard.REC_WORDS), -1, dtype=torch.int32, device=dev))
    words_total = int(((np.diff(offsets) + 31) // 32).sum())
    h_res = [torch.empty((C_RUN, shard.REC_WORDS), dtype=torch.int32).pin_memory() for _ in range(NSLOT)]
    h_msk = [torch.empty((max(words_total, 1),), dtype=torch.int32).pin_memory() for _ in range(NSLOT)]

    def upload(i, j):
        """flat arrays (24 B per correspondence): the resident pass, whose inputs are uploaded before the timed region"""
        blk = batches[j % NB]
        engines[i].set_problem_ids(blk["ids"])
        engines[i].pnp_upload(offsets, blk["p3d"], blk["p2d"], blk["s2"], [blk["K"]], prm, seeds=blk["seeds"])

    def upload_indexed(i, j, with_map=False):
        """the end-to-end pass: per batch the frames' keypoint tables (12 B per keypoint) and 6 B per correspondence --
        (keypoint u16, map point u32) -- from pinned memory; the map stays resident on the device"""
        blk = batches[j % NB]
        engines[i].set_problem_ids(blk["ids"])
        engines[i].pnp_upload_indexed(offsets, blk["kp_idx"].numpy(), blk["mp_idx"].numpy(), blk["K"], prm, seeds=blk["seeds"],
                                      kp_uv=blk["kp_uv"].numpy(), kp_sigma2=blk["kp_s2"].numpy(), mp_xyz=the_map if with_map else None)

    # multi-GPU: the all-gather of a batch's records (98 KB per rank, latency-bound) runs on a side stream, in issue
    # order; a ring of gather buffers, and every engine waits for the gather that last read its records
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    NG = NENG + 2
    d_gath = [torch.empty((world * C_RUN, shard.REC_WORDS), dtype=torch.int32, device=dev) for _ in range(NG)] if world > 1 else None
    gather_done = [None] * NENG
    last_gather = [None]

    def run_and_gather(i, k):
        st = streams[i]
        if gather_done[i] is not None:
            st.wait_event(gather_done[i])              # the gather that last read this engine's records
        engines[i].pnp_run(RUN_FLAGS, d_local[i].data_ptr())
        if world > 1:
            ev = torch.cuda.Event()
            ev.record(st)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(ev)
                dist.all_gather_into_tensor(d_gath[k % NG], d_local[i])
                gd = torch.cuda.Event()
                gd.record(comm_stream)
                gather_done[i] = gd
            last_gather[0] = d_gath[k % NG]
        else:
            last_gather[0] = d_local[i]

    def bound_in_flight(i, k, ring):
        # at most PIPE batches in flight: batch k starts after batch k - PIPE has finished
        prev = ring[(k - PIPE) % len(ring)] if k >= PIPE else None
        if prev is not None:
            streams[i].wait_event(prev)

    res_ring = [None] * (NRES + PIPE)
    e2e_ring = [None] * (NSLOT + PIPE)

    def run_resident(k):
        i = k % NRES
        with torch.cuda.stream(streams[i]):
            bound_in_flight(i, k, res_ring)
            run_and_gather(i, k)
            ev = torch.cuda.Event()
            ev.record(streams[i])
            res_ring[k % len(res_ring)] = ev

    def run_e2e(k):
        i = k % NSLOT
        with torch.cuda.stream(streams[i]):
            upload_indexed(i, k)                        # H2D of batch k mod NB
            bound_in_flight(i, k, e2e_ring)
            run_and_gather(i, k)
            ev = torch.cuda.Event()
            ev.record(streams[i])
            e2e_ring[k % len(e2e_ring)] = ev
            # D2H of this batch's records and inlier masks into pinned memory (async on the batch's stream)
            engines[i].pnp_download_async(h_res[i].data_ptr(), h_msk[i].data_ptr())

    issue_ms = [0.0]

    def run_e2e_flat(k):
        """the same with the flat wire format (24 B per correspondence, no resident map): reported beside the headline e2e"""
        i = k % NSLOT
        with torch.cuda.stream(streams[i]):
            upload(i, k)
            bound_in_flight(i, k, e2e_ring)
            run_and_gather(i, k)
            ev = torch.cuda.Event()
            ev.record(streams[i])
            e2e_ring[k % len(e2e_ring)] = ev
            engines[i].pnp_download_async(h_res[i].data_ptr(), h_msk[i].data_ptr())

    def timed(fn, steps):
        """`steps` steps = steps * RUNS batches on this rank; device time, max over ranks"""
        for r in (res_ring, e2e_ring):
            for q in range(len(r)):
                r[q] = None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        main_s = torch.cuda.current_stream()
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(main_s)
        for s in streams:
            s.wait_event(ev0)
        t_host = time.perf_counter()
        for k in range(steps * RUNS):
            fn(k)
        issue_ms[0] = (time.perf_counter() - t_host) * 1e3 / max(1, steps * RUNS)    # host time to ISSUE one batch
        for s in streams + ([comm_stream] if comm_stream is not None else []):
            e = torch.cuda.Event()
            e.record(s)
            main_s.wait_event(e)
        ev1.record(main_s)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = float(t.item())
        return ms

    for i in range(NRES):
        upload(i, i)
    torch.cuda.synchronize()
    timed(run_resident, max(1, (args.warmup * RUNS + RUNS - 1) // RUNS))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sum(e.launch_count() for e in engines)
    ms = timed(run_resident, args.steps)
    issue_res = issue_ms[0]
    launches = sum(e.launch_count() for e in engines) - launches0
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel durations: CUDA events around every launch of a NON-pipelined pass (one batch at a time on one
    # stream, eager launches), so that an event pair brackets exactly one kernel
    stage = {}
    e0 = engines[0]
    e0.profile_reset()
    e0.profile_enable(True)
    nprof = 8
    with torch.cuda.stream(streams[0]):
        for _ in range(nprof):
            e0.pnp_run(RUN_FLAGS, d_local[0].data_ptr())
    torch.cuda.synchronize()
    for k, (tms, nl) in e0.profile().items():
        stage[k] = [tms, nl]
    trace = e0.profile_trace()
    trace = trace[-(len(trace) // nprof):] if trace else []      # the launches of the last batch, in order
    e0.profile_enable(False)
    # stages and hypotheses actually solved and scored: mean over the batches this rank holds
    stats = [engines[i].pnp_phase_stats() for i in range(NB)]
    ha = stats[0][0]
    n_b = float(np.mean([st[1] for st in stats]))
    n_c = float(np.mean([st[2] for st in stats]))
    hyp_done = float(np.mean([st[3] for st in stats]))

    # ---- content checks (rank 0 asserts; every rank takes part in the collectives)
    torch.cuda.synchronize()
    checks = {}
    # (1) the resident records of batch 0, as this rank computed them
    with torch.cuda.stream(streams[0]):
        engines[0].pnp_run(RUN_FLAGS, d_local[0].data_ptr())
    torch.cuda.synchronize()
    res_local, msk_local = engines[0].pnp_download()
    # (2) end to end (timed), then one more end-to-end batch 0 whose downloaded records and masks must equal the resident run's
    for i in range(NSLOT):                              # the map becomes resident on the end-to-end engines (not timed)
        with torch.cuda.stream(streams[i]):
            upload_indexed(i, i, with_map=True)
    torch.cuda.synchronize()
    timed(run_e2e, 1)
    ms_e2e = timed(run_e2e, args.steps)
    issue_e2e = issue_ms[0]
    timed(run_e2e_flat, 1)
    ms_e2e_flat = timed(run_e2e_flat, max(1, args.steps // 2))
    steps_flat = max(1, args.steps // 2)
    with torch.cuda.stream(streams[0]):
        upload_indexed(0, 0)
        engines[0].pnp_run(RUN_FLAGS, d_local[0].data_ptr())
        engines[0].pnp_download_async(h_res[0].data_ptr(), h_msk[0].data_ptr())
    torch.cuda.synchronize()
    got = h_res[0].numpy().view(capi.RESULT_DTYPE).reshape(-1)
    assert got.tobytes() == res_local.tobytes(), "end-to-end records differ from the resident run's"
    assert (h_msk[0].numpy().view(np.uint32)[:words_total] == msk_local).all(), "end-to-end masks differ from the resident run's"
    checks["e2e_equals_resident"] = True
    # (3) N ranks == 1 rank: gather batch 0 (= sweeps 0 .. world-1) through torch.distributed and through the library's own
    # NCCL path; rank 0 recomputes those sweeps alone (unsharded) and compares byte for byte
    if world > 1:
        g_torch = torch.empty((world * C_RUN, shard.REC_WORDS), dtype=torch.int32, device=dev)
        g_native = torch.zeros_like(g_torch)
        dist.all_gather_into_tensor(g_torch, d_local[0])
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.Engine.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        engines[0].nccl_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        engines[0].nccl_allgather_results(d_local[0].data_ptr(), C_RUN, g_native.data_ptr())
        engines[0].sync()
        torch.cuda.synchronize()
        engines[0].nccl_destroy()
        rec_t = shard.records_from_tensor(g_torch)
        rec_n = shard.records_from_tensor(g_native)
        if rank == 0:
            assert rec_t.tobytes() == rec_n.tobytes(), "rsac_nccl_allgather_results differs from torch.distributed's all-gather"
            assert len(rec_t) == world * C_SWEEP and (rec_t["problem"] == np.arange(world * C_SWEEP)).all(), "gather lost candidates"
            one = capi.Engine(local_rank)
            for sw in range(world):
                bs = blocks_full[sw]
                one.set_problem_ids(None)
                one.set_problem_base(sw * C_SWEEP)
                r1, _ = one.pnp_solve((np.arange(C_SWEEP + 1) * N_MATCH).astype(np.int32), bs["p3d"], bs["p2d"], bs["sigma2"], [bs["K"]], prm,
                                      seeds=bs["seeds"], flags=RUN_FLAGS)
                assert r1.tobytes() == rec_t[sw * C_SWEEP:(sw + 1) * C_SWEEP].tobytes(), f"sweep {sw}: {world}-rank records differ from the 1-rank run"
            one.close()
            checks["n_rank_equals_1_rank_bytes"] = True
            checks["native_nccl_allgather_equals_torch"] = True
        n_ok = int(rec_t["ok"].sum())
    else:
        rec = shard.records_from_tensor(last_gather[0])
        assert len(rec) == C_RUN
        n_ok = int(rec["ok"].sum())

    h2d = int(C_RUN * N_MATCH * 6 + world * N_KP * 12 + C_RUN * 4 + C_RUN * 152 + C_RUN * 4)
    d2h = int(C_RUN * 96 + words_total * 4)
    C_STEP = BUNDLE * C_SWEEP
    value = C_STEP * args.steps / (ms * 1e-3)
    e2e_v = C_STEP * args.steps / (ms_e2e * 1e-3)
    hyp_frac = hyp_done / float(C_RUN * H_HYP)                   # share of the 300 x candidates hypotheses computed

    line = {"metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None, "dtype": "f64 solve / f32 score", "data": "synthetic",
            "config": {"workload": "cfg4", "candidates_per_sweep": C_SWEEP, "sweeps_per_step": BUNDLE, "candidates_per_step": C_STEP,
                       "candidates_per_gpu_per_sweep": count, "matches": N_MATCH, "hypotheses": H_HYP, "outliers": 0.5,
                       "mode": ("all H hypotheses solved and scored on the device, then reference-semantics replay + Refine per candidate"
                                if EXHAUSTIVE else
                                (f"reference semantics with early exit in stages: hypotheses [0,{ha}) of every candidate, the "
                                 f"following stages for the {n_b:.0f} of {C_RUN} candidates (mean over batches) without an acceptable "
                                 f"hypothesis so far ({100 * hyp_frac:.1f} % of the 300 x {C_RUN} hypotheses solved and scored), "
                                 "replay + Refine per candidate; records identical to the exhaustive run")) +
                               ("; 4-point null space by the 12x12 eigen-solve (RSAC_FLAG_EPNP_EIGEN)" if EIGEN else
                                "; 4-point null space by Householder QR (DESIGN.md section 2)"),
                       "blocks": (f"{NBLOCK} distinct synthetic sweeps are cycled (sweep s of a step is block s mod {NBLOCK}); a sweep is one frame of "
                                  f"{N_KP} keypoints matched against {C_SWEEP} candidate keyframes of a {N_MAP}-point map, {N_MATCH} (keypoint, map point) "
                                  "pairs per candidate, half of them wrong map points (Tracking.cpp:1196-1232)"),
                       "parallelism": (f"every sweep sharded over {world} GPU(s) in contiguous blocks of {count} candidates; a rank concatenates its "
                                       f"shards of {world} consecutive sweeps into one device batch of {C_RUN} candidates ({RUNS} batches per step "
                                       f"and GPU, {PIPE} in flight, one engine + stream each; CUDA graph per batch); all-gather of the batch's "
                                       "records (96 B per candidate) after every batch; e2e uploads every batch from pinned host memory in the indexed "
                                       "wire format (rsac_pnp_upload_indexed: the frames' keypoint tables + 6 B per correspondence; the map is "
                                       "resident on the device, as the map of a running SLAM system is) and reads records + masks back inside the "
                                       "timed region; the resident pass holds the same batches uploaded flat (24 B per correspondence), and the "
                                       "content check compares the two"),
                       "l2": (f"no explicit flush: {NB} distinct batches per GPU are cycled, each in its own engine (~90 MB of device buffers per "
                              "batch against 126 MB of L2); within a sweep the working set is L2-resident by design and every kernel is "
                              "compute- or latency-bound")},
            "e2e": {"value": e2e_v, "unit": "candidates/s", "h2d_bytes_per_step": h2d * RUNS * world, "d2h_bytes_per_step": d2h * RUNS * world,
                    "ms_per_step": ms_e2e / args.steps,
                    "ms_per_batch": ms_e2e / args.steps / RUNS, "host_issue_ms_per_batch": issue_e2e,
                    "resident_ms_per_batch": ms / args.steps / RUNS, "resident_host_issue_ms_per_batch": issue_res,
                    "flat_upload": {"value": C_STEP * steps_flat / (ms_e2e_flat * 1e-3), "unit": "candidates/s",
                                    "h2d_bytes_per_step": int(C_RUN * N_MATCH * 24 + C_RUN * 160) * RUNS * world,
                                    "note": "the same end-to-end pass with rsac_pnp_upload (24 B per correspondence, nothing resident)"}},
            "gpu_launches": int(launches),
            "checks": checks,
            "host_binding": ("rank 0 bound to NUMA node %d of its GPU (every rank binds itself the same way)" % numa_node) if numa_node is not None
                            else ("unbound (N = 1)" if world == 1 else "NUMA topology not exposed: ranks unbound"),
            "extras": {"evals_per_s": value * H_HYP * N_MATCH * hyp_frac, "e2e_evals_per_s": e2e_v * H_HYP * N_MATCH * hyp_frac,
                       "evals_note": "hypothesis x correspondence evaluations actually performed (early exit skips the rest)",
                       "ms_per_sweep": ms / args.steps / BUNDLE, "e2e_ms_per_sweep": ms_e2e / args.steps / BUNDLE,
                       "candidates_ok": n_ok,
                       "phases": {"first_phase": ha, "candidates_phase_b": n_b, "candidates_phase_c": n_c,
                                  "hypotheses_done_frac": hyp_frac},
                       "sweep_launches_ms": [[k, round(m, 4)] for k, m in trace],
                       "stage_ms_per_sweep": {k: v[0] / nprof for k, v in stage.items() if v[1]},
                       "stage_ms_per_launch": {k: (v[0] / v[1] if v[1] else None) for k, v in stage.items()}}}
    if clocks is not None:
        line["clocks"] = clocks

    if rank == 0:
        peaks, peak_src = measured_peaks()
        fp32_pk, fp64_pk = engines[0].measure_peaks()
        line["extras"]["measured_fp32_tflops"] = fp32_pk
        line["extras"]["measured_fp64_tflops"] = fp64_pk
        # per-kernel rooflines; "roofline" is the kernel with the largest share of the sweep.
        # Algorithmic FLOP per 4-point solve is counted by the oracle's instrumented build.
        flop_per_solve = epnp_flops_per_solve(b, eigen=EIGEN)
        per = {k: v[0] / nprof for k, v in stage.items() if v[1]}
        nl = {k: v[1] / nprof for k, v in stage.items() if v[1]}
        total_ms = sum(per.values()) or 1.0
        first_l = {}
        for k, m in trace:
            first_l.setdefault(k, m)
        kern_solve = ("epnp_minimal_kernel<eigen>" if EIGEN else "epnp_minimal_subwarp_kernel")
        roofs = {}
        if per.get("solve"):
            t = per["solve"] * 1e-3
            ach = flop_per_solve * hyp_done / t / 1e12
            roofs["solve"] = {"bound": "fp64", "kernel": kern_solve, "achieved": ach, "peak": fp64_pk,
                              "unit": "TFLOP/s", "frac": ach / fp64_pk,
                              "traffic": None,
                              "peak_source": "DFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                              "flop_per_solve": flop_per_solve, "solves_per_sweep": hyp_done, "launch_ms": per["solve"],
                              "launches_per_sweep": nl["solve"], "share_of_sweep": per["solve"] / total_ms}
            if not EXHAUSTIVE and first_l.get("solve"):
                # the stage-0 launch (every candidate's first hypotheses) is the dominant launch of the sweep; the later
                # stages launch the same kernel on a fraction of a wave (latency-bound by design, they overlap with other
                # sweeps).  `roofline` quotes the stage-0 launch; the launch-weighted figure over all launches stays beside it
                a1 = flop_per_solve * C_RUN * ha / (first_l["solve"] * 1e-3) / 1e12
                agg = dict(achieved=roofs["solve"]["achieved"], frac=roofs["solve"]["frac"], launch_ms=roofs["solve"]["launch_ms"],
                           solves=hyp_done, launches_per_sweep=nl["solve"])
                tr = ncu_traffic(kern_solve)
                roofs["solve"].update({"achieved": a1, "frac": a1 / fp64_pk, "launch_ms": first_l["solve"], "solves_per_launch": C_RUN * ha,
                                       "launch": "stage 0 (hypotheses [0,%d) of %d candidates)" % (ha, C_RUN),
                                       "all_launches_of_a_sweep": agg, "share_of_sweep": first_l["solve"] / total_ms,
                                       "traffic": tr["dram_bytes"] if tr else None,
                                       "traffic_source": (tr or {}).get("source")})
                roofs["solve"].pop("solves_per_sweep", None)
                roofs["solve"].pop("launches_per_sweep", None)
        if per.get("score"):
            t = per["score"] * 1e-3
            ach = FLOP_PER_EVAL * hyp_done * N_MATCH / t / 1e12
            roofs["score"] = {"bound": "fp32", "kernel": "score_kernel<HPL,0> (cfg4 sweep)", "achieved": ach, "peak": fp32_pk,
                              "unit": "TFLOP/s", "frac": ach / fp32_pk, "traffic": None,
                              "peak_source": "FFMA micro-kernel measured in this run", "flop_per_eval": FLOP_PER_EVAL,
                              "launch_ms": per["score"], "launches_per_sweep": nl["score"], "share_of_sweep": per["score"] / total_ms}
        if per.get("select"):
            roofs["select"] = {"bound": "latency", "kernel": "ransac_select_kernel<0> (replay + Refine, one CTA per candidate)",
                               "launch_ms": per["select"], "launches_per_sweep": nl["select"], "share_of_sweep": per["select"] / total_ms,
                               "note": "serial dense tails (2 sqrt + 1-2 div chains); no meaningful FLOP roofline"}
        dom = max(roofs, key=lambda k: per.get(k, 0.0) if "achieved" in roofs[k] else 0.0) if roofs else None   # largest share of a sweep
        if dom and "achieved" in roofs[dom]:
            line["roofline"] = roofs[dom]
        elif roofs.get("solve"):
            line["roofline"] = roofs["solve"]
        line["extras"]["rooflines"] = roofs
    if rank == 0 and world == 1 and not args.no_extras:
        line.update(extras_single_gpu(engines[0], peaks, peak_src, fp32_pk, line))
    if rank == 0:
        print(json.dumps(line))
    for e in engines:
        e.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def ncu_traffic(key):
    """per-launch DRAM bytes of a kernel from the committed ncu capture (profiles/r02_traffic.json): NOT measured in this
    run -- the entry names the capture it comes from (kernel, launch shape, solves per launch)"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        for k, v in d.items():
            if k.startswith(key):
                return {"dram_bytes": int(v["dram_bytes"]), "source": v.get("source", "profiles/r02_traffic.json")}
    except Exception:
        pass
    return None


def epnp_flops_per_solve(b, eigen=False):
    """algorithmic FP64 FLOP of one 4-point EPnP solve, from the oracle's operation counters"""
    try:
        import oracle_api as O
        O.build()
        pb = O.pnp_problem(b["p3d"][0], b["p2d"][0], b["sigma2"][0], b["K"])
        tab = O.index_table(int(b["seeds"][0]), N_MATCH, 4, 64)
        return float(O.epnp_flops(pb, tab, 0 if eigen else O.FLAG_EPNP_QR_NULLSPACE))
    except Exception:
        return 9.0e4   # order-of-magnitude fallback


def extras_single_gpu(eng, peaks, peak_src, fp32_pk, line):
    """cfg5 scoring roofline, cfg1 single-frame latency and the CPU baseline (rank 0, N = 1 only)."""
    import torch
    from ransac_b200 import capi, synth
    import oracle_api as O
    out = {}
    # ---- cfg5: 4096 poses x 10 000 correspondences, CheckInliers only
    H, N = 4096, 10000
    p = synth.scoring_stress(5000, H, N)
    max_err = (p["sigma2"] * np.float32(5.991)).astype(np.float32)
    eng.set_stream(None)
    eng.score_pnp_upload(p["poses"], p["p3d"], p["p2d"], max_err, p["K"])
    for _ in range(10):
        eng.score_pnp_run(True)
    eng.sync()
    reps = 200
    eng.timer_begin()
    for _ in range(reps):
        eng.score_pnp_run(True)
    ms = eng.timer_end() / reps
    ev = H * N / (ms * 1e-3)
    words = (N + 31) // 32
    alg_bytes = H * 48 + N * 24 + H * words * 4 + H * 4
    out["roofline_score"] = {"bound": "fp32", "kernel": "score_kernel<2,0> (cfg5: 4096 x 10000, masks + counts)",
                             "achieved": ev * FLOP_PER_EVAL / 1e12, "peak": fp32_pk, "unit": "TFLOP/s",
                             "frac": ev * FLOP_PER_EVAL / 1e12 / fp32_pk,
                             "peak_source": "FFMA micro-kernel measured in this run; theoretical 148 SM x 128 x 2 x 1.965 GHz = 74.4",
                             "evals_per_s": ev, "launch_ms": ms, "flop_per_eval": FLOP_PER_EVAL,
                             "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (ms * 1e-3) / 1e9,
                                     "peak_gbs": peaks["hbm_gbs"], "peak_source": peak_src + " (MEASURED_PEAKS.json)"},
                             "exact_path_evals": eng.score_exact_evals(),
                             "note": "launches run back to back on 5.6 MB of data: L2-resident by design (compute-bound kernel)"}
    # independent scoring jobs in flight (one engine + stream each), like the sweeps of the headline number: the ramp
    # and tail of one launch (19 % of its duration: sm__cycles_active 60.5 k of 74.5 k elapsed,
    # profiles/r01_score_cfg5_ncu_full.txt) overlap its neighbours.  Timed on the device: one event before all streams
    # start, one after they have all finished.
    single = dict(out["roofline_score"])
    try:
        NF = 6
        pool, pstreams = [], []
        for _ in range(NF):
            q = capi.Engine(0)
            st_ = torch.cuda.Stream()
            q.set_stream(st_.cuda_stream)
            q.score_pnp_upload(p["poses"], p["p3d"], p["p2d"], max_err, p["K"])
            for _ in range(3):
                q.score_pnp_run(True)
            pool.append(q)
            pstreams.append(st_)
        torch.cuda.synchronize()
        main_s = torch.cuda.current_stream()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(main_s)
        for st_ in pstreams:
            st_.wait_event(ev0)
        per = 100
        for _ in range(per):
            for q in pool:
                q.score_pnp_run(True)
        for st_ in pstreams:
            e_ = torch.cuda.Event()
            e_.record(st_)
            main_s.wait_event(e_)
        ev1.record(main_s)
        torch.cuda.synchronize()
        msf = ev0.elapsed_time(ev1) / (per * NF)
        evf = H * N / (msf * 1e-3)
        out["roofline_score"].update({
            "kernel": "score_kernel<2,0> (cfg5: 4096 x 10000, masks + counts), %d independent scoring jobs in flight" % NF,
            "achieved": evf * FLOP_PER_EVAL / 1e12, "frac": evf * FLOP_PER_EVAL / 1e12 / fp32_pk, "evals_per_s": evf,
            "launch_ms": msf, "launches_timed": per * NF,
            "timing": "CUDA events around %d launches on %d streams; launch_ms = timed region / launches" % (per * NF, NF),
            "single_stream": {k: single[k] for k in ("achieved", "frac", "evals_per_s", "launch_ms")}})
        out["roofline_score"]["hbm"]["achieved_gbs"] = alg_bytes / (msf * 1e-3) / 1e9
        for q in pool:
            q.close()
    except Exception as err:
        out["roofline_score"]["in_flight_error"] = repr(err)
    # ---- cfg1: one frame, one candidate (latency of the reference-facing call)
    b1 = synth.pnp_batch(1, 1, N_MATCH, 0.5)
    off1 = np.array([0, N_MATCH], np.int32)
    prm = capi.ransac_params(**PRM)
    for _ in range(3):
        eng.pnp_solve(off1, b1["p3d"], b1["p2d"], b1["sigma2"], [b1["K"]], prm, seeds=b1["seeds"])
    t0 = time.perf_counter()
    for _ in range(20):
        eng.pnp_solve(off1, b1["p3d"], b1["p2d"], b1["sigma2"], [b1["K"]], prm, seeds=b1["seeds"])
    out.setdefault("extras", dict(line["extras"]))["cfg1_single_candidate_ms"] = (time.perf_counter() - t0) / 20 * 1e3

    # ---- CPU baseline: the oracle port of PnPsolver on the host cores
    O.build()
    cores = os.cpu_count() or 1
    nb = 256
    bb = synth.pnp_batch(4, nb, N_MATCH, 0.5)
    pbs = [O.pnp_problem(bb["p3d"][c], bb["p2d"][c], bb["sigma2"][c], bb["K"]) for c in range(nb)]
    tabs = [O.index_table(int(s), N_MATCH, 4, H_HYP) for s in bb["seeds"]]
    oprm = O.params(**PRM)
    qr = O.FLAG_EPNP_QR_NULLSPACE
    dt1, ev1, _ = O.pnp_batch(pbs[:64], oprm, tabs[:64], qr, 1)
    dtn, evn, _ = O.pnp_batch(pbs, oprm, tabs, qr, cores)
    dtx, evx, _ = O.pnp_batch(pbs[:32], oprm, tabs[:32], O.FLAG_EXHAUSTIVE | qr, 1)
    dte, eve, _ = O.pnp_batch(pbs, oprm, tabs, 0, cores)
    out["cpu_baseline"] = {"value": nb / dtn, "unit": "candidates/s", "cores": cores, "kind": "port",
                           "sample": f"{nb} cfg4 candidates, reference semantics (early exit), one solver call per core, "
                                     "4-point null space by QR as on the device",
                           "single_thread_candidates_per_s": 64 / dt1, "single_thread_evals_per_s": ev1 / dt1,
                           "exhaustive_single_thread_evals_per_s": evx / dtx,
                           "exhaustive_single_thread_candidates_per_s": 32 / dtx,
                           "eigen_nullspace_candidates_per_s": nb / dte}

    # ---- cfg2: MLPnP, 64 frames x 1000 matches with bearing covariances; cfg3: Sim3, 200 matches, 300 iterations
    ex = out.setdefault("extras", dict(line["extras"]))
    try:
        C2, N2 = 64, 1000
        b2 = synth.pnp_batch(2, C2, N2, 0.5)
        cov = np.stack([synth.bearing_covariances(dict(K=b2["K"], sigma2=b2["sigma2"][c])) for c in range(C2)])
        Kf = np.array([b2["K"]], np.float32)
        off2 = (np.arange(C2 + 1) * N2).astype(np.int32)
        prm2 = capi.ransac_params(0.99, 10, 300, 6, 0.2, 5.991)
        eng.mlpnp_upload(off2, b2["p3d"], b2["p2d"], b2["sigma2"], Kf, prm2, cov=cov, seeds=b2["seeds"])
        for _ in range(2):
            eng.mlpnp_run()
        eng.sync()
        eng.timer_begin()
        for _ in range(5):
            eng.mlpnp_run()
        ms2 = eng.timer_end() / 5
        _, H2 = capi.pnp_ransac_setup(N2, prm2)
        res2, _ = eng.mlpnp_download()
        ex["cfg2_mlpnp"] = {"frames": C2, "matches": N2, "hypotheses": H2, "ms_per_batch": ms2,
                            "frames_per_s": C2 / (ms2 * 1e-3), "evals_per_s": C2 * H2 * N2 / (ms2 * 1e-3),
                            "frames_ok": int(res2["ok"].sum())}
        ex["cfg2_mlpnp"]["roofline"] = {"bound": "fp32 (scoring) / fp64 latency (one thread per 6-point solve)", "flop_per_eval": 30,
                                        "achieved": C2 * H2 * N2 * 30 / (ms2 * 1e-3) / 1e12, "peak": fp32_pk, "unit": "TFLOP/s",
                                        "frac": C2 * H2 * N2 * 30 / (ms2 * 1e-3) / 1e12 / fp32_pk,
                                        "note": "exhaustive: all 300 hypotheses of every frame are solved and scored"}
        # CPU port (oracle) on the host cores: reference semantics (stops at the first successful Refine)
        pbs2 = [O.mlpnp_problem(b2["p3d"][c], b2["p2d"][c], b2["sigma2"][c], tuple(Kf[0]), cov[c]) for c in range(C2)]
        tabs2 = [O.index_table(int(sd), N2, 6, H2) for sd in b2["seeds"]]
        oprm2 = O.params(0.99, 10, 300, 6, 0.2, 5.991)
        dta, eva, _ = O.mlpnp_batch(pbs2[:16], oprm2, tabs2[:16], 0, 1)
        dtb, evb, _ = O.mlpnp_batch(pbs2, oprm2, tabs2, 0, cores)
        ex["cfg2_mlpnp"]["cpu_port"] = {"single_thread_frames_per_s": 16 / dta, "single_thread_evals_per_s": eva / dta,
                                        "all_cores_frames_per_s": C2 / dtb, "cores": cores, "sample": "16 / 64 frames, early exit"}
        # a 64-frame batch is a fraction of one wave: independent batches in flight (one engine + stream each)
        pool = [capi.Engine(0) for _ in range(6)]
        for q in pool:
            q.mlpnp_upload(off2, b2["p3d"], b2["p2d"], b2["sigma2"], Kf, prm2, cov=cov, seeds=b2["seeds"])
            q.mlpnp_run()
        for q in pool:
            q.sync()
        t0 = time.perf_counter()
        reps2 = 20
        for _ in range(reps2):
            for q in pool:
                q.mlpnp_run()
        for q in pool:
            q.sync()
        dt2 = (time.perf_counter() - t0) / (reps2 * len(pool))
        ex["cfg2_mlpnp"]["batches_in_flight_6"] = {"ms_per_batch": dt2 * 1e3, "frames_per_s": C2 / dt2,
                                                   "evals_per_s": C2 * H2 * N2 / dt2, "timing": "wall clock around 120 batches, synchronised"}
        # the same with the staged early exit (RSAC_FLAG_EARLY_EXIT, MLPnPsolver.cpp:144-160: the reference returns at the first
        # successful Refine): fewer hypotheses per batch, so more batches share a wave; records equal the exhaustive run's
        try:
            ee_stats = {}
            for stg in ((128,), (96, 192), (48, 96, 192)):
                for q in pool:
                    q.set_stages(list(stg))
                    q.mlpnp_run(capi.FLAG_EARLY_EXIT)
                    q.mlpnp_run(capi.FLAG_EARLY_EXIT)
                for q in pool:
                    q.sync()
                r_ee, _ = pool[0].mlpnp_download()
                same = all((r_ee[f] == res2[f]).all() for f in ("ok", "n_inliers", "best_hyp", "n_refines", "n_hyp"))
                t0 = time.perf_counter()
                for _ in range(reps2):
                    for q in pool:
                        q.mlpnp_run(capi.FLAG_EARLY_EXIT)
                for q in pool:
                    q.sync()
                dte = (time.perf_counter() - t0) / (reps2 * len(pool))
                st2 = pool[0].mlpnp_phase_stats()
                ee_stats[",".join(str(v) for v in stg)] = {"ms_per_batch": dte * 1e3, "frames_per_s": C2 / dte, "records_equal_exhaustive": bool(same),
                                                           "hypotheses_done_frac": st2[3] / float(C2 * H2), "frames_in_stage_1": st2[1]}
            ex["cfg2_mlpnp"]["early_exit_batches_in_flight_6"] = ee_stats
            eng.set_stages([128])
            eng.mlpnp_run(capi.FLAG_EARLY_EXIT); eng.mlpnp_run(capi.FLAG_EARLY_EXIT); eng.sync()
            eng.timer_begin()
            for _ in range(5):
                eng.mlpnp_run(capi.FLAG_EARLY_EXIT)
            ex["cfg2_mlpnp"]["early_exit_one_batch_ms"] = eng.timer_end() / 5
            eng.set_stages([])
        except Exception as err:
            ex["cfg2_mlpnp"]["early_exit_error"] = repr(err)
        for q in pool:
            q.close()
        ps = [synth.sim3_problem(3000 + i, 200, 0.4, 1.0) for i in range(64)]
        cat = lambda k: np.concatenate([q[k] for q in ps])
        off3 = (np.arange(len(ps) + 1) * 200).astype(np.int32)
        K3 = np.array([ps[0]["K"]], np.float32)
        prm3 = capi.Sim3Params(0.99, 20, 300, 1)
        seeds3 = np.arange(len(ps), dtype=np.uint32) + 3000
        for C3 in (1, 64):
            o3 = off3[:C3 + 1]
            n3 = int(o3[-1])
            eng.sim3_upload(o3, cat("x1c")[:n3], cat("x2c")[:n3], cat("sigma2_1")[:n3], cat("sigma2_2")[:n3], K3, K3, prm3,
                            seeds=seeds3[:C3])
            for _ in range(2):
                eng.sim3_run()
            eng.sync()
            eng.timer_begin()
            for _ in range(10):
                eng.sim3_run()
            ms3 = eng.timer_end() / 10
            H3 = capi.sim3_ransac_setup(200, prm3)
            ex["cfg3_sim3_x%d" % C3] = {"candidates": C3, "matches": 200, "hypotheses": H3, "ms_per_batch": ms3,
                                        "evals_per_s": C3 * H3 * 200 / (ms3 * 1e-3),
                                        "roofline": {"bound": "fp32", "flop_per_eval": 63, "achieved": C3 * H3 * 200 * 63 / (ms3 * 1e-3) / 1e12,
                                                     "peak": fp32_pk, "unit": "TFLOP/s", "frac": C3 * H3 * 200 * 63 / (ms3 * 1e-3) / 1e12 / fp32_pk,
                                                     "note": "every hypothesis is scored (the reference stops at the first accepted one)"}}
        pbs3 = [O.sim3_problem(q["x1c"], q["x2c"], q["sigma2_1"], q["sigma2_2"], q["K"], q["K"], True) for q in ps]
        tabs3 = [O.index_table(int(sd), 200, 3, H3) for sd in seeds3]
        dt3a, ev3a, _ = O.sim3_batch(pbs3[:32], 0.99, 20, 300, tabs3[:32], 0, 1)
        dt3x, ev3x, _ = O.sim3_batch(pbs3[:32], 0.99, 20, 300, tabs3[:32], O.FLAG_EXHAUSTIVE, 1)
        ex["cfg3_sim3_cpu_port"] = {"single_thread_ms_per_candidate": 1e3 * dt3a / 32, "single_thread_evals_per_s": ev3a / dt3a,
                                    "exhaustive_single_thread_ms_per_candidate": 1e3 * dt3x / 32,
                                    "exhaustive_single_thread_evals_per_s": ev3x / dt3x, "sample": "32 candidates"}
    except Exception as err:   # the headline line must still be printed
        ex["cfg2_cfg3_error"] = repr(err)
    # ---- the relocalisation of ONE frame in the indexed wire format (first half): 1024 candidates x 500
    # (keypoint, map point) pairs over resident tables against the same sweep uploaded flat; one engine, one sweep at a time
    try:
        Ci, ni = 1024, 500
        fr = synth.reloc_frame(9, Ci, n_kp=2000, n_match=ni, n_map=200000)
        offi = (np.arange(Ci + 1) * ni).astype(np.int32)
        prmi = capi.ransac_params(**PRM)
        eng.pnp_upload_indexed(offi, fr["kp_idx"], fr["mp_idx"], fr["K"], prmi, seeds=fr["seeds"], kp_uv=fr["kp_uv"], kp_sigma2=fr["kp_sigma2"],
                               mp_xyz=fr["mp_xyz"])
        eng.pnp_run(capi.FLAG_EARLY_EXIT)
        ri, mi = eng.pnp_download()
        rf, mf = eng.pnp_solve(offi, fr["p3d"], fr["p2d"], fr["sigma2"], [fr["K"]], prmi, seeds=fr["seeds"], flags=capi.FLAG_EARLY_EXIT)
        assert ri.tobytes() == rf.tobytes() and (mi == mf).all(), "indexed upload differs from the flat upload"

        def sweep_indexed():
            eng.pnp_upload_indexed(offi, fr["kp_idx"], fr["mp_idx"], fr["K"], prmi, seeds=fr["seeds"])
            eng.pnp_run(capi.FLAG_EARLY_EXIT)
            return eng.pnp_download()

        def sweep_flat():
            return eng.pnp_solve(offi, fr["p3d"], fr["p2d"], fr["sigma2"], [fr["K"]], prmi, seeds=fr["seeds"], flags=capi.FLAG_EARLY_EXIT)

        tms = {}
        for name, fn in (("indexed", sweep_indexed), ("flat", sweep_flat)):
            for _ in range(3):
                fn()
            t0 = time.perf_counter()
            for _ in range(20):
                fn()
            tms[name] = (time.perf_counter() - t0) / 20 * 1e3
        ex["indexed_wire_format"] = {"candidates": Ci, "matches": ni, "keypoints": 2000, "map_points": 200000,
                                     "h2d_bytes_per_sweep_indexed": Ci * ni * 6, "h2d_bytes_per_sweep_flat": Ci * ni * 24,
                                     "ms_per_sweep_end_to_end_indexed": tms["indexed"], "ms_per_sweep_end_to_end_flat": tms["flat"],
                                     "records_identical": True, "candidates_ok": int(ri["ok"].sum()),
                                     "note": "host buffers pageable, one sweep at a time (upload + run + download, synchronous)"}
    except Exception as err:
        ex["indexed_wire_format_error"] = repr(err)
    # ---- KeyFrameDatabase::DetectRelocalizationCandidates, 64 lost frames against a resident database of
    # 4096 keyframes (KeyFrameDatabase.cpp:174-284); CPU port = the oracle's inverted-file walk, one thread
    try:
        Kd, Qd = 4096, 64
        dbk = synth.kf_database(21, K=Kd, n_places=256)
        qsk = [synth.kf_query(500 + i, dbk, (37 * i) % 256) for i in range(Qd)]
        eng.kfdb_upload(dbk)
        got = eng.kfdb_detect(qsk, mode=0)
        for _ in range(2):
            eng.kfdb_run()
        eng.sync()
        eng.timer_begin()
        for _ in range(10):
            eng.kfdb_run()
        msk = eng.timer_end() / 10
        t0 = time.perf_counter()
        for _ in range(3):
            eng.kfdb_detect(qsk, mode=0)
        e2ek = (time.perf_counter() - t0) / 3 * 1e3
        odbk = O.kfdb(dbk)
        stk = np.zeros(Kd, np.float32)
        want, dck = [], 0.0
        for (w, v) in qsk[:8]:
            want.append(O.detect_candidates(odbk, w, v, mode=0, score_state=stk))
            dck += O.kfdb_last_query_seconds() / 8
        nnz_db = int(dbk["bow_off"][-1])
        ex["candidate_retrieval"] = {"keyframes": Kd, "queries": Qd, "words_per_vector": nnz_db // Kd, "ms_per_batch": msk, "queries_per_s": Qd / (msk * 1e-3),
                                     "e2e_ms_per_batch": e2ek, "candidates_mean": float(np.mean([len(g) for g in got])),
                                     "first_8_equal_oracle": bool(all(g.tolist() == w.tolist() for g, w in zip(got[:8], want))),
                                     "cpu_port_single_thread_queries_per_s": 1.0 / dck,
                                     "hbm": {"algorithmic_bytes": Qd * nnz_db * 4, "achieved_gbs": Qd * nnz_db * 4 / (msk * 1e-3) / 1e9, "peak_gbs": peaks["hbm_gbs"]},
                                     "note": "KeyFrameDatabase.cpp:174-284 batched: every query against every keyframe's BowVector (the database is resident); "
                                             "CPU figure: the oracle's inverted-file walk + scoring of one query, the (per-call) rebuild of its inverted file excluded"}
    except Exception as err:
        ex["candidate_retrieval_error"] = repr(err)
    # ---- ORBmatcher::SearchBySim3 (ORBmatcher.cpp:948-1171), 64 keyframe pairs of ~1500 features each
    try:
        prs = [synth.kf_view_pair(700 + i, n_points=1200, n_extra=400, prematched=0.3) for i in range(8)]
        vws = [v for p_ in prs for v in (p_["kf1"], p_["kf2"])]
        NP = 64
        k1i, k2i = [2 * (i % 8) for i in range(NP)], [2 * (i % 8) + 1 for i in range(NP)]
        eng.sim3_search_upload(vws, k1i, k2i, [prs[i % 8]["K"] for i in range(NP)], [prs[i % 8]["R12"] for i in range(NP)],
                               [prs[i % 8]["t12"] for i in range(NP)], 7.5, [prs[i % 8]["matched12_in"] for i in range(NP)])
        for _ in range(3):
            eng.sim3_search_run()
        eng.sync()
        eng.timer_begin()
        for _ in range(20):
            eng.sim3_search_run()
        mss = eng.timer_end() / 20
        gm, gn = eng.sim3_search_download()
        ok1, ok2 = O.kf_view(prs[0]["kf1"]), O.kf_view(prs[0]["kf2"])
        t0 = time.perf_counter()
        for _ in range(5):
            wm, wn = O.search_by_sim3(ok1, ok2, prs[0]["K"], prs[0]["R12"], prs[0]["t12"], 7.5, prs[0]["matched12_in"])
        dcs = (time.perf_counter() - t0) / 5
        feats = sum(vws[a]["n_feat"] + vws[b_]["n_feat"] for a, b_ in zip(k1i, k2i))
        ex["search_by_sim3"] = {"pairs": NP, "features_per_keyframe_mean": feats / (2.0 * NP), "ms_per_batch": mss, "pairs_per_s": NP / (mss * 1e-3),
                                "matches_mean": float(np.mean(gn)), "pair_0_equals_oracle": bool(gm[0].tolist() == wm.tolist() and int(gn[0]) == wn),
                                "cpu_port_single_thread_pairs_per_s": 1.0 / dcs,
                                "note": "ORBmatcher.cpp:948-1171 batched: one thread per (pair, direction, map point), grid-window walk in the "
                                        "reference's order; 8 distinct synthetic pairs cycled; bit-identical to the oracle"}
    except Exception as err:
        ex["search_by_sim3_error"] = repr(err)
    # ---- ORBmatcher::SearchByProjection(Frame, KeyFrame, sAlreadyFound, 10, 100) (ORBmatcher.cpp:1317-1444), 64 pairs
    try:
        cs = [synth.proj_search_case(800 + i, n_points=1200, n_extra=400, clones=0.1) for i in range(8)]
        vwp = [v for c_ in cs for v in (c_["frame"], c_["kf"])]
        NPp = 64
        fi, ki = [2 * (i % 8) for i in range(NPp)], [2 * (i % 8) + 1 for i in range(NPp)]
        eng.proj_search_upload(vwp, fi, ki, [cs[i % 8]["K"] for i in range(NPp)], [cs[i % 8]["Rcw"] for i in range(NPp)],
                               [cs[i % 8]["tcw"] for i in range(NPp)], 10.0, 100, True, [cs[i % 8]["occupied"] for i in range(NPp)],
                               [cs[i % 8]["already_found"] for i in range(NPp)])
        for _ in range(3):
            eng.proj_search_run()
        eng.sync()
        eng.timer_begin()
        for _ in range(20):
            eng.proj_search_run()
        msp = eng.timer_end() / 20
        gmp, gnp, fellp, rndp = eng.proj_search_download()
        of, okf = O.kf_view(cs[0]["frame"]), O.kf_view(cs[0]["kf"])
        t0 = time.perf_counter()
        for _ in range(5):
            wmp, wnp = O.search_by_projection(of, okf, cs[0]["K"], cs[0]["Rcw"], cs[0]["tcw"], 10.0, 100, True, cs[0]["occupied"], cs[0]["already_found"])
        dcp = (time.perf_counter() - t0) / 5
        ex["search_by_projection"] = {"pairs": NPp, "ms_per_batch": msp, "pairs_per_s": NPp / (msp * 1e-3), "matches_mean": float(np.mean(gnp)),
                                      "assignment_rounds_max": int(rndp.max()), "sequential_fallbacks": int(fellp.sum()),
                                      "pair_0_equals_oracle": bool(gmp[0].tolist() == wmp.tolist() and int(gnp[0]) == wnp),
                                      "cpu_port_single_thread_pairs_per_s": 1.0 / dcp,
                                      "note": "ORBmatcher.cpp:1317-1444 batched; the reference's greedy keyframe-order assignment is reproduced "
                                              "exactly by preference lists + rounds (csrc/guided.cuh); bit-identical to the oracle"}
    except Exception as err:
        ex["search_by_projection_error"] = repr(err)
    # ---- the body of Tracking::Relocalization on the device (Tracking.cpp:1207-1284): SearchByBoW for every candidate -> PnP batch
    # built from the match arrays on the device (only the match counts visit the host) -> EPnP RANSAC sweep -> PoseOptimization
    try:
        Cw = 256
        wr = synth.reloc_world(77, C=Cw, n_kp=1500, n_kf_feat=1200)
        setsw = [wr["frame"]] + wr["kfs"]
        prmw = capi.ransac_params(0.99, 10, 300, 4, 0.5, 5.991)

        def reloc_once(first=False):
            eng.bow_upload(setsw, list(range(1, Cw + 1)), [0] * Cw, 0.75, True, 0) if first else None
            eng.bow_run()
            _, nmw = eng.bow_download()
            eng.pnp_upload_from_bow(nmw, wr["K"], prmw, wr["seeds"], 15, kp_uv=wr["kp_uv"] if first else None,
                                    kp_sigma2=wr["kp_sigma2"] if first else None, mp_xyz=wr["mp_xyz"] if first else None)
            eng.pnp_run(capi.FLAG_EARLY_EXIT)
            eng.poseopt_from_pnp()
            eng.poseopt_run()
            return eng.poseopt_download(), nmw

        (pow_, _), nmw = reloc_once(True)
        for _ in range(2):
            reloc_once()
        t0 = time.perf_counter()
        for _ in range(10):
            reloc_once()
        msw = (time.perf_counter() - t0) / 10 * 1e3
        Rw = np.stack([r_["Rf"].reshape(3, 3) for r_ in pow_])
        good = [int(c_) for c_ in range(Cw) if nmw[c_] >= 15 and pow_[c_]["n_inliers"] >= 50 and np.abs(Rw[c_] - wr["R"]).max() < 0.02]
        ex["relocalisation_pipeline"] = {"candidates": Cw, "frame_keypoints": 1500, "keyframe_features": 1200, "ms_per_relocalisation": msw,
                                         "candidates_per_s": Cw / (msw * 1e-3), "bow_matches_mean": float(np.mean(nmw)),
                                         "candidates_relocalised": len(good),
                                         "note": "SearchByBoW (resident keyframe sets) -> rsac_pnp_upload_from_bow -> EPnP RANSAC -> PoseOptimization, "
                                                 "wall clock per relocalisation of one frame against 256 candidates incl. the download of the "
                                                 "match counts and of the final records; a candidate counts as relocalised with >= 50 inliers "
                                                 "(Tracking.cpp:1318) at the true pose"}
    except Exception as err:
        ex["relocalisation_pipeline_error"] = repr(err)
    # ---- ORBmatcher::SearchByBoW, 1024 candidate keyframes against one frame (Tracking.cpp:1207-1232)
    try:
        Fb = synth.bow_frame(11, 1500, 100)
        kfs = [synth.bow_keyframe(1000 + i, Fb, 1200, shared=0.25, rot=7.0 * i) for i in range(16)]
        setsb = [Fb] + kfs
        qs = [1 + (i % 16) for i in range(1024)]
        tsb = [0] * 1024
        eng.bow_upload(setsb, qs, tsb, 0.75, True, 0)
        for _ in range(3):
            eng.bow_run()
        eng.sync()
        eng.timer_begin()
        for _ in range(20):
            eng.bow_run()
        msb = eng.timer_end() / 20
        mt, nmb = eng.bow_download()
        kq = [O.bow_features(s_) for s_ in setsb]
        t0 = time.perf_counter()
        for i in range(16):
            want, nw = O.search_by_bow(kq[1 + i], kq[0], 0.75, True, 0)
            assert nw == nmb[i] and (want == mt[i]).all(), "SearchByBoW differs from the oracle"
        dtb_ = (time.perf_counter() - t0) / 16
        # distances actually needed: per common node |KF features with a map point| x |frame features| (upper bound: none taken)
        alg_bytes = 32 * (1500 + 1024 * 1200) + 4 * 1024 * 1500
        ex["search_by_bow"] = {"pairs": 1024, "frame_features": 1500, "keyframe_features": 1200, "vocabulary_nodes": 100,
                               "ms_per_batch": msb, "pairs_per_s": 1024 / (msb * 1e-3), "matches_mean": float(np.mean(nmb)),
                               "cpu_port_single_thread_pairs_per_s": 1.0 / dtb_,
                               "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (msb * 1e-3) / 1e9, "peak_gbs": peaks["hbm_gbs"]},
                               "note": "ORBmatcher.cpp:110-239 batched; 16 distinct synthetic keyframes cycled; bit-identical to the oracle"}
    except Exception as err:
        ex["search_by_bow_error"] = repr(err)
    # ---- the reference's own null-space structure (RSAC_FLAG_EPNP_EIGEN: 12x12 eigen-solve per hypothesis), with early exit
    try:
        bce = synth.pnp_batch(4, 1024, 500, 0.5)
        offe = (np.arange(1025) * 500).astype(np.int32)
        eng.pnp_upload(offe, bce["p3d"], bce["p2d"], bce["sigma2"], [bce["K"]], capi.ransac_params(**PRM), seeds=bce["seeds"])
        fl = capi.FLAG_EARLY_EXIT | capi.FLAG_EPNP_EIGEN
        for _ in range(3):
            eng.pnp_run(fl)
        eng.sync()
        eng.timer_begin()
        for _ in range(10):
            eng.pnp_run(fl)
        mse = eng.timer_end() / 10
        rese, _ = eng.pnp_download()
        ex["eigen_mode"] = {"candidates": 1024, "ms_per_sweep": mse, "candidates_per_s": 1024 / (mse * 1e-3), "candidates_ok": int(rese["ok"].sum()),
                            "phases": list(eng.pnp_phase_stats()),
                            "cpu_port_all_cores_candidates_per_s": out.get("cpu_baseline", {}).get("eigen_nullspace_candidates_per_s"),
                            "note": "one engine, one sweep at a time; RSAC_BENCH_EIGEN=1 runs the whole bench in this mode"}
    except Exception as err:
        ex["eigen_mode_error"] = repr(err)
    # ---- Optimizer::PoseOptimization for every candidate of a cfg4-sized sweep (1024 frames x 250 matched
    # map points, 20 % outliers, monocular), one warp per frame; CPU port on a 64-frame sample beside it
    try:
        CP, NP = 1024, 250
        pp = [synth.poseopt_problem(4000 + i, NP, 0.2, 0.0) for i in range(CP)]
        offp = (np.arange(CP + 1) * NP).astype(np.int32)
        catp = lambda k: np.concatenate([q[k] for q in pp])
        Kp = np.stack([q["K"] for q in pp])
        Tp = np.stack([np.concatenate([q["Rcw"].ravel(), q["tcw"]]) for q in pp])
        argsp = (offp, catp("p3d"), catp("obs"), catp("inv_sigma2"), Kp, Tp)
        eng.poseopt_upload(*argsp)
        for _ in range(3):
            eng.poseopt_run()
        eng.sync()
        eng.timer_begin()
        for _ in range(20):
            eng.poseopt_run()
        msp = eng.timer_end() / 20
        resp, _ = eng.poseopt_download()
        t0 = time.perf_counter()
        for _ in range(10):
            eng.poseopt_solve(*argsp)
        dte = (time.perf_counter() - t0) / 10
        passes = float(resp["iterations"].sum() + resp["trials"].sum()) * NP       # edge visits of the LM loops
        import oracle_api as _O          # CPU baseline only (bench.py cpu_baseline leg)
        pbs = [_O.poseopt_problem(q["p3d"], q["obs"], q["inv_sigma2"], q["K"], q["Rcw"], q["tcw"]) for q in pp[:64]]
        t0 = time.perf_counter()
        for pb in pbs:
            _O.pose_optimization(pb)
        dtc = (time.perf_counter() - t0) / len(pbs)
        ex["pose_optimization"] = {"frames": CP, "edges_per_frame": NP, "ms_per_batch": msp, "frames_per_s": CP / (msp * 1e-3),
                                   "e2e_ms_per_batch": dte * 1e3, "e2e_frames_per_s": CP / dte,
                                   "lm_iterations_mean": float(resp["iterations"].mean()), "lm_trials_mean": float(resp["trials"].mean()),
                                   "edge_visits_per_s": passes / (msp * 1e-3), "inliers_mean": float(resp["n_inliers"].mean()),
                                   "cpu_port_single_thread_frames_per_s": 1.0 / dtc,
                                   "note": "Optimizer.cpp:205-424 batched; e2e = upload from host buffers + kernel + records and flags back"}
    except Exception as err:
        ex["pose_optimization_error"] = repr(err)
    # ---- the relocalisation chain on the device: cfg4 sweep (early exit) -> PoseOptimization of every verified candidate's
    # inliers from its RANSAC pose (rsac_poseopt_from_pnp: nothing crosses PCIe between the two)
    try:
        bc = synth.pnp_batch(4, 1024, 500, 0.5)
        offc = (np.arange(1025) * 500).astype(np.int32)
        prmc = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
        eng.pnp_upload(offc, bc["p3d"], bc["p2d"], bc["sigma2"], [bc["K"]], prmc, seeds=bc["seeds"])
        def chain():
            eng.pnp_run(capi.FLAG_EARLY_EXIT)
            eng.poseopt_from_pnp(0.0)
            eng.poseopt_run()
        for _ in range(3):
            chain()
        eng.sync()
        eng.timer_begin()
        for _ in range(10):
            chain()
        msc = eng.timer_end() / 10
        eng.timer_begin()
        for _ in range(10):
            eng.pnp_run(capi.FLAG_EARLY_EXIT)
        msr = eng.timer_end() / 10
        presc, _ = eng.poseopt_download()
        ex["relocalisation_chain"] = {"candidates": 1024, "matches": 500, "ms_per_sweep_ransac_only": msr, "ms_per_sweep_with_pose_optimization": msc,
                                      "candidates_per_s": 1024 / (msc * 1e-3), "frames_optimised": int((presc["rounds"] > 0).sum()),
                                      "inliers_after_optimization_mean": float(presc["n_inliers"].mean()),
                                      "note": "one engine, one sweep at a time (no sweeps in flight), inputs resident"}
    except Exception as err:
        ex["relocalisation_chain_error"] = repr(err)
    # ---- Optimizer::OptimizeSim3 for 256 loop candidates x 100 matches (fixed scale, th2 = 10)
    try:
        CS, NS = 256, 100
        sp = [synth.sim3opt_problem(8000 + i, NS, 0.15) for i in range(CS)]
        offs = (np.arange(CS + 1) * NS).astype(np.int32)
        cats = lambda k: np.concatenate([q[k] for q in sp])
        Ks = np.stack([q["K"] for q in sp])
        argss = (offs, cats("x1c"), cats("x2c"), cats("obs1"), cats("obs2"), cats("inv_sigma2_1"), cats("inv_sigma2_2"), Ks, Ks,
                 np.stack([q["S12"] for q in sp]), 10.0)
        ress, _ = eng.sim3opt_solve(*argss)
        for _ in range(3):
            eng.sim3opt_run()
        eng.sync()
        eng.timer_begin()
        for _ in range(20):
            eng.sim3opt_run()
        mss = eng.timer_end() / 20
        import oracle_api as _O          # CPU baseline only
        pbs = [_O.sim3opt_problem(q["x1c"], q["x2c"], q["obs1"], q["obs2"], q["inv_sigma2_1"], q["inv_sigma2_2"], q["K"], q["K"], q["S12"])
               for q in sp[:32]]
        t0 = time.perf_counter()
        for pb in pbs:
            _O.optimize_sim3(pb)
        dts = (time.perf_counter() - t0) / len(pbs)
        ex["optimize_sim3"] = {"pairs": CS, "matches_per_pair": NS, "ms_per_batch": mss, "pairs_per_s": CS / (mss * 1e-3),
                               "lm_iterations_mean": float(ress["iterations"].mean()), "lm_trials_mean": float(ress["trials"].mean()),
                               "inliers_mean": float(ress["n_inliers"].mean()), "cpu_port_single_thread_pairs_per_s": 1.0 / dts,
                               "note": "Optimizer.cpp:1054-1249 batched, numeric Jacobians (30 projections per match and build pass)"}
    except Exception as err:
        ex["optimize_sim3_error"] = repr(err)
    return out


def _quiet_main():
    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    out = sys.stdout
    sys.stdout = os.fdopen(saved, "w")
    try:
        return main()
    finally:
        sys.stdout.flush()
        sys.stdout = out


if __name__ == "__main__":
    sys.exit(_quiet_main())
