#!/usr/bin/env python
"""bench.py -- relocalisation-sweep throughput of the B200 RANSAC engine (BASELINE.json metric).

Workload (config.workload = "cfg4"): 1024 candidate keyframes x 500 2D-3D matches, 50 % outliers,
PnPsolver EPnP RANSAC with SetRansacParameters(0.99,10,300,4,0.2,5.991) => H = 300 hypotheses per
candidate.  The reference stops at the first hypothesis whose Refine() succeeds; the device does the same
in phases (RSAC_FLAG_EARLY_EXIT: the first 55 hypotheses of every candidate, the rest only for the candidates
that still need them), then replays the reference's sequential semantics (PnPsolver::iterate / Refine) per
candidate -- records and masks are identical to solving and scoring all 300 (RSAC_BENCH_EXHAUSTIVE=1 times
that mode: 153.6 M evaluations + 307 200 minimal solves per sweep).  With N GPUs every GPU works on its own
1024-candidate sweep (weak scaling; RSAC_BENCH_STRONG=1 shards one sweep instead) and the per-candidate
records (96 B) are all-gathered over NCCL every sweep.

A "step" is one sweep.  Sweeps are independent, so they are pipelined over a few engine
instances / CUDA streams; the timed region is K sweeps between barriers, timed with CUDA events,
max over ranks.
  value : candidates/s, inputs resident in HBM
  e2e   : the same through the host-buffer C-ABI call sequence (H2D of every sweep's inputs from
          pinned memory, D2H of results + inlier masks inside the timed region)
  roofline      : dominant kernel of the sweep (EPnP minimal solver, FP64 CUDA cores)
  roofline_score: CheckInliers kernel on cfg5 (4096 poses x 10 000 correspondences, FP32 CUDA cores)
  cpu_baseline  : the CPU oracle (port of the reference, see oracle/) on the host cores

--impl reference runs the reference arm: the oracle port of PnPsolver on all host threads, same
workload, same metric (the reference itself cannot be compiled here: it needs Eigen/OpenCV).
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

C_TOTAL = 1024
N_MATCH = 500
PRM = dict(prob=0.99, min_inliers=10, max_its=300, min_set=4, eps=0.2, th2=5.991)
H_HYP = 300
METRIC = "relocalization candidates/s (PnP EPnP RANSAC sweep; hyp x corr evals/s in extras)"
FLOP_PER_EVAL = 31            # PnP CheckInliers
PIPE = int(os.environ.get("RSAC_BENCH_PIPE", "0"))   # sweeps in flight per GPU (0: default, one)
EXHAUSTIVE = os.environ.get("RSAC_BENCH_EXHAUSTIVE", "0") == "1"   # solve and score all 300 hypotheses of every candidate


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


def make_shard(first, count):
    from ransac_b200 import synth
    b = synth.pnp_batch(4, count, N_MATCH, 0.5, first=first)
    offsets = (np.arange(count + 1, dtype=np.int64) * N_MATCH).astype(np.int32)
    return b, offsets


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU implementation of the path (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle_api as O
    O.build()
    cores = os.cpu_count() or 1
    strong = os.environ.get("RSAC_BENCH_STRONG", "0") == "1"
    n_cand = C_TOTAL if strong else C_TOTAL * max(1, args.gpus)     # the GPU arm's config at this N
    # the GPU arm's data: NBLOCK blocks of 1024 candidates; step k works on blocks (k + r) mod NBLOCK, r < gpus
    nblock = 1 if strong else max(1, int(os.environ.get("RSAC_BENCH_BLOCKS", "8")))
    prm = O.params(**PRM)
    oflags = O.FLAG_EPNP_QR_NULLSPACE   # the port's faster mode (same arithmetic as the device path)
    blocks = []
    for j in range(nblock):
        bj, _ = make_shard(j * C_TOTAL, C_TOTAL)
        blocks.append(([O.pnp_problem(bj["p3d"][c], bj["p2d"][c], bj["sigma2"][c], bj["K"]) for c in range(C_TOTAL)],
                       [O.index_table(int(sd), N_MATCH, 4, H_HYP) for sd in bj["seeds"]]))

    def step_work(k):
        pbs, tabs = [], []
        for r in range(1 if strong else max(1, args.gpus)):
            p_, t_ = blocks[(k + r) % nblock]
            pbs += p_
            tabs += t_
        return pbs, tabs

    for _ in range(max(1, min(args.warmup, 1))):
        O.pnp_batch(blocks[0][0][:64], prm, blocks[0][1][:64], oflags, cores)
    t_tot, ev_tot = 0.0, 0
    for k in range(args.steps):
        pbs, tables = step_work(k)
        dt, ev, res = O.pnp_batch(pbs, prm, tables, oflags, cores)
        t_tot += dt
        ev_tot += ev
    val = n_cand * args.steps / t_tot
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "candidates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": "cfg4", "candidates": n_cand, "matches": N_MATCH,
                                            "hypotheses": H_HYP, "mode": "reference semantics (early exit)",
                                            "blocks": f"{nblock} blocks of {C_TOTAL} synthetic candidates; step k works on blocks (k + r) mod {nblock}, r < {max(1, args.gpus)}"},
            "cpu_baseline": {"value": val, "unit": "candidates/s", "cores": cores, "kind": "port",
                             "sample": "full cfg4 sweep per step, one solver call per core (BASELINE.md mode B); "
                                       "4-point null space by Householder QR as on the device (the port's faster mode)",
                             "evals_per_s": ev_tot / t_tot},
            "e2e": {"value": val, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=8)
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
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    # Sweeps are independent, so several are in flight per GPU (one engine + stream each): every phase of a sweep is
    # one latency-bound wave, and a rare candidate that needs several Refine() calls, or the clean-up phase, stretches
    # a whole sweep (0.52 -> 1.16 ms for one such candidate) -- overlapping sweeps fills those holes
    # (measured on the 8-block mix, resident, two stages: 1 in flight 0.72 ms per sweep, 2: 0.59, 3: 0.55; three stages: 4 in
    # flight 0.48, 6: 0.45, 8: 0.46)
    global PIPE, C_TOTAL
    if PIPE <= 0:
        PIPE = 6
    # weak scaling (task statement, section 5): the path shards by candidate with no data-path collective, so
    # every GPU works on its own 1024-candidate sweep and the job processes 1024 x N candidates per step;
    # RSAC_BENCH_STRONG=1 keeps the total at 1024 instead (each GPU then gets 1024/N candidates)
    strong = os.environ.get("RSAC_BENCH_STRONG", "0") == "1"
    C_PER_GPU = C_TOTAL
    if not strong:
        C_TOTAL = C_PER_GPU * world
    RUN_FLAGS = 0 if EXHAUSTIVE else capi.FLAG_EARLY_EXIT
    first, count = shard.block_range(C_TOTAL, rank, world)
    cap = shard.per_rank_capacity(C_TOTAL, world)
    # The synthetic set is NBLOCK blocks of 1024 candidates; at step k rank r works on block (k + r) mod NBLOCK, so
    # every rank meets every block equally often at every N (blocks differ: some hold a candidate whose refines
    # fail, which costs the clean-up phase) and the ranks of one step work on different blocks.
    NBLOCK = 1 if strong else max(1, int(os.environ.get("RSAC_BENCH_BLOCKS", "8")))
    prm = capi.ransac_params(**PRM)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    blocks = []
    for j in range(NBLOCK):
        bj = (j + rank) % NBLOCK
        bb, offsets = make_shard(first if strong else bj * C_PER_GPU, count)
        blocks.append(dict(p3d=pin(bb["p3d"].reshape(-1, 3)), p2d=pin(bb["p2d"].reshape(-1, 2)), s2=pin(bb["sigma2"].reshape(-1)),
                           seeds=bb["seeds"], K=bb["K"], block=bj))
    b = dict(K=blocks[0]["K"], p3d=blocks[0]["p3d"].numpy().reshape(count, N_MATCH, 3), p2d=blocks[0]["p2d"].numpy().reshape(count, N_MATCH, 2),
             sigma2=blocks[0]["s2"].numpy().reshape(count, N_MATCH), seeds=blocks[0]["seeds"])

    NRES = max(NBLOCK, PIPE)             # resident pass: one engine per block (block j of this rank stays uploaded in engine j)
    NSLOT = PIPE + 1                     # end-to-end pass: engines 0 .. PIPE are re-uploaded every step
    NENG = max(NRES, NSLOT)
    engines, streams, d_local = [], [], []
    for i in range(NENG):
        e = capi.Engine(local_rank)
        s = torch.cuda.Stream(device=dev)
        e.set_stream(s.cuda_stream)
        e.set_problem_base(first)
        engines.append(e)
        streams.append(s)
        d_local.append(torch.full((cap, shard.REC_WORDS), -1, dtype=torch.int32, device=dev))
    words_total = int(((np.diff(offsets) + 31) // 32).sum())
    h_res = [torch.empty((count, shard.REC_WORDS), dtype=torch.int32).pin_memory() for _ in range(NSLOT)]
    h_msk = [torch.empty((max(words_total, 1),), dtype=torch.int32).pin_memory() for _ in range(NSLOT)]

    def upload(i, j):
        blk = blocks[j % NBLOCK]
        engines[i].pnp_upload(offsets, blk["p3d"].numpy(), blk["p2d"].numpy(), blk["s2"].numpy(), [blk["K"]], prm, seeds=blk["seeds"])

    # multi-GPU: the all-gather of a sweep's records (98 KB per rank, latency-bound) runs on a side stream, in
    # issue order; a ring of gather buffers, and every engine waits for the gather that last read its records
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    NG = NENG + 2
    d_gath = [torch.empty((world * cap, shard.REC_WORDS), dtype=torch.int32, device=dev) for _ in range(NG)] if world > 1 else None
    gather_done = [None] * NENG
    slot_done = [None] * NENG          # completion of the last sweep that ran on engine i
    last_gather = [None]
    GATHER = os.environ.get("RSAC_BENCH_GATHER", "overlap")   # overlap | none (diagnostic)

    def run_and_gather(i, k):
        st = streams[i]
        if gather_done[i] is not None:
            st.wait_event(gather_done[i])              # the gather that last read this engine's records
        engines[i].pnp_run(RUN_FLAGS, d_local[i].data_ptr())
        if world > 1 and GATHER != "none":
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
        # at most PIPE sweeps in flight: sweep k starts after sweep k - PIPE has finished
        prev = ring[(k - PIPE) % len(ring)] if k >= PIPE else None
        if prev is not None:
            streams[i].wait_event(prev)

    res_ring = [None] * (NRES + PIPE)
    e2e_ring = [None] * (NSLOT + PIPE)

    def step_resident(k):
        i = k % NRES
        with torch.cuda.stream(streams[i]):
            bound_in_flight(i, k, res_ring)
            run_and_gather(i, k)
            ev = torch.cuda.Event()
            ev.record(streams[i])
            res_ring[k % len(res_ring)] = ev

    def step_e2e(k):
        i = k % NSLOT
        with torch.cuda.stream(streams[i]):
            upload(i, k)                                # H2D of block (k + rank) mod NBLOCK, plans, tables
            bound_in_flight(i, k, e2e_ring)
            run_and_gather(i, k)
            ev = torch.cuda.Event()
            ev.record(streams[i])
            e2e_ring[k % len(e2e_ring)] = ev
            # D2H of this sweep's records and inlier masks into pinned memory (async on the sweep's stream)
            engines[i].pnp_download_async(h_res[i].data_ptr(), h_msk[i].data_ptr())

    def timed(fn, steps):
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
        for k in range(steps):
            fn(k)
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
    timed(step_resident, max(args.warmup, NRES))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sum(e.launch_count() for e in engines)
    ms = timed(step_resident, args.steps)
    launches = sum(e.launch_count() for e in engines) - launches0
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel durations: CUDA events around every launch of a NON-pipelined pass (one sweep at a time on one
    # stream), so that an event pair brackets exactly one kernel; in the pipelined region above kernels of
    # different sweeps overlap and a bracket would also count the neighbours
    stage = {}
    e0 = engines[0]
    e0.profile_reset()
    e0.profile_enable(True)
    nprof = max(3, min(10, args.steps))
    with torch.cuda.stream(streams[0]):
        for _ in range(nprof):
            e0.pnp_run(RUN_FLAGS, d_local[0].data_ptr())
    torch.cuda.synchronize()
    for k, (tms, nl) in e0.profile().items():
        stage[k] = [tms, nl]
    trace = e0.profile_trace()
    trace = trace[-(len(trace) // nprof):] if trace else []      # the launches of the last sweep, in order
    e0.profile_enable(False)
    # phases and hypotheses actually solved and scored: mean over the blocks this rank holds
    stats = [engines[i].pnp_phase_stats() for i in range(NRES)]
    ha = stats[0][0]
    n_b = float(np.mean([st[1] for st in stats]))
    n_c = float(np.mean([st[2] for st in stats]))
    hyp_done = float(np.mean([st[3] for st in stats]))

    # correctness guard on the gathered records (cheap): every candidate reported once, in order
    torch.cuda.synchronize()
    if world > 1 and GATHER == "none":
        n_ok = -1
    else:
        rec = shard.records_from_tensor(last_gather[0])
        assert len(rec) == C_TOTAL and (rec["problem"] == np.arange(C_TOTAL)).all(), "gather lost candidates"
        n_ok = int(rec["ok"].sum())

    # end-to-end through host buffers
    timed(step_e2e, max(4, 2 * NSLOT))
    ms_e2e = timed(step_e2e, args.steps)
    h2d = int(count * N_MATCH * 24 + count * 4 + count * 152)
    d2h = int(count * 96 + words_total * 4)

    value = C_TOTAL * args.steps / (ms * 1e-3)
    e2e_v = C_TOTAL * args.steps / (ms_e2e * 1e-3)
    hyp_frac = hyp_done / float(count * H_HYP)                   # share of the 300 x candidates hypotheses computed

    line = {"metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64 solve / f32 score", "data": "synthetic",
            "config": {"workload": "cfg4", "candidates": C_TOTAL, "candidates_per_gpu": count, "matches": N_MATCH,
                       "hypotheses": H_HYP, "outliers": 0.5,
                       "mode": ("all H hypotheses solved and scored on the device (4-point null space by QR), then "
                                "reference-semantics replay + Refine per candidate") if EXHAUSTIVE else
                               (f"reference semantics with early exit in phases: hypotheses [0,{ha}) of every candidate, the "
                                f"remaining ones for the {n_b:.0f} of {count} candidates (mean over blocks) without an acceptable "
                                f"hypothesis so far ({100 * hyp_frac:.1f} % of the 300 x {count} hypotheses solved and scored), "
                                "replay + Refine per candidate; records identical to the exhaustive run"),
                       "blocks": f"{NBLOCK} blocks of {C_PER_GPU} synthetic candidates; at step k rank r works on block (k + r) mod {NBLOCK}",
                       "parallelism": f"one 1024-candidate sweep per GPU and step (x{world} GPUs), {PIPE} independent sweeps in flight per "
                                      "GPU (one engine + stream each); e2e uploads every step's block from pinned host memory and "
                                      "reads records + masks back inside the timed region",
                       "l2": (f"no explicit flush: {NBLOCK} blocks are cycled, each in its own engine (~90 MB of device buffers per block, "
                              f"{NBLOCK * 90} MB in total against 126 MB of L2), so a step's data was last touched {NBLOCK} steps earlier; "
                              "within a sweep the working set is L2-resident by design and every kernel is compute- or latency-bound")},
            "e2e": {"value": e2e_v, "unit": "candidates/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "extras": {"evals_per_s": value * H_HYP * N_MATCH * hyp_frac, "e2e_evals_per_s": e2e_v * H_HYP * N_MATCH * hyp_frac,
                       "evals_note": "hypothesis x correspondence evaluations actually performed (early exit skips the rest)",
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
        flop_per_solve = epnp_flops_per_solve(b)
        # per sweep: with early exit a kernel is launched once per phase (A, B, clean-up); work = what the launches
        # of a sweep actually did, time = their summed durations, so achieved = work / time is the launch-weighted
        # average; the first launch of each kind (phase A: count x first_phase hypotheses) is also given alone
        per = {k: v[0] / nprof for k, v in stage.items() if v[1]}
        nl = {k: v[1] / nprof for k, v in stage.items() if v[1]}
        total_ms = sum(per.values()) or 1.0
        first = {}
        for k, m in trace:
            first.setdefault(k, m)
        roofs = {}
        if per.get("solve"):
            t = per["solve"] * 1e-3
            ach = flop_per_solve * hyp_done / t / 1e12
            roofs["solve"] = {"bound": "fp64", "kernel": "epnp_minimal_kernel<QR>" if EXHAUSTIVE else "epnp_minimal_range_kernel",
                              "achieved": ach, "peak": fp64_pk,
                              "unit": "TFLOP/s", "frac": ach / fp64_pk,
                              "traffic": ncu_traffic("epnp_minimal_kernel<QR>" if EXHAUSTIVE else "epnp_minimal_range_kernel") if count == 1024 else None,
                              "peak_source": "DFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                              "flop_per_solve": flop_per_solve, "solves_per_sweep": hyp_done, "launch_ms": per["solve"],
                              "launches_per_sweep": nl["solve"], "share_of_sweep": per["solve"] / total_ms}
            if not EXHAUSTIVE and first.get("solve"):
                # the stage-A launch (every candidate's first hypotheses: one full resident wave) is the dominant launch
                # of the sweep; the later stages launch the same kernel on a fraction of a wave (latency-bound by design,
                # they overlap with other sweeps).  `roofline` quotes the stage-A launch; the launch-weighted figure over
                # all launches of a sweep stays beside it
                a1 = flop_per_solve * count * ha / (first["solve"] * 1e-3) / 1e12
                agg = dict(achieved=roofs["solve"]["achieved"], frac=roofs["solve"]["frac"], launch_ms=roofs["solve"]["launch_ms"],
                           solves=hyp_done, launches_per_sweep=nl["solve"])
                roofs["solve"].update({"achieved": a1, "frac": a1 / fp64_pk, "launch_ms": first["solve"], "solves_per_launch": count * ha,
                                       "launch": "stage A (hypotheses [0,%d) of %d candidates)" % (ha, count),
                                       "all_launches_of_a_sweep": agg, "share_of_sweep": first["solve"] / total_ms})
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
        dom = max(roofs, key=lambda k: roofs[k]["launch_ms"] if "achieved" in roofs[k] else 0.0) if roofs else None
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
    """per-launch DRAM bytes of a kernel from the committed ncu capture (profiles/r01_traffic.json), or None"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        for k, v in d.items():
            if k.startswith(key) and "phase B" not in k:
                return int(v["dram_bytes"])
    except Exception:
        pass
    return None


def epnp_flops_per_solve(b):
    """algorithmic FP64 FLOP of one 4-point EPnP solve, from the oracle's operation counters"""
    try:
        import oracle_api as O
        O.build()
        pb = O.pnp_problem(b["p3d"][0], b["p2d"][0], b["sigma2"][0], b["K"])
        tab = O.index_table(int(b["seeds"][0]), N_MATCH, 4, 64)
        return float(O.epnp_flops(pb, tab, O.FLAG_EPNP_QR_NULLSPACE))
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
                                        "evals_per_s": C3 * H3 * 200 / (ms3 * 1e-3)}
    except Exception as err:   # the headline line must still be printed
        ex["cfg2_cfg3_error"] = repr(err)
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
