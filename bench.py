#!/usr/bin/env python
"""bench.py — headline benchmark of the estimate_transform hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this implementation (CUDA, C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU port of the reference's path (oracle/)

Workload (config.workload): BASELINE.json configs[1] — synthetic 8k x 8k-nucleus embryo-like pair
(N2 = 8000 fixed, N1 = 7200 moving after 10 % dropout, jitter 2 px), unsupervised
estimate_transform: shape context -> 4 chi^2 cost matrices -> 4 assignments -> 4 affine RANSACs
(8000 trials each, the widget default) -> ICP (50 iterations).
A step = one full registration of one pair.  metric = registrations/sec (whole job, all ranks).
  value   inputs already resident in HBM when the timed region starts
  e2e     the public Python API with HOST (pinned) inputs: H2D of both clouds and D2H of the results
          inside the timed region
Multi-GPU (torchrun, one rank per GPU): independent specimen pairs are sharded across ranks (no
data-path collective) -> weak scaling; timing = max over ranks of the device time.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "synthetic 8k x 8k-nucleus pair (N1=7200 moving, N2=8000 fixed), unsupervised estimate_transform"
N_FIXED = 8000
TRIALS = 8000
ICP_ITERS = 50
FLOP_PER_PAIR = 1801.0        # 5 FLOP per bin pair x 360 + 1
METRIC = "registrations/sec"
CHI2_NCU_DRAM_BYTES = 186.0e6     # measured once with ncu at the headline size (10.6 MB read + 175.4 MB written)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-fixed", type=int, default=N_FIXED, help="override the workload size (not the headline)")
    ap.add_argument("--trials", type=int, default=TRIALS)
    ap.add_argument("--bid-rounds", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stages", action="store_true", help="print a per-stage device-time breakdown to stderr")
    ap.add_argument("--in-flight", type=int, default=3,
                    help="registrations in flight per GPU for the secondary `pipelined` throughput number (0 = skip)")
    ap.add_argument("--no-rows", action="store_true", help="skip the secondary measurements of the widened rows")
    ap.add_argument("--lean", action="store_true",
                    help="profiling aid: warm-up + timed resident steps only (no e2e, stage, roofline or CPU legs)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(pair, trials, icp_iters, threads=None):
    """Bounded sample of the oracle (CPU port of the reference's path) on the same workload, extrapolated
    to one full registration.  Sampled loops are embarrassingly parallel over the sampled index, so the
    extrapolation is a plain ratio; the assignment (superlinear) is solved in full for the true hypothesis."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    if threads:
        O.set_num_threads(threads)
    m, f = pair["moving"], pair["fixed"]
    n1, n2 = m.shape[1], f.shape[1]
    FR = 8                                   # sample 1/8 of the query nuclei / cost rows / trials
    t = {}
    t0 = time.perf_counter()
    mc, fc = O.get_centroid(m, False), O.get_centroid(f, False)
    md, fd = O.get_mean_distance(m, False), O.get_mean_distance(f, False)
    t["mean_distance"] = time.perf_counter() - t0
    xm, xf = O.pca_first_axis(m.T), O.pca_first_axis(f.T)
    t0 = time.perf_counter()
    O.shape_context_counts(m.T, mc, md, xm, 1, query_range=(0, n1 // FR))
    O.shape_context_counts(f.T, fc, fd, xf, 2, query_range=(0, n2 // FR))
    ts = time.perf_counter() - t0
    # reference builds 2 moving + 4 fixed descriptor sets (shape_context.py:170-185)
    t["descriptors"] = ts * FR * (2 * n1 + 4 * n2) / (n1 + n2)
    # full descriptors of the true hypothesis (needed for a real cost matrix for the LAP); not timed
    um = O.normalise_counts(O.shape_context_counts(m.T, mc, md, xm, 1)[0])
    uf = O.normalise_counts(O.shape_context_counts(f.T, fc, fd, xf, 2)[0])
    t0 = time.perf_counter()
    O.unary_distance_matrix(um[: n1 // FR], uf)
    t["chi2_per_matrix"] = (time.perf_counter() - t0) * FR
    U = O.unary_distance_matrix(um, uf)
    t0 = time.perf_counter()
    r, c = O.linear_sum_assignment(U)
    t["lap_per_matrix"] = time.perf_counter() - t0
    idx = O.ransac_sample_indices(len(r), 4, max(trials // FR, 1), seed=0)
    t0 = time.perf_counter()
    A, inl = O.do_ransac(m[:, r], f[:, c], 4, len(idx), 16, sample_indices=idx)
    t["ransac_per_matrix"] = (time.perf_counter() - t0) * trials / len(idx)
    it = max(icp_iters // 10, 1)
    t0 = time.perf_counter()
    O.perform_icp(O.apply_affine_transform(m, A), f, it)
    t["icp"] = (time.perf_counter() - t0) * icp_iters / it
    # the reference evaluates 8 hypotheses; 4 are algebraically distinct (what the GPU arm computes)
    H = 4
    total = t["mean_distance"] + t["descriptors"] + H * (t["chi2_per_matrix"] + t["lap_per_matrix"] +
                                                         t["ransac_per_matrix"]) + t["icp"]
    sample = ("oracle (C/numpy port), %d threads: full mean-distance; 1/%d of query nuclei for descriptors; 1/%d of "
              "rows of one chi2 matrix; full LAP of the true hypothesis; 1/%d of %d RANSAC trials; %d of %d ICP "
              "iterations; scaled to 6 descriptor sets + 4 hypotheses" % (O.num_threads(), FR, FR, FR, trials, it, icp_iters))
    return dict(seconds_per_registration=total, stages=t, cores=O.num_threads(), sample=sample,
                gpairs_per_s=n1 * n2 / t["chi2_per_matrix"] / 1e9)


def run_reference(args):
    """--impl reference: the reference's CPU path.  The reference is pure Python and does not travel to
    the GPU box (and would need ~27 h per 8k registration), so this arm times the oracle port of it
    (oracle/, pinned to the reference by tests/golden) with all host threads, one bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from platymatch_b200.synthetic import make_pair
    pair = make_pair(args.n_fixed)
    times = []
    last = None
    for s in range(args.warmup + args.steps):
        last = cpu_sample(pair, args.trials, ICP_ITERS)
        if s >= args.warmup:
            times.append(last["seconds_per_registration"])
    sec = float(np.mean(times))
    val = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "registrations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_fixed": args.n_fixed, "ransac_trials": args.trials,
                       "icp_iterations": ICP_ITERS, "hypotheses": 4},
            "cpu_baseline": {"value": val, "unit": "registrations/s", "cores": last["cores"], "kind": "port",
                             "sample": last["sample"], "stages_s": last["stages"],
                             "cost_matrix_gpairs_per_s": last["gpairs_per_s"]},
            "e2e": {"value": val, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- widened rows
def bench_label_row(torch, D, hbm_peak, with_cpu):
    """label volume -> detections (reference _dock_widget.py:497-521).  HBM-bound streaming
    kernel: algorithmic bytes = one read of the volume; roofline against the measured copy bandwidth."""
    from platymatch_b200.synthetic import make_label_volume
    shape, n_nuclei = (384, 512, 512), 6000
    rng = np.random.default_rng(1)
    d = rng.normal(size=(n_nuclei, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    centers = np.array(shape) / 2.0 + d * (np.array(shape) * 0.42) + rng.normal(0, 4.0, size=(n_nuclei, 3))
    vol = make_label_volume(shape, radius=(3.0, 5.0), seed=2, centers=centers, dtype=np.int32)
    dev = torch.from_numpy(vol).cuda()
    nbytes = vol.size * 4
    for _ in range(3):
        ids, cen, sizes = D.label_centroids(dev, 1.0, table_size=n_nuclei + 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    ms = 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0.record()
        D.label_centroids(dev, 1.0, table_size=n_nuclei + 1, sync=False)          # memset + accumulate + compact
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1) / reps
    out = {"workload": "label volume %dx%dx%d int32, %d nuclei (%.0f %% foreground)" %
                       (shape + (int(ids.numel()), 100.0 * float((vol > 0).mean()))),
           "ms": ms, "voxels_per_s": vol.size / (ms * 1e-3),
           "roofline": {"kernel": "pm_label_accumulate_kernel", "bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak, "of": "measured",
                        "algorithmic_bytes": nbytes, "l2": "160 MB flush between repetitions"}}
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        k = 4                                          # the reference makes one np.where pass over the volume per id
        sub = vol.copy()
        keep = np.isin(sub, np.unique(sub[sub > 0])[:k])
        sub[~keep] = 0
        t0 = time.perf_counter()
        O.detections_from_labels(sub, 1.0)
        t = (time.perf_counter() - t0) / k * int(ids.numel())
        out["cpu_baseline"] = {"value": vol.size / t, "unit": "voxels/s", "cores": 1, "kind": "port",
                               "sample": "oracle restatement of the per-id np.where loop: %d of %d ids timed, scaled" % (k, int(ids.numel()))}
    return out


# ------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for k, nme in enumerate(names):
                    if p[5 + k].lower().startswith("active"):
                        reasons.add(nme)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    import platymatch_b200 as pm
    from platymatch_b200 import device as D, pipeline as P
    from platymatch_b200._lib import load
    from platymatch_b200.synthetic import make_pair

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load()
    bid_rounds = args.bid_rounds if args.bid_rounds is not None else 2048
    kw = dict(ransac_trials=args.trials, icp_iterations=ICP_ITERS, max_bid_rounds=bid_rounds)

    # 4 specimen pairs cycled over the steps; every rank registers the same sequence of pairs, so the per-GPU
    # work is identical at every N and for every K, and the N-GPU value measures scaling, not pair difficulty
    # (the assignment stage of the third pair takes twice as long as that of the others)
    n_slots = 4
    pairs = [make_pair(args.n_fixed, seed=args.n_fixed + 97 * s) for s in range(n_slots)]
    n1, n2 = pairs[0]["moving"].shape[1], pairs[0]["fixed"].shape[1]
    dev_pairs = [(D.to_device_points(p["moving"]), D.to_device_points(p["fixed"])) for p in pairs]
    host_pairs = [(torch.from_numpy(np.ascontiguousarray(p["moving"].T)).pin_memory(),
                   torch.from_numpy(np.ascontiguousarray(p["fixed"].T)).pin_memory()) for p in pairs]
    cost_buf = torch.empty((4, n1, (n2 + 3) // 4 * 4), dtype=torch.float32, device="cuda")

    def step_resident(s, hook=None):
        m, f = dev_pairs[s % n_slots]
        dm = P.describe_cloud(m, 1, transposed=True)
        if hook: hook("describe_moving")
        df = P.describe_cloud(f, 4, transposed=True)
        if hook: hook("describe_fixed")
        return P.register_described(dm, df, seed=s, cost_out=cost_buf, stage_hook=hook, **kw)

    def step_e2e(s):
        hm, hf = host_pairs[s % n_slots]
        m = hm.to("cuda", non_blocking=True)
        f = hf.to("cuda", non_blocking=True)
        dm = P.describe_cloud(m, 1, transposed=True)
        df = P.describe_cloud(f, 4, transposed=True)
        res = P.register_described(dm, df, seed=s, cost_out=cost_buf, **kw)
        out = torch.cat([res["transform"], res["transform_sc"], res["transform_icp"],
                         res["inliers"].to(torch.float64)]).cpu()       # D2H of the step's result
        return res, out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for s in range(args.warmup):
        step_resident(s)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = lib.pm_launch_count()
    ms = timed(step_resident, args.steps)
    launches = lib.pm_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    if args.lean:
        if rank == 0:
            print(json.dumps({"lean": True, "ms_per_step": ms / args.steps, "gpu_launches": int(launches)}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    for s in range(min(args.warmup, 2)):
        step_e2e(s)
    ms_e2e = timed(step_e2e, args.steps)

    # ---- per-stage breakdown + roofline of the chi2 kernel (CUDA events on the launching stream) ----
    marks = []

    def hook(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    stage_ms = {}
    reps = max(2, min(args.steps, 5))
    step_resident(0, lambda name: None)      # untimed: the first batched (non-overlapped) pass pays one-time allocations
    torch.cuda.synchronize()
    for s in range(reps):
        marks.clear()
        hook("start")
        res = step_resident(s, hook)
        torch.cuda.synchronize()
        for (n0, a), (n1_, b) in zip(marks[:-1], marks[1:]):
            stage_ms[n1_] = stage_ms.get(n1_, 0.0) + a.elapsed_time(b) / reps
            if args.stages and n1_ == "lap":
                print("rep %d (pair slot %d): lap %.2f ms" % (s, s % n_slots, a.elapsed_time(b)), file=sys.stderr)
    lap_stats = res["lap_stats"].cpu().numpy().tolist()

    # chi2 kernel alone: algorithmic FLOPs per launch / average launch duration
    dm = P.describe_cloud(dev_pairs[0][0], 1, transposed=True)
    df = P.describe_cloud(dev_pairs[0][1], 4, transposed=True)
    a_t, b_t = dm.operand(1), df.operand(2)
    for _ in range(3):
        D.chi2_cost(a_t, b_t, out=cost_buf[0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nrep = 8
    e0.record()
    for q in range(nrep):
        D.chi2_cost(a_t, b_t, out=cost_buf[q % 4])      # 4 x 230 MB outputs cycled: > L2
    e1.record()
    torch.cuda.synchronize()
    chi2_ms = e0.elapsed_time(e1) / nrep
    chi2_tflops = FLOP_PER_PAIR * n1 * n2 / (chi2_ms * 1e-3) / 1e12
    chi2_gpairs = n1 * n2 / (chi2_ms * 1e-3) / 1e9
    # FP32 FMA peak measured on this GPU, now (burst, kernel alone)
    sink = torch.zeros(4, dtype=torch.float32, device="cuda")
    import ctypes
    flops = ctypes.c_double(0)
    sm = lib.pm_sm_count(local)
    for _ in range(2):
        lib.pm_probe_fp32_fma(sm * 8, 2000, D.ptr(sink), ctypes.byref(flops), D.stream_ptr())
    fp32_peak = 0.0
    for _ in range(3):                       # best of 3: the denominator is a peak, not an average
        e0.record()
        lib.pm_probe_fp32_fma(sm * 8, 20000, D.ptr(sink), ctypes.byref(flops), D.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        fp32_peak = max(fp32_peak, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    chi2_bytes = 4.0 * n1 * n2 + 4.0 * 360 * (n1 + n2)      # algorithmic: write the matrix, read both operands once

    # ---- secondary: independent registrations overlapped on separate streams (one host thread each) ----
    pipelined = None
    if args.in_flight > 1:
        import threading
        nfl = args.in_flight
        bufs = [cost_buf] + [torch.empty_like(cost_buf) for _ in range(nfl - 1)]
        streams = [torch.cuda.Stream() for _ in range(nfl)]

        def worker(k, steps):
            torch.cuda.set_device(local)
            with torch.cuda.stream(streams[k]):
                for s_ in range(k, steps, nfl):
                    m, f = dev_pairs[s_ % n_slots]
                    dmk = P.describe_cloud(m, 1, transposed=True)
                    dfk = P.describe_cloud(f, 4, transposed=True)
                    P.register_described(dmk, dfk, seed=s_, cost_out=bufs[k], overlap_hypotheses=False, **kw)

        def run_all(steps):
            ts = [threading.Thread(target=worker, args=(k, steps)) for k in range(nfl)]
            for t_ in ts: t_.start()
            for t_ in ts: t_.join()

        run_all(2 * nfl)                                   # warm-up
        barrier()
        t0 = time.perf_counter()
        run_all(args.steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        pipelined = {"value": world * args.steps / dt, "unit": "registrations/s", "in_flight_per_gpu": nfl,
                     "ms_per_registration": dt / args.steps * 1e3,
                     "note": "independent registrations overlapped on %d CUDA streams per GPU (host clock, inputs "
                             "resident); the headline `value` registers one pair at a time" % nfl}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total = world * args.steps
    value = total / (ms * 1e-3)
    e2e_val = total / (ms_e2e * 1e-3)
    h2d = (n1 + n2) * 3 * 8
    d2h = (16 * 3 + 4) * 8
    line = {
        "metric": METRIC, "value": value, "unit": "registrations/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 descriptors/duals/transforms, f32 cost matrix", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_fixed": n2, "n_moving": n1, "hypotheses": 4, "ransac_trials": args.trials,
                   "icp_iterations": ICP_ITERS, "lap_bid_rounds": bid_rounds, "parallelism": "pairs sharded, 1 rank/GPU",
                   "l2": "4 x %.0f MB cost matrices rewritten every step (> 126 MB L2); 4 input pairs cycled" %
                         (n1 * n2 * 4 / 1e6)},
        "e2e": {"value": e2e_val, "unit": "registrations/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "pipelined": pipelined,
        "cost_matrix_gpairs_per_s": chi2_gpairs * world,
        "roofline": {"kernel": "pm_chi2_kernel", "bound": "fp32", "achieved": chi2_tflops, "peak": fp32_peak,
                     "unit": "TFLOP/s", "frac": chi2_tflops / fp32_peak,
                     "peak_source": "FFMA probe kernel run in this process (MEASURED_PEAKS.json has no FP32 CUDA-core "
                                    "figure); nominal 2*128*148*1.965 GHz = 74.4",
                     "flop_per_pair": FLOP_PER_PAIR, "launch_ms": chi2_ms,
                     "traffic": CHI2_NCU_DRAM_BYTES if (n1, n2) == (7200, 8000) else None,
                     "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                       "(profiles/r1_chi2_kernel.txt); algorithmic bytes: %.0f" % chi2_bytes,
                     "hbm": {"achieved": chi2_bytes / (chi2_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": chi2_bytes / (chi2_ms * 1e-3) / 1e9 / hbm_peak, "of": "measured"}},
        "stages_ms": stage_ms,
        "lap_stats": {"bid_rounds": [s[0] for s in lap_stats], "rows_after_bidding": [s[1] for s in lap_stats],
                      "augmentations": [s[2] for s in lap_stats], "dijkstra_steps": [s[3] for s in lap_stats],
                      "bids": [s[5] for s in lap_stats], "refreshes": [s[6] for s in lap_stats],
                      "retries": [s[7] for s in lap_stats], "parked": [s[8] for s in lap_stats],
                      "refresh_cycles": [s[9] for s in lap_stats], "auction_cycles": [s[10] for s in lap_stats],
                      "bulk_bids": [s[11] for s in lap_stats], "sap_dense_relax": [s[12] for s in lap_stats]},
    }
    # the CPU baseline and the secondary rows are single-GPU measurements (rank 0 at N = 1 only)
    if not args.no_rows and world == 1:
        line["rows"] = {"label_centroids": bench_label_row(torch, D, hbm_peak, not args.no_cpu_baseline)}
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_sample(pairs[0], args.trials, ICP_ITERS)
        line["cpu_baseline"] = {"value": 1.0 / cb["seconds_per_registration"], "unit": "registrations/s",
                                "cores": cb["cores"], "kind": "port", "sample": cb["sample"], "stages_s": cb["stages"],
                                "cost_matrix_gpairs_per_s": cb["gpairs_per_s"]}
    if args.stages:
        print(json.dumps(stage_ms, indent=1), file=sys.stderr)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
