#!/usr/bin/env python
"""bench.py — MonoSLAM EKF hot path on B200: frames/s of predict + active-search match + update.

Workload (BASELINE.json configs[1]): single filter, 500 inverse-depth features (state dim 3014),
synthetic 640x480 sequence, fp64.  One "step" = one camera frame through captureNewFrame -> predict
-> update (match + RANSAC + both EKF corrections + book-keeping), exactly the call sequence of
ImageConverter::imageCb (monoslam_ransac.cpp:404-557).

  value  : frames/s with the frames already resident in HBM (ekf_capture_frame_device)
  e2e    : frames/s through the public API with HOST frames (H2D inside the timed step) and the
           state / covariance / feature flags read back to the host every step
  N > 1  : the single-filter path does not shard ("replicas only"): every rank runs an
           independent replica, value = total frames / max-over-ranks time, scaling "weak"
  --impl reference : the CPU oracle (dense reference algebra, all host threads) on the same
           workload, rank 0 only

Timing: W warm-up steps, then K steps each bracketed by CUDA events on the launching stream; L2 is
flushed (256 MiB write) between timed steps, outside the event pairs; max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "EKF frames/s (predict+match+update), single filter, N=500 features"


def metric_for(workload):
    n = WORKLOADS[workload][0]
    return f"EKF frames/s (predict+match+update), single filter, N={n} features"
UNIT = "frames/s"
WORKLOADS = {
    # name: (features, width, height)
    "cfg1_n50": (50, 640, 480),
    "cfg2_n500": (500, 640, 480),
    # BASELINE configs[2]: 4096 independent filters x 30 features per GPU, sharded by filter
    "cfg3_batch4096": (30, 640, 480),
    # BASELINE configs[3]: large map, 2000 features (n = 12014, Sigma = 1.15 GB), un-partitioned on one B200
    "cfg4_n2000": (2000, 1920, 1080),
    # BASELINE configs[4]: stateless active-search NCC matching, 256 frames x 200 features, 11x11 patches
    "cfg5_match": (200, 1920, 1080),
}
MATCH_FRAMES = 256
METRIC_MATCH = "active-search NCC matches/s (Patch::findMatch, 11x11 templates, 200 features per 1920x1080 frame, 256 frames)"
UNIT_MATCH = "matches/s"
BATCH_FILTERS = 4096
METRIC_BATCH = "batched EKF filter-steps/s (predict+match+update), 4096 independent filters x N=30 features per GPU"
UNIT_BATCH = "filter-steps/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_scene(pkg, workload, n_frames, seed=1235, visible_every=1):
    nfeat, W, H = WORKLOADS[workload]
    # slow motion so that all seeded features stay inside the image for the whole run (fixed N)
    return pkg.synth.Scene(n_features=nfeat, width=W, height=H, n_frames=n_frames, seed=seed, speed=0.1, omega=0.02,
                           accel_sigma=0.002, border=44, visible_every=visible_every)


def seed_filter(filt, scene):
    filt.captureNewFrame(scene.frame(0), scene.stamps[0])
    added = sum(filt.addFeature(*p) for p in scene.feature_pixels)
    return added


def dgemm_peak_tflops(torch, n=4096):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / best / 1e9


def run_ours(args):
    import torch
    import ekfb200
    pkg = ekfb200.load_package()
    pkg.lib()  # fails loudly if the CUDA library is missing
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, Wm = args.steps, max(args.warmup, 3)
    nfeat, width, height = WORKLOADS[args.workload]
    scene = make_scene(pkg, args.workload, 1 + Wm + K, seed=1235 + (0 if (world > 1 and args.workload.startswith("cfg4")) else rank),
                       visible_every=max(1, args.match_every))
    frames = [scene.frame(t) for t in range(scene.n_frames)]
    over = scene.config_overrides()
    if args.match_every > 1:   # SURVEY.md 8(d) "also report m = N/4": unmatched features must stay in the map (fixed n)
        over["quality_ratio"] = 1.0e9
    cfg = pkg.default_config(**over)
    stream = torch.cuda.current_stream()
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    # cfg4 (large map) at N > 1: ONE filter whose stacked update is partitioned by covariance row blocks
    # (every rank holds a replica and runs the same calls); other single-filter workloads: replicas only
    partitioned = world > 1 and args.workload.startswith("cfg4")

    def new_filter():
        f = pkg.VSlamFilter(cfg, feature_capacity=nfeat + 4, device=local)
        f.set_stream(stream.cuda_stream)
        f.set_symmetric_downdate(not args.full_square)
        added = seed_filter(f, scene)
        assert added == nfeat, f"seeded {added} of {nfeat}"
        if partitioned:
            pkg.dist.attach_row_partition(f, dist, torch.device("cuda", local))
        return f

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_pass(step_fn, filt, profile=False):
        times = []
        for t in range(1, 1 + Wm):
            step_fn(filt, t)
        torch.cuda.synchronize()
        if profile:
            filt.set_profiling(True); filt.profile(reset=True)
        l0 = filt.stats().kernel_launches
        barrier()
        wall0 = time.perf_counter()
        for t in range(1 + Wm, 1 + Wm + K):
            flush_buf.fill_(t & 255)  # L2 flush, outside the event pair
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn(filt, t)
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        barrier()
        wall = time.perf_counter() - wall0
        launches = filt.stats().kernel_launches - l0
        prof = filt.profile(reset=True) if profile else None
        if profile:
            filt.set_profiling(False)
        return np.array(times), wall, launches, prof

    # ---- pass A: frames resident in HBM -------------------------------------------------------
    dev_frames = [torch.from_numpy(f).cuda() for f in frames]
    picks = [scene.picks(t, nfeat) for t in range(scene.n_frames)]

    def step_resident(f, t):
        f.captureNewFrame_device(dev_frames[t].data_ptr(), width, height, width, scene.stamps[t])
        f.predict()
        f.update(picks[t])

    filt = new_filter()
    sampler = ClockSampler(local)
    sampler.start()
    tA, wallA, launches, _ = timed_pass(step_resident, filt)          # the reported value: no per-kernel events in the stream
    clocks = sampler.stop()
    stA = filt.stats()
    n_state = filt.state_dim()
    del filt
    filt = new_filter()                                                 # same sequence again on a fresh filter,
    tP, _, _, prof = timed_pass(step_resident, filt, profile=True)      # with CUDA-event brackets per kernel class
    stA = filt.stats()
    n_state = filt.state_dim()
    del filt

    # ---- pass B: end to end through the public API with host buffers ---------------------------
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    last = {}

    def step_e2e(f, t):
        f.captureNewFrame(pinned[t].numpy(), scene.stamps[t])   # H2D inside the step
        f.predict()
        f.update(picks[t])                                       # D2H of the packed step record inside
        last["state"] = f.getState(); last["sigma"] = f.getSigma()  # accessors the ROS node reads

    filt = new_filter()
    tB, wallB, _, _ = timed_pass(step_e2e, filt)
    h2d = width * height + 4 * nfeat
    d2h = 210 * 8 + (16 + 3 * nfeat) * 4 + 2 * 88 + 14 * 8 + 196 * 8
    peer_memory = bool(filt.dist_info()["peer_memory"]) if partitioned else False
    del filt

    def agg(times):
        tot = float(times.sum())
        if dist is not None:
            tt = torch.tensor([tot], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tot = float(tt.item())
        return tot

    totA, totB = agg(tA), agg(tB)
    nfil = 1 if partitioned else world   # partitioned: all ranks step the same filter
    value = nfil * K / (totA / 1e3)
    e2e_value = nfil * K / (totB / 1e3)

    out = None
    if rank == 0:
        gemm_ms, gemm_launches = prof["downdate_gemm"]
        # algorithmic flops of one rank-128 downdate launch: 2 n^2 k for the full square, n (n + 128) k
        # when only tiles touching the lower triangle are computed (SYRK form)
        flops_per_launch = 2.0 * n_state * n_state * 128 if args.full_square else 1.0 * n_state * (n_state + 128) * 128
        if partitioned:   # a rank downdates its row block over all columns
            r0, r1, _ = pkg.dist.row_block(rank, world, n_state)
            flops_per_launch = 2.0 * (r1 - r0) * n_state * 128
        peak = dgemm_peak_tflops(torch)
        achieved = flops_per_launch * gemm_launches / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tfile):
            try:
                traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        step_ms = float(tP.mean())   # share_of_step is taken inside the profiled pass
        out = {
            "metric": metric_for(args.workload), "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": round(totA / K, 4), "higher_is_better": True, "scaling": "strong" if partitioned else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: single filter, {nfeat} inverse-depth features (n={n_state}), "
                                   f"{width}x{height} u8 frames, predict+match+update per frame, "
                                   + ("all features matched" if args.match_every <= 1 else f"every {args.match_every}th feature visible")
                                   + f" (n_li={stA.n_li})",
                       "l2": "flushed between timed steps (256 MiB write outside the event pairs)",
                       "multi_gpu": ("one filter, stacked update partitioned by covariance row blocks, look-ahead pipeline; partial S blocks "
                                     "and V panels exchanged " + ("by peer-memory stores from inside the producing kernels (NVLink)"
                                                                  if peer_memory else "by NCCL all-reduce / all-gather")
                                     + ", row blocks of Sigma all-gathered by NCCL once per update (strong scaling)") if partitioned else
                                    ("replicas only (one independent filter per rank)" if world > 1 else "n/a")},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "k_gemm_nt_sub (Sigma -= V V^T, DMMA.8x8x4, K=128 per launch, "
                                   + ("rank 0's row block x all columns" if partitioned else
                                      ("full square" if args.full_square else "lower-triangle tiles + mirror")) + ")", "bound": "tensor",
                         "achieved": round(achieved, 3), "peak": round(peak, 2), "unit": "TFLOP/s",
                         "frac": round(achieved / peak, 4) if peak > 0 else None, "traffic": traffic,
                         "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no fp64 entry; "
                                        "DMMA issue peak measured by tools/fp64_probe: 37.05 TFLOP/s)",
                         "flops_per_launch": flops_per_launch, "launches": int(gemm_launches),
                         "avg_launch_ms": round(gemm_ms / max(gemm_launches, 1), 5),
                         "share_of_step": round(gemm_ms / K / step_ms, 4)},
            "kernel_ms_per_step": {k: round(v[0] / K, 5) for k, v in prof.items() if v[1] > 0 or v[0] > 0},
        }
        if world == 1 and not args.no_cpu_baseline:
            if nfeat > 600:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                       "sample": "not run: one dense reference frame at n = 12014 is ~2.8e13 flops (minutes per frame); "
                                                 "see the cfg2 line for the measured CPU baseline"}
            else:
                out["cpu_baseline"] = cpu_baseline(pkg, args.workload, budget_s=args.cpu_budget)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def run_match(args):
    """BASELINE configs[4]: the stateless batched matcher (ekf_match_batch), frames sharded across ranks."""
    import torch
    import ekfb200
    pkg = ekfb200.load_package()
    pkg.lib()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, Wm = args.steps, max(args.warmup, 3)
    M, width, height = WORKLOADS[args.workload]
    F = args.frames
    w = 11
    d = pkg.synth.match_batch_inputs(n_frames=F, features_per_frame=M, width=width, height=height, window=w, seed=1239 + rank)
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    host = {k: torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory() for k in ("frames", "templates", "h", "S")}
    res = {k: host[k].to(dev) for k in host}
    uv = torch.zeros((F * M, 2), dtype=torch.int32, device=dev); sc = torch.zeros(F * M, dtype=torch.float32, device=dev)
    uv_h = torch.zeros((F * M, 2), dtype=torch.int32).pin_memory(); sc_h = torch.zeros(F * M, dtype=torch.float32).pin_memory()

    def launch(t):
        pkg.match_batch(t["frames"].data_ptr(), F, width, height, width, t["templates"].data_ptr(), M, w, t["h"].data_ptr(),
                        t["S"].data_ptr(), uv.data_ptr(), sc.data_ptr(), sigma_size=3.0, stream=stream.cuda_stream)

    def step_resident():
        launch(res)

    stage = {k: torch.empty_like(res[k]) for k in res}

    def step_e2e():
        for k in stage:
            stage[k].copy_(host[k], non_blocking=True)          # H2D inside the step
        launch(stage)
        uv_h.copy_(uv, non_blocking=True); sc_h.copy_(sc, non_blocking=True)   # D2H inside the step

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(Wm):
            fn()
        barrier()
        ts = []
        for _ in range(K):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        barrier()
        return np.array(ts)

    sampler = ClockSampler(local)
    sampler.start()
    tA = timed(step_resident)
    clocks = sampler.stop()
    tB = timed(step_e2e)
    found = int((uv_h[:, 0] >= 0).sum())
    assert np.array_equal(uv_h.numpy()[uv_h.numpy()[:, 0] >= 0], d["truth"][uv_h.numpy()[:, 0] >= 0]), "matches differ from the planted truth"
    totA = pkg.dist.max_over_ranks(float(tA.sum()), dist, "cuda"); totB = pkg.dist.max_over_ranks(float(tB.sum()), dist, "cuda")
    out = None
    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm_peak = float(peaks.get("hbm_gbs", 6456.5))
        delta = 12  # 3 sigma of S = diag(16,16)
        win_bytes = (2 * delta + w) * (2 * delta + w) + w * w + 16 + 32 + 12
        bytes_per_launch = float(win_bytes) * F * M
        ms = float(tA.mean())
        cand = 3.14159 * delta * delta            # in-ellipse candidates per feature
        out = {"metric": METRIC_MATCH, "value": round(world * F * M * K / (totA / 1e3), 1), "unit": UNIT_MATCH, "n_gpus": world,
               "steps": K, "warmup": Wm, "ms_per_step": round(totA / K, 4), "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"{args.workload}: {F} frames {width}x{height} u8 per GPU x {M} features, 11x11 templates, "
                                      f"S = diag(16,16) +- 10 % => ~{cand:.0f} in-ellipse candidates per feature; {found} of {F * M} accepted",
                          "l2": f"inputs larger than L2 ({F * width * height / 1e6:.0f} MB of frames)",
                          "multi_gpu": "sharded by frame, no collective" if world > 1 else "n/a"},
               "e2e": {"value": round(world * F * M * K / (totB / 1e3), 1), "unit": UNIT_MATCH,
                       "h2d_bytes_per_step": int(sum(host[k].numel() * host[k].element_size() for k in host)),
                       "d2h_bytes_per_step": int(uv_h.numel() * 4 + sc_h.numel() * 4)},
               "gpu_launches": K, "clocks": clocks,
               "roofline": {"kernel": "k_match_batch (one CTA per feature; instruction-issue bound on DP4A / integer box sums, not HBM: "
                                      "see issue)", "bound": "hbm",
                            "achieved": round(bytes_per_launch / (ms / 1e3) / 1e9, 2), "peak": hbm_peak, "unit": "GB/s",
                            "frac": round(bytes_per_launch / (ms / 1e3) / 1e9 / hbm_peak, 5), "traffic": None,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs", "bytes_per_launch": bytes_per_launch,
                            "avg_launch_ms": round(ms, 4), "share_of_step": 1.0,
                            "issue": {"candidates_per_launch": cand * F * M,
                                      "ns_per_candidate": round(ms * 1e6 / (cand * F * M), 4),
                                      "note": "profiles/r1g-r1i: smsp issue slots ~70 % busy; every candidate costs an 11x11 u8 dot product "
                                              "(33 DP4A) + exact-integer NCC; only the guard-band candidate is re-scored in fp64"}}}
        if world == 1 and not args.no_cpu_baseline:
            import orc
            orc.build()
            nf = 2
            t0 = time.perf_counter()
            orc.lib(omp=True).orc_set_num_threads(os.cpu_count() or 1)
            orc.match_batch(d["frames"][:nf], d["templates"][:nf * M], d["h"][:nf * M], d["S"][:nf * M], sigma_size=3.0, omp=True)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": round(nf * M / dt, 1), "unit": UNIT_MATCH, "cores": orc.lib(omp=True).orc_num_threads(), "kind": "port",
                                   "sample": f"{nf} frames x {M} features through the oracle's Patch::findMatch (OpenMP over features), {dt:.2f} s"}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def run_batch(args):
    """BASELINE configs[2]: a Monte-Carlo ensemble of independent filters, sharded by filter across
    ranks with no data-path collective (weak scaling: BATCH_FILTERS filters per GPU)."""
    import torch
    import ekfb200
    pkg = ekfb200.load_package()
    pkg.lib()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, Wm = args.steps, max(args.warmup, 3)
    B = args.filters
    nfeat, width, height = WORKLOADS[args.workload]
    scene = make_scene(pkg, args.workload, 1 + Wm + K, seed=1235)
    frames = [scene.frame(t) for t in range(scene.n_frames)]
    cfg = pkg.default_config(**scene.config_overrides())
    stream = torch.cuda.current_stream()

    def new_batch():
        f = pkg.VSlamFilter(cfg, feature_capacity=nfeat + 2, device=local)
        added = seed_filter(f, scene)
        assert added == nfeat
        b = pkg.FilterBatch(cfg, B, feature_capacity=nfeat + 2, device=local)
        b.set_stream(stream.cuda_stream)
        b.seed_from(f)
        # per-hypothesis perturbation of the camera state: filters [rank*B, (rank+1)*B) of the ensemble
        b.set_camera_states(pkg.dist.ensemble_camera_states(f.getState(), pkg.dist.ensemble_slice(rank, world, B)))
        del f
        return b

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    picks = [scene.picks(t, nfeat) for t in range(scene.n_frames)]

    def timed(step_fn, batch):
        times, cls = [], {"predict": 0.0, "match": 0.0, "update": 0.0}
        for t in range(1, 1 + Wm):
            step_fn(batch, t)
        torch.cuda.synchronize()
        l0 = batch.kernel_launches()
        barrier()
        for t in range(1 + Wm, 1 + Wm + K):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn(batch, t)
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            for k, v in batch.last_step_ms().items():
                cls[k] += v
        barrier()
        return np.array(times), batch.kernel_launches() - l0, {k: v / K for k, v in cls.items()}

    dev_frames = [torch.from_numpy(f).cuda() for f in frames]

    def step_resident(b, t):
        b.captureNewFrame_device(dev_frames[t].data_ptr(), width, height, width, scene.stamps[t])
        b.step(picks[t])

    batch = new_batch()
    sampler = ClockSampler(local)
    sampler.start()
    tA, launches, cls = timed(step_resident, batch)
    clocks = sampler.stop()
    mu14, st = batch.camera_states()
    n_state = batch.state_dim(0)
    li_mean = float(st[:, 2].mean())
    del batch

    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    last = {}

    def step_e2e(b, t):
        b.captureNewFrame(pinned[t].numpy(), scene.stamps[t])   # H2D inside the step
        b.step(picks[t])                                        # D2H of camera states + counters inside
        last["mu"], last["st"] = b.camera_states()

    batch = new_batch()
    tB, _, _ = timed(step_e2e, batch)
    del batch
    h2d = width * height + 4 * nfeat
    d2h = B * 14 * 8 + B * 8 * 4

    def agg(times):
        return pkg.dist.max_over_ranks(float(times.sum()), dist, "cuda")

    totA, totB = agg(tA), agg(tB)
    all_cams = pkg.dist.gather_camera_states(last["mu"], dist, "cuda")   # reporting only
    assert all_cams.shape == (world * B, 14) and np.isfinite(all_cams).all()
    out = None
    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm_peak = float(peaks.get("hbm_gbs", 6456.5))
        # k_batch_update: algorithmic bytes = read + write of every filter's covariance (16 n^2)
        bytes_per_launch = 16.0 * n_state * n_state * B
        upd_ms = cls["update"]
        achieved = bytes_per_launch / (upd_ms / 1e3) / 1e9 if upd_ms > 0 else 0.0
        out = {
            "metric": METRIC_BATCH, "value": round(world * B * K / (totA / 1e3), 1), "unit": UNIT_BATCH, "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": round(totA / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {B} independent filters per GPU x {nfeat} inverse-depth features (n={n_state}), "
                                   f"one shared {width}x{height} u8 frame per step, predict+match+update per filter "
                                   f"(mean n_li={li_mean:.1f})",
                       "l2": f"inputs larger than L2 ({B * n_state * n_state * 8 / 1e6:.0f} MB of covariance per step vs 126 MB)",
                       "multi_gpu": "sharded by filter, no data-path collective" if world > 1 else "n/a",
                       "filters_per_gpu": B},
            "e2e": {"value": round(world * B * K / (totB / 1e3), 1), "unit": UNIT_BATCH, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"kernel": "k_batch_update (one CTA per filter: RANSAC, W = Sigma H^T, Cholesky gain, DMMA downdate from smem)",
                         "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                         "frac": round(achieved / hbm_peak, 4), "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs", "bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": round(upd_ms, 4), "share_of_step": round(upd_ms / float(tA.mean()), 4)},
            "kernel_ms_per_step": {k: round(v, 4) for k, v in cls.items()},
        }
        tfile = os.path.join(ROOT, "profiles", "batch_update_traffic.json")
        if os.path.exists(tfile):
            try:
                out["roofline"]["traffic"] = json.load(open(tfile)).get("dram_bytes_per_launch")
            except Exception:
                pass
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_batch(pkg, args.workload, budget_s=args.cpu_budget)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def cpu_baseline_batch(pkg, workload, budget_s=20.0):
    """The oracle (dense reference algebra) on a bounded sample of the ensemble: filters are
    independent, so they are spread over the host cores with one oracle filter per thread."""
    import concurrent.futures as cf
    import orc
    orc.build()
    nfeat = WORKLOADS[workload][0]
    scene = make_scene(pkg, workload, 3)
    cfg = pkg.default_config(**scene.config_overrides())
    cores = os.cpu_count() or 1
    L = orc.lib(omp=False)

    def one(_):
        o = orc.OracleFilter(cfg, kind=0, omp=False)
        seed_filter(o, scene)
        t0 = time.perf_counter()
        n = 0
        for t in (1, 2):
            o.captureNewFrame(scene.frame(t), scene.stamps[t]); o.predict(); o.update(scene.picks(t, nfeat))
            n += 1
        return n, time.perf_counter() - t0

    nfil = max(cores, 8)
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(max_workers=cores) as ex:   # ctypes releases the GIL inside the oracle calls
        res = list(ex.map(one, range(nfil)))
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    busy = sum(r[1] for r in res)
    return {"value": round(steps / (busy / cores), 3), "unit": UNIT_BATCH, "cores": cores, "kind": "port",
            "sample": f"{nfil} filters x 2 steps of {workload} on {cores} threads (one oracle filter per thread, dense reference "
                      f"algebra, fp64), {busy / steps * 1e3:.1f} ms per filter-step per core, wall {wall:.1f} s incl. seeding"}


def _oracle_seeded(pkg, orc, workload, n_frames, omp=True):
    """An oracle filter holding the workload's seeded map.  Seeding goes through a structured
    (O(n)-per-feature) twin when a GPU is present, else through the oracle's own dense addFeature."""
    nfeat, width, height = WORKLOADS[workload]
    scene = make_scene(pkg, workload, n_frames)
    cfg = pkg.default_config(**scene.config_overrides())
    o = orc.OracleFilter(cfg, kind=0, omp=omp)
    o.captureNewFrame(scene.frame(0), scene.stamps[0])
    seeded = False
    try:
        import torch
        if torch.cuda.is_available():
            g = pkg.VSlamFilter(cfg, feature_capacity=nfeat + 4)
            seed_filter(g, scene)
            o.import_from(g)   # identical (mu, Sigma, templates) as the GPU arm starts from
            del g
            seeded = True
    except Exception as e:  # no GPU: dense seeding (slow for large maps)
        log("oracle seeding via GPU failed:", e)
    if not seeded:
        seed_filter(o, scene)
    return o, scene


def cpu_baseline(pkg, workload, budget_s=30.0):
    """The oracle (dense reference algebra, vslamRansac.cpp as written) on this host's cores, on a
    bounded sample of the same workload."""
    import orc
    orc.build()
    L = orc.lib(omp=True)
    L.orc_set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1; the baseline uses every host core
    cores = L.orc_num_threads()
    nfeat = WORKLOADS[workload][0]
    o, scene = _oracle_seeded(pkg, orc, workload, 4)
    times = []
    t_start = time.perf_counter()
    for t in range(1, 4):
        img = scene.frame(t)
        t0 = time.perf_counter()
        o.captureNewFrame(img, scene.stamps[t]); o.predict(); o.update(scene.picks(t, nfeat))
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    per = float(np.mean(times))
    return {"value": round(1.0 / per, 5), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} frame(s) of {workload} (dense n x n algebra as the reference executes it, "
                      f"fp64, OpenMP over GEMM rows), {per:.2f} s/frame"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ekfb200
    pkg = ekfb200.load_package()
    import orc
    orc.build()
    L = orc.lib(omp=True)
    L.orc_set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core
    cores = L.orc_num_threads()
    nfeat = WORKLOADS[args.workload][0]
    K, Wm = args.steps, args.warmup
    budget = args.ref_budget
    o, scene = _oracle_seeded(pkg, orc, args.workload, 1 + Wm + K)
    t_begin = time.perf_counter()
    times = []
    done_w = 0
    for t in range(1, 1 + Wm + K):
        img = scene.frame(t)
        t0 = time.perf_counter()
        o.captureNewFrame(img, scene.stamps[t]); o.predict(); o.update(scene.picks(t, nfeat))
        dt = time.perf_counter() - t0
        if done_w < Wm:
            done_w += 1
        else:
            times.append(dt)
        if time.perf_counter() - t_begin > budget and len(times) >= 1:
            break
    per = float(np.mean(times))
    val = 1.0 / per
    n_state = o.state_dim()
    out = {"impl": "reference", "metric": metric_for(args.workload), "value": round(val, 5), "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
           "steps": len(times), "warmup": done_w, "ms_per_step": round(per * 1e3, 2), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{args.workload}: single filter, {nfeat} inverse-depth features (n={n_state}), CPU oracle of the "
                                  f"reference's dense algebra (the reference itself needs Eigen/OpenCV/ROS, absent here)",
                      "bounded": f"stopped after {len(times)} timed frame(s) (time budget {budget:.0f} s)"},
           "cpu_baseline": {"value": round(val, 5), "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{len(times)} timed frame(s), {per:.2f} s/frame"},
           "e2e": {"value": round(val, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    # keep stdout to the single JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_n500", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--match-every", type=int, default=1,
                    help="single-filter workloads: only every k-th feature is visible after frame 0 (m = N / k matches per frame; "
                         "the headline uses k = 1, all features matched, the stated worst case)")
    ap.add_argument("--full-square", action="store_true", help="downdate all n x n tiles instead of lower triangle + mirror")
    ap.add_argument("--frames", type=int, default=MATCH_FRAMES, help="frames per GPU of the matcher workload")
    ap.add_argument("--filters", type=int, default=BATCH_FILTERS, help="filters per GPU of the batched workload")
    ap.add_argument("--cpu-budget", type=float, default=30.0)
    ap.add_argument("--ref-budget", type=float, default=150.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload.startswith("cfg3"):
        run_batch(args)
    elif args.workload.startswith("cfg5"):
        run_match(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
