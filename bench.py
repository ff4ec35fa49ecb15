#!/usr/bin/env python
"""bench.py -- EKF predict+update steps/s at N landmarks, with the covariance sweep's HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one whole Robot::localize (slam_ros/Robot.cpp:126-943): prediction, association of the m
observed lines against every map landmark, the matched updates (folded into one rank-2m covariance
sweep), augmentation.  Workloads (BASELINE.json `configs`):

    10k      configs[2]  single filter, 10 000 line landmarks (P 20003^2 fp64 = 3.2 GB), m = 8   [default]
    1k       configs[1]  single filter, 1 000 landmarks (P fits L2: launch-bound, not HBM-bound); its CPU baseline /
                         reference arm is the LITERAL reference compiled with LINESIZE = 1000 (oracle/_ref/libslamref1k.so)
    40k      configs[4]  single filter, 40 000 landmarks (51 GB); with --gpus N > 1 row-sharded over N ranks
    mc       configs[3]  Monte-Carlo batch, 4096 independent filters x 50 landmarks, sharded over ranks
    room     configs[0]  the reference-sized filter (LINESIZE=100) on the synthetic room; its CPU baseline /
                         reference arm is the LITERAL reference (oracle/_ref: Robot.cpp compiled over the GSL shim)
    extract  (8f row 2)  line extraction of 361-beam scans (scans/s); CPU baseline = the reference's own
                         lineFitting.cpp (oracle/_ref/libslamlines.so) where built, else the restatement

With --gpus N > 1 (launched by torchrun, one rank per GPU) the default workload runs N independent
filters (replicas; the path needs no collective) -> "scaling": "weak"; `40k` runs ONE filter row-sharded
over the ranks with the per-update NCCL exchange -> "scaling": "strong".

Timing: W warm-up steps, then exactly K steps bracketed by CUDA events recorded on the library's own
stream (ekf_timer_start/stop) after a barrier + synchronize; max over ranks.  The working set (3.2 GB)
is far larger than L2 (126 MB), so no flush is needed between iterations (stated in config.l2).
`value` has the K steps' inputs already resident in HBM (ekf_scan_device); `e2e` drives the public
host-buffer call (ekf_scan: pinned H2D of the step's inputs, D2H of the matches + pose, every step).

The oracle (oracle/) is executed here only as the CPU baseline / reference arm, never as the product.
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
sys.path.insert(0, ROOT)

WORKLOADS = {
    "10k": dict(N=10000, m=8, headroom=1024, desc="configs[2]: single filter, 10k line landmarks, m=8 batched multi-line updates"),
    "1k": dict(N=1000, m=8, headroom=512, desc="configs[1]: single filter, 1k line landmarks, m=8"),
    "40k": dict(N=40000, m=8, headroom=1024, desc="configs[4]: single filter, 40k line landmarks, m=8"),
    "mc": dict(N=50, m=8, headroom=14, filters=4096, desc="configs[3]: Monte-Carlo batch, 4096 filters x 50 landmarks, m=8"),
    "room": dict(N=100, m=9, headroom=0, desc="configs[0]: the reference-sized filter (LINESIZE=100) on the synthetic 2-D room, 361-beam scans"),
    "extract": dict(N=0, m=0, headroom=0, desc="SURVEY 8f-2: line extraction (mapping_cb + LineExtraction) of 361-beam scans of the synthetic room"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the structured oracle (port of Robot::localize), all host threads
# ------------------------------------------------------------------------------------------------
def cpu_run(workload, steps, warmup, budget_s, threads=None):
    """Times the CPU restatement of the reference on the same workload.  Returns (steps_per_s, info)."""
    from oracle.oracle import StructuredOracle, build
    from slam_ros_b200 import scenario as sc
    build()
    w = WORKLOADS[workload]
    N, m = w["N"], w["m"]
    threads = threads or (os.cpu_count() or 1)
    if workload == "mc":
        # independent filters: time a bounded number of filters for `steps` scans each, one thread each is
        # the natural CPU mapping; we time filters sequentially on one thread and report filter-steps/s x 1
        so = StructuredOracle(N + w["headroom"], threads=1)
        scn = sc.map_scenario(N, warmup + steps, m=m, seed=1000)
        so.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
        for s in range(warmup):
            so.scan(scn["u"][s], scn["z"][s], scn["R"][s])
        t0 = time.perf_counter()
        done = 0
        for s in range(warmup, warmup + steps):
            so.scan(scn["u"][s], scn["z"][s], scn["R"][s])
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        fps = done / dt                         # filter-steps per second on one core
        return fps / w["filters"], {"kind": "port", "cores": 1, "value": fps / w["filters"],
                                    "sample": "%d scans of ONE 50-landmark filter on one thread; batch steps/s = filter-steps/s / %d filters (the reference has no threads)" % (done, w["filters"])}
    so = StructuredOracle(N + w["headroom"], threads=threads)
    scn = sc.map_scenario(N, warmup + steps, m=m, seed=1)
    so.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
    t_w = time.perf_counter()
    for s in range(min(warmup, 1)):
        so.scan(scn["u"][s], scn["z"][s], scn["R"][s])
    t_w = time.perf_counter() - t_w
    t0 = time.perf_counter()
    done = 0
    for s in range(min(warmup, 1), min(warmup, 1) + steps):
        so.scan(scn["u"][s], scn["z"][s], scn["R"][s])
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    val = done / dt
    n = 3 + 2 * N
    # the literal Robot::localize spends 7.1 ms in its dense n^3 prediction dgemms and 0.45 ms per
    # matched line in n^2 passes at n = 203 (LINESIZE = 100, this container's host); scaled, NOT measured
    # + ~0.05 ms per (line, landmark) gate pair; check: the LINESIZE = 1000 build of the reference itself measures 12.7 s
    # per step at n = 2003, m = 8 (bench.py --workload 1k), this formula gives 11.1 s
    lit_ms = 7.1 * (n / 203.0) ** 3 + m * 0.45 * (n / 203.0) ** 2 + m * N * 0.05 * (n / 203.0)
    info = {"kind": "port", "cores": so.threads, "value": val, "unit": "steps/s",
            "literal_reference_extrapolated": "EXTRAPOLATED, not measured: the reference's own dense GSL path would need about "
                                              "%.3g s per step at n = %d (2 n^3 MACs per prediction dgemm)" % (lit_ms / 1e3, n),
            "sample": "%d full steps (of %d requested) of the same workload on the structured oracle (oracle/ekf_oracle.cpp, "
                      "the runtime-capacity restatement of Robot::localize; OpenMP over the %d host threads for the n^2 "
                      "row sweeps; the literal reference is fixed at LINESIZE=100 and cannot run this size)" % (done, steps, so.threads)}
    return val, info


def _have_literal_1k():
    from oracle.oracle import have_literal_1k
    return have_literal_1k()


def literal_1k_run(steps, budget_s):
    """configs[1] on the reference ITSELF: oracle/_ref/libslamref1k.so = slam_ros/Robot.cpp compiled with
    LINESIZE = 1000 (the two size macros of Robot.h rewritten into a generated header), single thread.  980 landmarks
    (its map resets above LINESIZE - 10); one call is ~12 s (dense n^3 prediction), so the sample is 1-2 steps."""
    from oracle.oracle import LiteralReference
    from slam_ros_b200 import scenario as sc
    N, m = 980, 8
    scn = sc.map_scenario(N, steps + 1, m=m, seed=1)
    lit = LiteralReference(big=True)
    zero = np.zeros(3)
    lit.localize(scn["seed_z"], scn["seed_R"], sc.encoder_for(zero, zero))
    t0 = time.perf_counter(); done = 0
    for s in range(steps):
        y, P, L, pose = lit.state()
        lit.localize(scn["z"][s], scn["R"][s], sc.encoder_for(pose, scn["u"][s]))
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    val = done / dt
    return val, {"kind": "reference", "cores": 1, "value": val, "unit": "steps/s",
                 "sample": "%d Robot::localize calls of the literal reference compiled with LINESIZE=1000 (n = 2003; g++ -O2, GSL shim, "
                           "cout disabled) on a 980-landmark map, m = 8 -- the structured oracle is bitwise equal to it at this size "
                           "(tests/test_oracle.py)" % done}


def literal_run(steps, warmup, budget_s):
    """configs[0] on the reference ITSELF: oracle/_ref/libslamref.so = slam_ros/Robot.cpp (Q1-patched on a pipe)
    compiled -O2 over the GSL shim, single thread (the reference has none), std::cout disabled (Q14)."""
    from oracle.oracle import LiteralReference, have_literal
    from slam_ros_b200 import scenario as sc
    if not have_literal():
        return None, {"kind": "reference", "unavailable": "oracle/_ref/libslamref.so not present (built only where /root/reference exists)"}
    room = sc.room_scenario(steps=warmup + steps, seed=7, range_sigma=5e-5)
    lit = LiteralReference()
    def step(s):
        m = room["count"][s]
        y, P, L, pose = lit.state()
        lit.localize(room["z"][s, :m], room["R"][s, :m], sc.encoder_for(pose, room["u"][s]))
    for s in range(warmup):
        step(s)
    t0 = time.perf_counter()
    done = 0
    for s in range(warmup, warmup + steps):
        step(s)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    val = done / dt
    return val, {"kind": "reference", "cores": 1, "value": val, "unit": "steps/s",
                 "sample": "%d Robot::localize calls of the literal reference (LINESIZE=100, g++ -O2, GSL shim, cout disabled) on the room scenario, including its state read-back through the harness" % done}


def run_room(args, rank, world, local):
    """configs[0] on the GPU through the host-buffer call (the drop-in's real use: one small filter, 10 Hz node)."""
    import torch
    from slam_ros_b200 import EkfFilter, scenario as sc
    K, W = args.steps, args.warmup
    torch.cuda.set_device(local)
    room = sc.room_scenario(steps=W + K, seed=7, range_sigma=5e-5)
    f = EkfFilter(capacity_lines=100, device=local)
    f.profile_enable(True)
    for s in range(W):
        m = room["count"][s]
        f.scan(room["u"][s], room["z"][s, :m], room["R"][s, :m])
    f.sync(); f.profile_read()
    clocks = ClockSampler(local); clocks.start()
    t0 = time.perf_counter()
    for s in range(W, W + K):
        m = room["count"][s]
        f.scan(room["u"][s], room["z"][s, :m], room["R"][s, :m])
    f.sync()
    ms = (time.perf_counter() - t0) * 1e3
    clk = clocks.stop()
    prof = f.profile_read()
    peak, peak_src = measured_peaks()
    n = 203
    val = K / (ms / 1e3)
    sweep_ms = prof["sweep_ms"] / max(prof["sweeps"], 1)
    bytes_per_sweep = 8.0 * n * (n + 1)
    return {
        "metric": "EKF predict+update steps/s at N landmarks", "value": val, "unit": "steps/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS["room"]["desc"], "capacity_lines": 100, "lines_per_scan_mean": float(room["count"][W:W + K].mean()),
                   "l2": "P is 330 KB: resident in L2, the path is launch / latency bound (value == e2e: host buffers every step)"},
        "roofline": {"bound": "hbm", "kernel": "k_sweep_pipe", "achieved": bytes_per_sweep / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0,
                     "peak": peak, "unit": "GB/s", "frac": (bytes_per_sweep / (sweep_ms * 1e-3) / 1e9 / peak) if sweep_ms > 0 else 0.0,
                     "peak_source": peak_src, "launch_ms": sweep_ms, "launches_timed": prof["sweeps"],
                     "algorithmic_bytes_per_launch": bytes_per_sweep, "traffic": None,
                     "note": "not an HBM-bound configuration: 330 KB per sweep"},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": (6 + 6 * 9) * 8, "d2h_bytes_per_step": 4 * 9 + 128, "ms_per_step": ms / K},
        "gpu_launches": prof["launches"], "clocks": clk,
    }


def lines_cpu_run(steps, budget_s):
    """Line extraction on the host: the reference's own sources (deterministic build) when present, else the port."""
    from oracle.oracle import LinesOracle, LiteralLineExtraction, have_literal_lines
    from slam_ros_b200 import scenario as sc
    kind = "reference" if have_literal_lines() else "port"
    ex = LiteralLineExtraction() if kind == "reference" else LinesOracle()
    S = sc.room_scans(steps=max(steps, 4), seed=17, range_sigma=2e-3)
    ex.extract(S["scans"][0])
    t0 = time.perf_counter(); done = 0
    for s in range(steps):
        ex.extract(S["scans"][s]); done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    what = ("LineExtraction of slam_ros/lineFitting.cpp itself (g++ -O2, GSL shim, zero-initialised locals; it also writes "
            "three text files per call, as in the node)") if kind == "reference" else "oracle/lines_oracle.cpp (g++ -O2)"
    return done / dt, {"kind": kind, "cores": 1, "value": done / dt, "unit": "scans/s",
                       "sample": "%d scans of the same payloads through %s" % (done, what)}


def run_extract(args, rank, world, local):
    """payload -> lines.  value: payloads resident in HBM, results left in HBM (ekf_lx_extract_device,
    K scans enqueued back to back); e2e: host payload in, host lines out, one sync per scan (ekf_lx_extract)."""
    import torch
    from slam_ros_b200 import LineExtractor, scenario as sc
    K, W = args.steps, args.warmup
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    S = sc.room_scans(steps=W + K, seed=17, range_sigma=2e-3)
    scans = S["scans"]; beams = scans.shape[1]
    lx = LineExtractor(device=local)
    d_scans = torch.tensor(scans, dtype=torch.float32, device=dev)
    for s in range(W):
        lx.extract_device(d_scans[s].data_ptr(), beams)
    lx.sync(); torch.cuda.synchronize()
    clocks = ClockSampler(local); clocks.start()
    t0 = time.perf_counter()
    for s in range(W, W + K):
        lx.extract_device(d_scans[s].data_ptr(), beams)
    lx.sync()
    ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); nl = 0
    for s in range(W, W + K):
        rows, n = lx.extract(scans[s]); nl += n
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clk = clocks.stop()
    peak, peak_src = measured_peaks()
    bytes_per_scan = 8.0 * beams + 6 * 8.0 * beams + 10 * 8.0 * (nl / K)     # payload + point arrays + lines
    return {
        "metric": "line-extraction scans/s (361 beams -> (alfa, r, C_AR, end points) per line)", "value": K / (ms / 1e3), "unit": "scans/s",
        "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS["extract"]["desc"], "beams": int(beams), "lines_per_scan_mean": nl / K,
                   "l2": "a scan is 3 KB: the path is three small dependent kernels, launch / latency bound (no L2 flush applies)"},
        "roofline": {"bound": "hbm", "kernel": "k_lx_segments (one thread block per 0.5 m segment: split recursion + finite-difference covariance)",
                     "achieved": bytes_per_scan / (ms / K * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bytes_per_scan / (ms / K * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_per_scan, "traffic": None,
                     "note": "latency-bound by construction (25 KB per scan); reported for the contract, not a roofline claim"},
        "e2e": {"value": K / (e2e_ms / 1e3), "unit": "scans/s", "h2d_bytes_per_step": int(8 * beams), "d2h_bytes_per_step": 4 + 80 * 128,
                "ms_per_step": e2e_ms / K},
        "gpu_launches": 3 * 2 * K, "clocks": clk,
    }


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    if args.workload == "extract":
        val, info = lines_cpu_run(args.steps, budget_s=150.0)
        print(json.dumps({"impl": "reference", "metric": "line-extraction scans/s (361 beams -> (alfa, r, C_AR, end points) per line)",
                          "value": val, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": {"workload": w["desc"], "beams": 361}, "cpu_baseline": info,
                          "e2e": {"value": val, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    if args.workload == "1k" and args.lines == 0 and _have_literal_1k():
        val, info = literal_1k_run(min(args.steps, 8), budget_s=150.0)
    elif args.workload == "room":
        val, info = literal_run(args.steps, args.warmup, budget_s=150.0)
        if val is None:
            print(json.dumps({"impl": "reference", "unavailable": info["unavailable"]}), flush=True)
            return
    else:
        val, info = cpu_run(args.workload, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": "EKF predict+update steps/s at N landmarks", "value": val, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / val,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "landmarks": w["N"], "lines_per_scan": w["m"]},
        "cpu_baseline": info,
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
def run_single_or_replicas(args, rank, world, local, sharded):
    import torch
    import torch.distributed as dist
    from slam_ros_b200 import EkfFilter, scenario as sc
    from slam_ros_b200.ekf import nccl_unique_id

    w = WORKLOADS[args.workload]
    N, m = w["N"], w["m"]
    K, W = args.steps, args.warmup
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)

    shard = None
    if sharded and world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(uid, 0)
        shard = (rank, world, bytes(uid.cpu().tolist()))
    seed = 1 if (sharded or world == 1) else 1 + rank
    total_steps = W + K + K            # device-resident leg, then the host-buffer (e2e) leg
    scn = sc.map_scenario(N, total_steps, m=m, seed=seed)
    f = EkfFilter(capacity_lines=N + w["headroom"], device=local, shard=shard)
    exchange = None
    if shard is not None:
        # fused exchange over NVLink peer memory inside the line-loop kernel (EKF_SHARD_NCCL=1 keeps the NCCL path)
        from slam_ros_b200.parallel import connect_shards
        fused = os.environ.get("EKF_SHARD_NCCL", "0") != "1" and connect_shards(f, dev)
        exchange = "in-kernel NVLink stores (CUDA IPC peer memory)" if fused else "ncclAllReduce per matched line"
    rc, j, pose = f.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
    assert rc == 0 and f.lines == N, (rc, f.lines)

    # ---- leg 1: inputs resident in HBM --------------------------------------------------------
    d_u = torch.tensor(scn["u"], dtype=torch.float64, device=dev)
    d_z = torch.tensor(scn["z"], dtype=torch.float64, device=dev)
    d_R = torch.tensor(scn["R"], dtype=torch.float64, device=dev)
    d_j = torch.full((total_steps, m), -7, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def dev_step(s):
        f.scan_device(d_u[s].data_ptr(), m, d_z[s].data_ptr(), d_R[s].data_ptr(), d_j[s].data_ptr())

    for s in range(W):
        dev_step(s)
    f.sync()
    f.profile_read()
    f.profile_enable(True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    f.timer_start()
    for s in range(W, W + K):
        dev_step(s)
    ms = f.timer_stop()
    torch.cuda.synchronize()
    clk = clocks.stop() if rank == 0 else None
    lprof = f.profile_read_lines()
    prof = f.profile_read()
    f.profile_enable(False)
    matched = int((d_j[W:W + K] >= 0).sum().item())
    pose_dev, L_dev, st_dev = f.state()

    # ---- leg 2: end to end through the host-buffer call -------------------------------------------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_matched = 0
    for s in range(W + K, W + 2 * K):
        rc, jj, pose = f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
        e2e_matched += int((jj >= 0).sum())
    f.sync()
    e2e_ms = (time.perf_counter() - t0) * 1e3

    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    filters = 1 if sharded else world
    value = filters * K / (ms_max / 1e3)
    e2e_value = filters * K / (e2e_ms_max / 1e3)

    if rank != 0:
        return None
    n = 3 + 2 * L_dev
    peak, peak_src = measured_peaks()
    sweep_ms = prof["sweep_ms"] / max(prof["sweeps"], 1)
    # algorithmic bytes of one sweep: upper triangle read + written once, plus K and KS of the folded
    # terms.  Row-sharded: each rank sweeps 1/world of the triangle (rank 0's share is timed).
    mean_terms = matched / max(K, 1)
    bytes_per_sweep = (8.0 * n * (n + 1) + 32.0 * n * mean_terms) / (world if sharded else 1)
    achieved = bytes_per_sweep / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0
    line = {
        "metric": "EKF predict+update steps/s at N landmarks", "value": value, "unit": "steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
        "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "landmarks": N, "state_dim": n, "lines_per_scan": m,
                   "matched_per_step": mean_terms, "parallelism": ("row-sharded P x%d" % world) if sharded else ("independent filters x%d" % world),
                   "exchange": exchange,
                   "l2": "per-step working set %.2f GB read + %.2f GB written >> 126 MB L2: no flush needed" % (8.0 * n * (n + 1) / 2 / 1e9, 8.0 * n * (n + 1) / 2 / 1e9),
                   "seed": seed},
        "roofline": {"bound": "hbm", "kernel": "k_sweep_quad (P -= (K S) K' over the upper triangle; TMA + mbarrier ring, 8x4 register tiles, runs under the next scan's line loop)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                     "launch_ms": sweep_ms, "launches_timed": prof["sweeps"], "algorithmic_bytes_per_launch": bytes_per_sweep,
                     "sweep_share_of_step": (prof["sweep_ms"] / ms) if ms > 0 else None,
                     "line_stream_ms_per_step": (lprof["line_ms"] / lprof["scans"]) if lprof["scans"] else None,
                     "traffic": 3.167e9 if (N == 10000 and not sharded and m == 8) else None,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, profiles/r1_ncu_sweep_quad.csv (10k workload, 8 terms: 1.615 GB read + 1.552 GB written)"},
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": (6 + 2 * m) * 8 + 4 * m * 8,
                "d2h_bytes_per_step": 4 * m + 128, "ms_per_step": e2e_ms_max / K, "matched_per_step": e2e_matched / max(K, 1)},
        "gpu_launches": prof["launches"],
        "clocks": clk,
    }
    return line


def run_monte_carlo(args, rank, world, local):
    import torch
    import torch.distributed as dist
    from slam_ros_b200 import EkfBatch, scenario as sc

    w = WORKLOADS["mc"]
    N, m, B_total = w["N"], w["m"], w["filters"]
    K, W = args.steps, args.warmup
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    from slam_ros_b200.parallel import filters_of_rank
    weak = getattr(args, "mc_per_gpu", 0) > 0
    if weak:                                    # weak scaling: a fixed batch per GPU (the 4096-filter batch is the 1-GPU case)
        B_total = args.mc_per_gpu * world
    ids = filters_of_rank(B_total, rank, world)
    B = len(ids)
    cap = N + w["headroom"]
    total = W + 2 * K
    rng_scn = [sc.map_scenario(N, total, m=m, seed=1000 + int(i)) for i in ids[:64]]   # 64 distinct filters, tiled
    reps = (B + len(rng_scn) - 1) // len(rng_scn)
    seed_z = np.concatenate([np.stack([s["seed_z"] for s in rng_scn])] * reps)[:B]
    seed_R = np.concatenate([np.stack([s["seed_R"] for s in rng_scn])] * reps)[:B]
    U = np.concatenate([np.stack([s["u"] for s in rng_scn])] * reps)[:B]            # (B, total, 3)
    Z = np.concatenate([np.stack([s["z"] for s in rng_scn])] * reps)[:B]            # (B, total, m, 2)
    Rr = np.concatenate([np.stack([s["R"] for s in rng_scn])] * reps)[:B]
    bt = EkfBatch(B, capacity_lines=cap, device=local)
    rc, j, pose = bt.scan(np.zeros((B, 3)), seed_z, seed_R)
    d_u = torch.tensor(np.ascontiguousarray(U.transpose(1, 0, 2)), dtype=torch.float64, device=dev)        # (total, B, 3)
    d_z = torch.tensor(np.ascontiguousarray(Z.transpose(1, 0, 2, 3)), dtype=torch.float64, device=dev)     # (total, B, m, 2)
    d_R = torch.tensor(np.ascontiguousarray(Rr.transpose(1, 0, 2, 3)), dtype=torch.float64, device=dev)
    d_j = torch.zeros((B, m), dtype=torch.int32, device=dev)
    for s in range(W):
        bt.scan_device(d_u[s].data_ptr(), m, d_z[s].data_ptr(), d_R[s].data_ptr(), d_j.data_ptr())
    bt.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    t0 = time.perf_counter()
    for s in range(W, W + K):
        bt.scan_device(d_u[s].data_ptr(), m, d_z[s].data_ptr(), d_R[s].data_ptr(), d_j.data_ptr())
    bt.sync()
    ms = (time.perf_counter() - t0) * 1e3
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(W + K, W + 2 * K):
        rc, jj, pose = bt.scan(np.ascontiguousarray(U[:, s]), np.ascontiguousarray(Z[:, s]), np.ascontiguousarray(Rr[:, s]))
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    ms_max, e2e_max = float(t[0]), float(t[1])
    n = 3 + 2 * N                               # live dimension of every filter
    peak, peak_src = measured_peaks()
    bytes_per_launch = 8.0 * n * (n + 1) * B    # upper triangle of every filter read + written once per scan
    achieved = bytes_per_launch / (ms_max / K * 1e-3) / 1e9
    return {
        "metric": "EKF predict+update steps/s at N landmarks", "value": K / (ms_max / 1e3), "unit": "steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
        "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "filters_total": B_total, "filters_per_gpu": B, "landmarks": N,
                   "lines_per_scan": m, "filter_steps_per_s": B_total * K / (ms_max / 1e3),
                   "parallelism": "independent filters, %d per GPU, no collective" % B,
                   "l2": "batch state %.0f MB > 126 MB L2" % (B * (3 + 2 * cap) ** 2 * 8 / 1e6)},
        "roofline": {"bound": "hbm", "kernel": "k_batch_scan (whole localize per CTA; hot state + pending gains in shared memory, one deferred sweep of the cold upper triangle)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "launch_ms": ms_max / K, "algorithmic_bytes_per_launch": bytes_per_launch, "traffic": None},
        "e2e": {"value": K / (e2e_max / 1e3), "unit": "steps/s", "h2d_bytes_per_step": (3 + 6 * m) * 8 * B,
                "d2h_bytes_per_step": (4 * m + 32) * B, "ms_per_step": e2e_max / K},
        "gpu_launches": K, "clocks": clk,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="10k", choices=sorted(WORKLOADS))
    ap.add_argument("--lines", type=int, default=0, help="observed lines per scan (default: the workload's m = 8; "
                    "configs[2] also names m = 32 and 64: one rank-2m sweep per scan)")
    ap.add_argument("--mc-per-gpu", type=int, default=0, help="Monte-Carlo workload: filters PER GPU (weak scaling) instead "
                    "of splitting the 4096-filter batch over the ranks (strong scaling, default)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline sample")
    args = ap.parse_args()
    if args.lines > 0 and args.workload in ("10k", "1k", "40k"):
        w = WORKLOADS[args.workload]
        w["desc"] = w["desc"].replace("m=%d" % w["m"], "m=%d" % args.lines)
        w["m"] = args.lines
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- libekfcuda has no CPU fallback (use --impl reference for the CPU arm)")
    if args.workload == "mc":
        line = run_monte_carlo(args, rank, world, local)
    elif args.workload == "room":
        line = run_room(args, rank, world, local) if rank == 0 else None
    elif args.workload == "extract":
        line = run_extract(args, rank, world, local) if rank == 0 else None
    else:
        sharded = args.workload == "40k" and world > 1
        line = run_single_or_replicas(args, rank, world, local, sharded)
    if rank == 0 and line is not None:
        if world == 1 and not args.no_cpu_baseline:
            if args.workload == "extract":
                val, info = lines_cpu_run(steps=200, budget_s=args.cpu_budget)
            elif args.workload == "room":
                val, info = literal_run(steps=1000, warmup=3, budget_s=args.cpu_budget)
            elif args.workload == "1k" and args.lines == 0 and _have_literal_1k():
                val, info = literal_1k_run(steps=2, budget_s=args.cpu_budget)
            else:
                val, info = cpu_run(args.workload, steps=1000, warmup=1, budget_s=args.cpu_budget)
            line["cpu_baseline"] = info
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
