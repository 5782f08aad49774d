#!/usr/bin/env python
"""Headline benchmark: BA observations/sec + descriptor-pairs/sec (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path, host cores

Workload (config.workload): the synthetic cube scene of BASELINE.json configs[3],
500 cameras / 200k points / 2M observations (exactly 10 observations per point), which fits
one B200 and is the configuration the north-star target is quoted on.  One *step* is

  BA     one full `bundle()` of that scene: Levenberg-Marquardt to convergence (SoftLOneLoss,
         cameras optimised, <= 100 iterations) from the seed-43 perturbed start;
  MATCH  symmetric brute-force matching of every image with its 8 nearest cameras
         (unique unordered pairs), 128-D integer-valued float32 descriptors, ratio 0.8.

`value` = BA observations/sec = N_obs x LM iterations / CUDA-event time of the LM loop with the
problem resident in HBM; `match.value` = descriptor pairs (2 directions) / device time of the
batch with descriptors resident.  `e2e` is the same metric through the public Python API with
host buffers (H2D of the problem / descriptors and D2H of the results inside the timed region).
Multi-GPU: BA shards observations by point (one all-reduce of the reduced camera system per LM
iteration), MATCH shards the pair list; total work is fixed => "scaling": "strong".
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: cameras, points, obs/point, neighbours per image for matching
    "c4": dict(cameras=500, points=200000, obs_per_point=10, neighbours=8,
               label="synthetic cube 500 cameras / 200k points / 2M observations (BASELINE configs[3])"),
    "c2": dict(cameras=50, points=5000, obs_per_point=None, neighbours=49,
               label="synthetic cube 50 cameras / 5k points (BASELINE configs[1])"),
    "tiny": dict(cameras=12, points=1500, obs_per_point=6, neighbours=4, label="tiny smoke workload"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def build_workload(name):
    from opensfm_b200 import synthetic as syn

    w = WORKLOADS[name]
    sc = syn.cube_scene(w["cameras"], w["points"], 1.0, seed=42, with_descriptors=True,
                        max_obs_per_point=w["obs_per_point"])
    pb = syn.scene_to_problem(sc)
    # per-image descriptor matrices (rows of the visible points); observations are shot-major
    starts = np.searchsorted(sc.obs_shot, np.arange(sc.num_shots + 1))
    feats = [np.ascontiguousarray(sc.track_descriptors[sc.obs_point[starts[s]:starts[s + 1]]]) for s in range(sc.num_shots)]
    # pair list: each image with its k nearest cameras, unique unordered pairs (pairs_selection-style)
    k = min(w["neighbours"], sc.num_shots - 1)
    d = np.linalg.norm(sc.origins[:, None, :] - sc.origins[None, :, :], axis=2)
    np.fill_diagonal(d, np.inf)
    nn = np.argsort(d, axis=1, kind="stable")[:, :k]
    pairs = sorted({(min(i, int(j)), max(i, int(j))) for i in range(sc.num_shots) for j in nn[i]})
    pairs = [p for p in pairs if len(feats[p[0]]) and len(feats[p[1]])]
    return pb, feats, pairs, w


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path on the host cores
# --------------------------------------------------------------------------------------
def cpu_ba_sample(pb, iterations):
    """CPU restatement of the Ceres path (oracle, OpenMP over all host threads): LM iterations on the
    same problem, stopped after `iterations` (bounded sample)."""
    from oracle import ba_lm

    res = ba_lm.solve(pb, stop_after_iterations=iterations)
    its = max(res["iterations"], 1)
    return pb.num_observations * its / res["time_run"], its, res["time_run"]


def cpu_match_sample(feats, pairs, npairs, threads):
    """The reference matcher itself (cv2 BFMatcher through opensfm/matching.py:723-777 semantics) on a
    bounded sample of the pair list, in a joblib *threading* pool like opensfm/context.py:47-67."""
    import cv2
    from joblib import Parallel, delayed

    from oracle import match_oracle as mo

    sample = pairs[:npairs]
    cfg = {"lowes_ratio": 0.8}
    cv2.setNumThreads(0)  # context.py:52-53

    def one(p):
        return len(mo.match_brute_force_symmetric(feats[p[0]], feats[p[1]], cfg))

    t0 = time.perf_counter()
    Parallel(n_jobs=threads, backend="threading")(delayed(one)(p) for p in sample)
    dt = time.perf_counter() - t0
    work = sum(2 * len(feats[a]) * len(feats[b]) for a, b in sample)
    return work / dt, len(sample), dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    pb, feats, pairs, w = build_workload(args.workload)
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    ba_vals, mt_vals = [], []
    ba_its = 2 if args.workload == "c4" else 5
    npairs = min(len(pairs), max(cores // 2, 16))
    for _ in range(args.warmup):
        cpu_match_sample(feats, pairs, min(npairs, 8), cores)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, its, dt = cpu_ba_sample(pb, ba_its)
        ba_vals.append((v, dt))
        m, n, dtm = cpu_match_sample(feats, pairs, npairs, cores)
        mt_vals.append((m, dtm))
    total = time.perf_counter() - t_all
    ba_v = float(np.mean([v for v, _ in ba_vals]))
    mt_v = float(np.mean([v for v, _ in mt_vals]))
    line = {
        "impl": "reference", "metric": "BA observations/sec", "value": ba_v, "unit": "observations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "flush": "n/a (CPU)"},
        "cpu_baseline": {"value": ba_v, "unit": "observations/s", "cores": cores, "kind": "port",
                         "sample": "%d LM iterations of the same problem per step (restated Ceres path, OpenMP; "
                                   "Ceres itself cannot be built in this image)" % ba_its},
        "e2e": {"value": ba_v, "unit": "observations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "match": {"metric": "descriptor-pairs/sec", "value": mt_v, "unit": "descriptor-pairs/s",
                  "cpu_baseline": {"value": mt_v, "unit": "descriptor-pairs/s", "cores": cores, "kind": "reference",
                                   "sample": "%d symmetric pairs per step through cv2 BFMatcher in a joblib threading pool"
                                             % npairs},
                  "e2e": {"value": mt_v, "unit": "descriptor-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    from opensfm_b200 import _lib, bundle, dist as odist, matching

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    allreduce = None
    if world > 1:
        odist.init_from_env("nccl")
        import torch.distributed as tdist

        allreduce = "nccl"  # the library's own NCCL communicator (opensfm_b200.dist.make_allreduce is the callback form)
    L = _lib.load()
    pk = peaks()

    pb, feats, pairs, w = build_workload(args.workload)
    nobs = pb.num_observations
    sizes = {i: len(f) for i, f in enumerate(feats)}
    my_pairs = matching.shard_pairs(pairs, sizes, world)[rank] if world > 1 else pairs
    pair_work_total = sum(2 * sizes[a] * sizes[b] for a, b in pairs)
    my_pair_work = sum(2 * sizes[a] * sizes[b] for a, b in my_pairs)
    cfg = {"lowes_ratio": 0.8, "symmetric_matching": True}

    # resident descriptors for the device-timed leg
    pm = matching.PairMatcher(device=local)
    needed = sorted({i for p in my_pairs for i in p})
    for i in needed:
        pm.add(i, feats[i])

    # e2e legs: host buffers are page-locked (the contract's "pinned host memory"); every step copies
    # them to the device and reads the results back into page-locked arrays
    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    for name in ("obs_shot", "obs_point", "obs_xy", "obs_sigma", "points"):
        setattr(pb, name, pinned(getattr(pb, name)))
    feats = [pinned(f) for f in feats]
    ba_out = {"points": pinned(np.zeros((len(pb.points), 3))), "reprojection_errors": pinned(np.zeros((nobs, 3)))}

    def ba_step():
        t0 = time.perf_counter()
        res = bundle.solve(pb, device=local, rank=rank, world=world, allreduce=allreduce, out=ba_out)
        return res, time.perf_counter() - t0

    def match_resident():
        pm.submit(my_pairs, cfg["lowes_ratio"], True)
        pm.sync()
        return pm.device_ms()

    pm2 = matching.PairMatcher(device=local)  # long-lived matcher; every step re-uploads all descriptors

    def match_e2e():
        t0 = time.perf_counter()
        pm2.clear()
        pm2.add_many([(i, feats[i]) for i in needed])  # H2D of every descriptor matrix
        out = pm2.match_pairs(my_pairs, cfg)  # kernels + D2H of the match lists
        return time.perf_counter() - t0, out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        ba_step()
        match_resident()
    if args.warmup:
        match_e2e()

    launches0 = L.osfm_kernel_launch_count()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    t_begin = time.perf_counter()
    ba_dev_ms, ba_wall, ba_iters, ba_sum = 0.0, 0.0, 0, None
    mt_dev_ms, mt_kernel_ms, mt_wall = 0.0, 0.0, 0.0
    for _ in range(args.steps):
        res, dt = ba_step()
        s = res["summary"]
        ba_sum = s
        ba_dev_ms += s["time_device_ms"]
        ba_wall += dt
        ba_iters += s["iterations"]
        tot, ker = match_resident()
        mt_dev_ms += tot
        mt_kernel_ms += ker
        dte, _ = match_e2e()
        mt_wall += dte
    barrier()
    t_total = time.perf_counter() - t_begin
    clk = clocks.stop()
    launches = L.osfm_kernel_launch_count() - launches0

    # max over ranks
    vals = torch.tensor([t_total, ba_dev_ms, ba_wall, mt_dev_ms, mt_kernel_ms, mt_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        tdist.all_reduce(vals, op=tdist.ReduceOp.MAX)
    t_total, ba_dev_ms, ba_wall, mt_dev_ms, mt_kernel_ms, mt_wall = vals.tolist()

    K = args.steps
    ba_value = nobs * ba_iters / (ba_dev_ms * 1e-3)
    ba_e2e = nobs * ba_iters / ba_wall
    mt_value = pair_work_total * K / (mt_dev_ms * 1e-3)
    mt_e2e = pair_work_total * K / mt_wall

    # ---- rooflines (rank 0's kernels) ----
    s = ba_sum
    nloc = s["num_observations_local"]
    plane_bytes = s["jac_planes"] * 8
    kern = {}
    if s["schur_launches"]:
        # ba_schur reads the residual/Jacobian planes of every observation once (U, g_c) and once more
        # for the W blocks, plus the 4-byte shot index; writes V^-1 / g_p per point.
        per_launch = nloc * (plane_bytes + 4) + len(pb.points) // world * 72
        dur = s["time_schur_ms"] / s["schur_launches"] * 1e-3
        kern["ba_schur"] = dict(bytes=per_launch, ms=dur * 1e3, share=s["time_schur_ms"] / s["time_device_ms"])
    if s["linearize_launches"]:
        # ba_linearize reads the 32-byte observation record + parameters, writes the planes
        per_launch = nloc * (32 + 8 + plane_bytes)
        dur = s["time_linearize_ms"] / s["linearize_launches"] * 1e-3
        kern["ba_linearize"] = dict(bytes=per_launch, ms=dur * 1e3, share=s["time_linearize_ms"] / s["time_device_ms"])
    kern["pcg"] = dict(ms=s["time_pcg_ms"], share=s["time_pcg_ms"] / s["time_device_ms"],
                       iterations=s["pcg_iterations"], reduced_dim=s["reduced_dim"])
    # DRAM traffic per launch from the committed `ncu --set full` capture of this command (scripts/extract_traffic.py)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))

    def dram(*names):
        vals = [traffic[n]["dram_bytes_per_launch"] for n in names if n in traffic]
        return float(sum(vals)) if vals else None

    if "ba_schur" in kern:
        # the Schur phase = ba_point_blocks + ba_schur_mma (+ ba_schur for points off the fast path); it is
        # latency bound (per-segment fixed costs), the HBM figure is reported because the contract asks for it
        npts = len(pb.points) // world
        kk = nloc / max(npts, 1)
        wc_ = s["jac_planes"] / 2.0 - 4.0  # jac_planes = nres * (wc + 4), nres = 2
        fma = npts * (kk * wc_) ** 2 * 3
        kern["ba_schur"]["kernels"] = "ba_point_blocks + ba_schur_mma<wc> (fp64 mma.m8n8k4)"
        kern["ba_schur"]["fp64_tflops"] = 2.0 * fma / (kern["ba_schur"]["ms"] * 1e-3) / 1e12
        kern["ba_schur"]["fp64_note"] = "2*3*(k*wc)^2 flop per point, k = observations per point, wc = camera-side width (full square)"
    dom = max((k for k in kern if "bytes" in kern[k]), key=lambda k: kern[k]["share"])
    ach = kern[dom]["bytes"] / (kern[dom]["ms"] * 1e-3) / 1e9
    dom_traffic = dram("ba_point_blocks", "ba_schur_mma<9>", "ba_schur_mma<0>", "ba_schur") if dom == "ba_schur" else dram("ba_linearize<1>")
    roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": ach / pk["hbm"], "traffic": dom_traffic, "peak_source": pk["source"], "kernels": kern}
    flops = 2.0 * 128.0 * my_pair_work * K
    tc_ach = flops / (mt_kernel_ms * 1e-3) / 1e12
    mt_roof = {"kernel": "bf_top2_tc" if pm.last_kernel() == 2 else "bf_top2_simt", "bound": "tensor",
               "achieved": tc_ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": tc_ach / pk["bf16_sustained"],
               "traffic": dram("bf_top2_tc"), "peak_source": pk["source"] + ", sustained bf16",
               "note": "2*128 flop per descriptor pair per direction (SURVEY 8d); kernel time = distance kernel only"}

    line = None
    if rank == 0:
        h2d_ba = sum(a.nbytes for a in (pb.obs_shot, pb.obs_point, pb.obs_xy, pb.obs_sigma, pb.points, pb.inst, pb.cam_params))
        d2h_ba = pb.points.nbytes + pb.inst.nbytes + pb.cam_params.nbytes + nobs * 24
        h2d_mt = sum(feats[i].nbytes for i in needed)
        d2h_mt = sum(4 * sizes[a] for a, _ in my_pairs)
        line = {
            "metric": "BA observations/sec", "value": ba_value, "unit": "observations/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["label"], "observations": nobs, "lm_iterations_per_step": ba_iters / K,
                       "parallelism": "points sharded over %d GPU(s); pair list sharded" % world,
                       "flush": "inputs larger than L2 (Jacobian planes %.0f MB, descriptors %.0f MB)" % (
                           nobs * plane_bytes / 1e6, sum(f.nbytes for f in feats) / 1e6),
                       "loss": pb.loss_name, "termination": s["message"]},
            "ba_ms_per_step": ba_dev_ms / K, "match_ms_per_step": mt_dev_ms / K,
            "e2e": {"value": ba_e2e, "unit": "observations/s", "h2d_bytes_per_step": int(h2d_ba),
                    "d2h_bytes_per_step": int(d2h_ba)},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline,
            "match": {"metric": "descriptor-pairs/sec", "value": mt_value, "unit": "descriptor-pairs/s", "dtype": "bf16->f32",
                      "pairs": len(pairs), "descriptor_pairs_per_step": pair_work_total,
                      "e2e": {"value": mt_e2e, "unit": "descriptor-pairs/s", "h2d_bytes_per_step": int(h2d_mt),
                              "d2h_bytes_per_step": int(d2h_mt)},
                      "roofline": mt_roof},
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            os.environ.setdefault("OMP_NUM_THREADS", str(cores))
            its = 2 if args.workload == "c4" else 5
            v, it, dt = cpu_ba_sample(pb, its)
            line["cpu_baseline"] = {"value": v, "unit": "observations/s", "cores": cores, "kind": "port",
                                    "sample": "%d LM iterations of the same problem (%.1f s); restated Ceres path, OpenMP" % (it, dt)}
            npairs = min(len(pairs), max(cores // 2, 16))
            m, n, dtm = cpu_match_sample(feats, pairs, npairs, cores)
            line["match"]["cpu_baseline"] = {"value": m, "unit": "descriptor-pairs/s", "cores": cores, "kind": "reference",
                                             "sample": "%d symmetric pairs via cv2 BFMatcher, joblib threading (%.1f s)" % (n, dtm)}
        print(json.dumps(line))
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
