#!/usr/bin/env python
"""bench.py — loop-closure queries/s of the B200 path on a synthetic map (BASELINE.json metric).

One step = one batch of query keyframes through the whole hot path (project -> IMI kNN ->
covisibility voting/clustering -> PnP-RANSAC verdicts). `value` is measured with the batch
resident in HBM, `e2e` through the C-ABI with host buffers (H2D/D2H inside the timed region).
Contract: one JSON line on stdout from rank 0. See DESIGN.md "Measurement".
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

HBM_FALLBACK_GBS = 6650.0
ENTRY_BYTES = 44  # imi: 10 fp32 + int32 id


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--landmarks", type=int, default=1_000_000, help="landmarks PER GPU (weak scaling)")
    ap.add_argument("--queries", type=int, default=1000, help="query keyframes per step")
    ap.add_argument("--words", type=int, default=1000)
    ap.add_argument("--cpu-queries", type=int, default=0, help="cpu_baseline sample size (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage-report", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_world(args, rank, world):
    """Seeded synthetic map + queries + vocabulary (identical on every rank)."""
    from maplab_b200 import synthetic
    total_landmarks = args.landmarks * world
    m = synthetic.make_map(total_landmarks, seed=1)
    blob, _ = synthetic.make_vocabulary(m["bits"][:: max(len(m["bits"]) // 100_000, 1)][:100_000],
                                        num_words=args.words, seed=7)
    q = synthetic.make_queries(m, args.queries, seed=11)
    return m, blob, q


def run_b200(args):
    import torch
    import torch.distributed as dist
    from maplab_b200 import capi

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    hbm_peak, peak_src = peaks()

    t0 = time.time()
    m, blob, q = build_world(args, rank, world)
    det = capi.Detector(blob, capi.default_settings(device=local, shard_rank=rank, shard_count=world))
    frames = capi.make_frames(*(m["frames"][k] for k in ("timestamp_ns", "vertex_id", "mission_id",
                                                          "frame_index", "num_descriptors")))
    # database build: project on the GPU in chunks, insert, freeze
    n_db = len(m["bits"])
    proj = np.empty((n_db, det.dim), np.float32)
    for s in range(0, n_db, 1 << 20):
        proj[s:s + (1 << 20)] = det.project(m["bits"][s:s + (1 << 20)])
    det.insert_batch(frames, proj, m["landmarks"])
    t1 = time.time()
    det.initialize()
    torch.cuda.synchronize()
    t_build = time.time() - t1
    k = det.num_neighbors()
    nq_kf = args.queries
    qbits_h = torch.from_numpy(q["bits"]).pin_memory()
    n_q = qbits_h.shape[0]
    qbits_d = qbits_h.to(dev)
    qproj_d = torch.empty((n_q, det.dim), dtype=torch.float32, device=dev)
    idx_d = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    dist_d = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    if world > 1:
        gidx = torch.empty((world, n_q, k), dtype=torch.int32, device=dev)
        gdist = torch.empty((world, n_q, k), dtype=torch.float32, device=dev)
        midx = torch.empty_like(idx_d)
        mdist = torch.empty_like(dist_d)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device():
        det.project_device(qbits_d.data_ptr(), 64, n_q, qproj_d.data_ptr(), stream)
        det.knn_device(qproj_d.data_ptr(), n_q, k, idx_d.data_ptr(), dist_d.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(gidx, idx_d)
            dist.all_gather_into_tensor(gdist, dist_d)
            det.merge_topk_device(gidx.data_ptr(), gdist.data_ptr(), world, n_q, k, midx.data_ptr(),
                                  mdist.data_ptr(), stream)

    qbits_np = q["bits"]

    def step_e2e():
        p = det.project(qbits_np)
        return det.knn(p, k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = capi.kernel_launch_count()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    launches_per_step = (capi.kernel_launch_count() - launches0) // max(args.warmup, 3)

    # --- timed: device-resident ---
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    scan_ms = []
    barrier()
    wall0 = time.time()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # evict L2 between timed iterations
        ev[i][0].record()
        step_device()
        ev[i][1].record()
        st = det.last_scan_stats()  # syncs; reads the scan kernel's own CUDA-event time
        scan_ms.append(st["scan_ms"])
    barrier()
    wall = time.time() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = float(np.sum([a.elapsed_time(b) for a, b in ev]))
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps

    # --- timed: end to end through the C-ABI with host buffers ---
    for _ in range(2):
        step_e2e()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - e0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    st = det.last_scan_stats()
    scan_ms_avg = float(np.mean(scan_ms))
    achieved = st["algorithmic_bytes"] / (scan_ms_avg * 1e-3) / 1e9 if scan_ms_avg > 0 else 0.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {
        "metric": "loop-closure queries/sec (query keyframes fully processed per second)",
        "value": nq_kf * world / (ms_per_step * 1e-3) if False else nq_kf / (ms_per_step * 1e-3),
        "unit": "query keyframes/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8->s32 (projection), f32 (distances)", "data": "synthetic",
        "config": {"workload": f"synthetic map, {args.landmarks * world} landmarks "
                               f"({n_db} db descriptors, {len(frames)} keyframes), 512-bit FREAK, "
                               f"{nq_kf} query keyframes x 500 descriptors per step, k={k}, nw=10, "
                               f"W={args.words}x{args.words} cells",
                   "stages": ["project", "imi_knn"] + (["allgather_merge"] if world > 1 else []),
                   "l2": "256 MiB flush buffer written between timed iterations",
                   "db_build_s": round(t_build, 3)},
        "roofline": {"kernel": "imi_scan_kernel", "bound": "hbm", "achieved": achieved,
                     "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": st["algorithmic_bytes"],
                     "launch_ms": scan_ms_avg, "traffic": None},
        "e2e": {"value": nq_kf / (e2e_s / args.steps), "unit": "query keyframes/s",
                "h2d_bytes_per_step": int(qbits_np.nbytes + n_q * det.dim * 4),
                "d2h_bytes_per_step": int(n_q * det.dim * 4 + 2 * n_q * k * 4)},
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks,
        "setup_s": round(time.time() - t0, 1), "timed_wall_s": round(wall, 3),
    }
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, m, blob, q, proj, frames, k)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, m, blob, q, proj, frames, k, sample_queries=None):
    """The oracle (CPU restatement of maplab's path) on a bounded sample of the same workload."""
    from oracle import pyoracle as po
    ora = po.Engine(blob)
    t0 = time.time()
    at = 0
    lm = m["landmarks"]
    for i in range(len(frames)):
        n = int(frames["num_descriptors"][i])
        ora.insert(int(frames["timestamp_ns"][i]), int(frames["vertex_id"][i]), 0,
                   int(frames["mission_id"][i]), proj[at:at + n], lm[at:at + n])
        at += n
    t_build = time.time() - t0
    nq = sample_queries or args.cpu_queries or 40
    nq = min(nq, args.queries)
    nd = int(np.sum(q["frames"]["num_descriptors"][:nq]))
    t1 = time.time()
    qp = ora.project(q["bits"][:nd])
    ora.knn(qp, k)
    dt = time.time() - t1
    return {"value": nq / dt, "unit": "query keyframes/s", "cores": 1, "kind": "port",
            "sample": f"first {nq} of {args.queries} query keyframes ({nd} descriptors): project + kNN "
                      f"on the full database, single thread; oracle db build {t_build:.1f}s not included"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    m, blob, q = build_world(args, 0, max(int(os.environ.get("WORLD_SIZE", 1)), 1))
    from maplab_b200 import capi
    from oracle import pyoracle as po
    frames = capi.make_frames(*(m["frames"][k] for k in ("timestamp_ns", "vertex_id", "mission_id",
                                                          "frame_index", "num_descriptors")))
    ora = po.Engine(blob)
    proj = ora.project(m["bits"])
    k = 6 if len(proj) < 1e7 else 8
    cb = cpu_baseline(args, m, blob, q, proj, frames, k, sample_queries=20)
    out = {"impl": "reference", "metric": "loop-closure queries/sec (query keyframes fully processed per second)",
           "value": cb["value"], "unit": "query keyframes/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "data": "synthetic", "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "query keyframes/s", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
