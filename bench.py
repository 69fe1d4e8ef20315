#!/usr/bin/env python
"""Benchmark of the composition hot path (BASELINE.json metric) — prints ONE JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload merge]

Workload ``merge`` (default, BASELINE config 2 / SURVEY §8(d) C2): the 3-way vicuna-7B-shaped merge
``out = Σ_m w_m · src_m`` with w = (0.333, 0.333, 0.333) over 291 bf16 tensors (6,738,415,616 elements per
source, random-init N(0, 0.02) generated on the device), sharded BY PARAMETER TENSOR over the N ranks (greedy
size balancing, no collective).  A step = one pass of the merge kernel over the rank's whole shard.
``value`` = whole-job algorithmic GB/s = (3 reads + 1 write) x 2 B x elements of ALL ranks / max-over-ranks step time.

``--impl reference``: the reference's CPU torch merge arithmetic (oracle port of
merge_unimodal_modelcompose.py:105-112 in its weighted form) on the host's cores, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WEIGHTS = (0.333, 0.333, 0.333)
SOURCE_SEEDS = (1000, 1001, 1002)  # video, audio, vision (README.md:86-91 order)
L2_BYTES = 126 * 1024 * 1024


# ------------------------------------------------------------------------------------------------ helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------ merge workload
def merge_shard(world: int, rank: int):
    from modelcompose_b200 import synthetic as syn
    shapes = syn.dense_7b_tensor_shapes()
    sizes = [int(torch.Size(s).numel()) for _, s in shapes]
    mine = syn.shard_tensors_greedy(sizes, world)[rank]
    return shapes, sizes, mine


def make_device_sources(shapes, mine, device):
    """3 sources x this rank's tensors, bf16 N(0, 0.02), one allocation per tensor (as a loaded checkpoint has)."""
    srcs = []
    for m, seed in enumerate(SOURCE_SEEDS):
        g = torch.Generator(device=device).manual_seed(seed)
        lst = []
        for i in mine:
            t = torch.empty(shapes[i][1], dtype=torch.bfloat16, device=device)
            t.normal_(0.0, 0.02, generator=g)
            lst.append(t)
        srcs.append(lst)
    return srcs


def cpu_merge_rate(sample_tensors, min_seconds: float, max_passes: int = 50):
    """Times the oracle port of the reference CPU merge arithmetic on host cores; returns (GB/s, seconds, passes)."""
    from oracle import merge_oracle as MO
    nbytes = sum(t[0].numel() for t in sample_tensors) * 2 * (len(WEIGHTS) + 1)
    MO.weighted_merge(sample_tensors[0], WEIGHTS)  # warm the allocator / thread pool
    t0 = time.perf_counter()
    passes = 0
    while True:
        for ts in sample_tensors:
            MO.weighted_merge(ts, WEIGHTS)
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or passes >= max_passes:
            break
    return nbytes * passes / dt / 1e9, dt, passes


def cpu_sample_tensors():
    """One decoder layer of each source (202,383,360 elements: q,k,v,o,gate,up,down + 2 norms), seeded on CPU."""
    from modelcompose_b200 import synthetic as syn
    shapes = [s for n, s in syn.dense_7b_tensor_shapes() if n.startswith("model.layers.0.")]
    out = []
    for shp in shapes:
        ts = []
        for seed in SOURCE_SEEDS:
            g = torch.Generator().manual_seed(seed + 7)
            # cheap deterministic fill (randn over 200M elements x3 would dominate the bench's wall time)
            base = torch.randn(4096, generator=g) * 0.02
            t = base.repeat((int(torch.Size(shp).numel()) + 4095) // 4096)[: int(torch.Size(shp).numel())]
            ts.append(t.to(torch.bfloat16).reshape(shp))
        out.append(ts)
    return out, "one vicuna-7B decoder layer x 3 sources (202,383,360 elements per source, 1.62 GB algorithmic)"


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    sample, desc = cpu_sample_tensors()
    from oracle import merge_oracle as MO
    nbytes = sum(t[0].numel() for t in sample) * 2 * (len(WEIGHTS) + 1)
    for _ in range(max(args.warmup, 1)):
        for ts in sample:
            MO.weighted_merge(ts, WEIGHTS)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for ts in sample:
            MO.weighted_merge(ts, WEIGHTS)
    dt = time.perf_counter() - t0
    gbs = nbytes * args.steps / dt / 1e9
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "3x7B merge GB/s", "value": round(gbs, 3), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "3-way vicuna-7B-shaped merge video=0.333,audio=0.333,vision=0.333 (C2)",
                   "step": "bounded sample: " + desc, "l2": "sample larger than L2/LLC"},
        "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": "port", "sample": desc,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_merge(args):
    import torch.distributed as dist
    from modelcompose_b200 import merge as M
    rank, local_rank, world = dist_env()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    if args.emulate_world > 1:  # profiling aid: rank 0's shard of a K-way job in one process (never a bench value)
        shapes, sizes, mine = merge_shard(args.emulate_world, 0)
    else:
        shapes, sizes, mine = merge_shard(world, rank)
    srcs = make_device_sources(shapes, mine, device)
    outs = [torch.empty(shapes[i][1], dtype=torch.bfloat16, device=device) for i in mine]
    plan = M.MergePlan(srcs, outs, tuning=args.tuning)
    my_bytes = plan.algorithmic_bytes
    total_bytes = sum(sizes) * 2 * (len(WEIGHTS) + 1)
    if args.emulate_world > 1:
        total_bytes = my_bytes

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        plan.run(WEIGHTS)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        t_start.record()
        for a, b in ev:
            a.record()
            plan.run(WEIGHTS)
            b.record()
        t_end.record()
        barrier()
    total_ms = t_start.elapsed_time(t_end)
    launch_ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    ms_per_step = max_ms / args.steps
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    # parity spot-check outside the timed region: smallest and one mid-size tensor vs the CPU oracle
    from oracle import merge_oracle as MO
    order = sorted(range(len(mine)), key=lambda j: sizes[mine[j]])
    for j in (order[0], order[len(order) // 2]):
        want = MO.weighted_merge([srcs[s][j].cpu() for s in range(3)], WEIGHTS)
        if not torch.equal(outs[j].cpu().view(torch.int16), want.view(torch.int16)):
            raise SystemExit(f"PARITY FAILURE on tensor {shapes[mine[j]][0]}")

    # ---- e2e: same merge through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timing)
    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"profiling_only": True, "ms_per_step": round(ms_per_step, 4), "GBps": round(value, 1),
                              "launch_ms": round(launch_ms, 4), "emulate_world": args.emulate_world}), flush=True)
        return
    e2e = run_merge_e2e(args, M, srcs, outs, shapes, sizes, mine, device, world, barrier, dist)

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = my_bytes / (launch_ms * 1e-3) / 1e9
        cpu_sample, desc = cpu_sample_tensors()
        cpu_gbs, cpu_s, passes = cpu_merge_rate(cpu_sample, min_seconds=10.0)
        line = {
            "metric": "3x7B merge GB/s", "value": round(value, 2), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "3-way vicuna-7B-shaped merge video=0.333,audio=0.333,vision=0.333 (C2)",
                       "tensors": len(sizes), "elements_per_source": sum(sizes), "sharding": f"by-tensor greedy x{world}",
                       "algorithmic_bytes": total_bytes, "l2": "inputs larger than L2 (%.2f GB per GPU vs 0.13 GB)" % (my_bytes / 1e9),
                       "tuning": args.tuning},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                         "kernel": "mc::merge_kernel<3,bf16,bf16>", "launch_ms": round(launch_ms, 4),
                         "frac_of_8TBps_nominal": round(achieved / 8000.0, 4)},
            "cpu_baseline": {"value": round(cpu_gbs, 3), "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{desc}, {passes} passes in {cpu_s:.1f} s", "host_cpus": os.cpu_count()},
            "e2e": e2e,
            "gpu_launches": args.steps * world,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_merge_e2e(args, M, srcs, outs, shapes, sizes, mine, device, world, barrier, dist):
    """Whole-shard merge through ``mc_merge_host`` from pinned host buffers; value = all ranks' bytes / max time."""
    import psutil
    n_src = len(srcs)
    my_elems = sum(sizes[i] for i in mine)
    need = my_elems * 2 * (n_src + 1)
    avail = psutil.virtual_memory().available / max(world, 1) * (1 if world == 1 else 1)
    # keep pinned memory well below what the host has: use every k-th tensor when RAM is short
    stride = 1
    while need / stride > 0.45 * psutil.virtual_memory().available / world and stride < 64:
        stride *= 2
    pick = list(range(0, len(mine), stride))
    h_src = [[torch.empty(srcs[s][j].shape, dtype=torch.bfloat16).pin_memory() for j in pick] for s in range(n_src)]
    for s in range(n_src):
        for hj, j in zip(h_src[s], pick):
            hj.copy_(srcs[s][j])
    elems = sum(srcs[0][j].numel() for j in pick)
    h2d, d2h = elems * 2 * n_src, elems * 2
    h_out = [torch.empty(srcs[0][j].shape, dtype=torch.bfloat16).pin_memory() for j in pick]
    from modelcompose_b200 import _cabi
    lib = _cabi.lib()
    sp = _cabi.ptr_array([h_src[s][k].data_ptr() for s in range(n_src) for k in range(len(pick))])
    dp = _cabi.ptr_array([o.data_ptr() for o in h_out])
    ne = _cabi.i64_array([o.numel() for o in h_out])
    w = _cabi.f32_array(WEIGHTS)

    def step():
        _cabi.check(lib.mc_merge_host(len(pick), n_src, sp, dp, ne, w, _cabi.MC_MERGE_WEIGHTED, _cabi.MC_BF16,
                                      _cabi.MC_BF16, 0), "mc_merge_host")
    steps = max(1, min(args.steps, 3))
    step()  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()  # returns after the last D2H byte landed (synchronous contract)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=device)
    b = torch.tensor([float(elems * 2 * (n_src + 1))], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
    ok = all(torch.equal(h_out[k].view(torch.int16), outs[j].cpu().view(torch.int16)) for k, j in
             list(zip(range(len(pick)), pick))[:3])
    if not ok:
        raise SystemExit("PARITY FAILURE: host-streamed merge differs from the device-resident merge")
    return {"value": round(float(b.item()) / (float(t.item()) / steps) / 1e9, 2), "unit": "GB/s",
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": steps,
            "api": "mc_merge_host (pinned host buffers, H2D/kernel/D2H pipelined, returns after last D2H)",
            "sample": "whole shard" if stride == 1 else f"every {stride}th tensor of the shard (host RAM bound)",
            "timer": "host wall clock around the synchronous call, max over ranks"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="merge", choices=["merge"])
    ap.add_argument("--tuning", type=int, default=0)
    ap.add_argument("--emulate-world", type=int, default=1,
                    help="profiling aid: run rank 0's shard of a K-way job on one GPU (ncu captures)")
    ap.add_argument("--no-e2e", action="store_true", help="profiling aid: skip the host-buffer e2e and CPU legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)
    run_merge(args)


if __name__ == "__main__":
    main()
