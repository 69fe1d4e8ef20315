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


def measured_traffic(kernel: str, world: int, launch_bytes: int):
    """DRAM bytes per launch from the committed ncu --set full captures of this kernel (profiles/traffic.json).  A capture of
    exactly this launch (same algorithmic bytes) is returned as measured; otherwise the capture's traffic / algorithmic ratio
    (1.000 - 1.012 for every capture so far) is applied to this launch's algorithmic bytes.  None when the kernel has no capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        d = json.load(f)
    best = None
    for k, v in d.items():
        if not k.startswith(kernel) or "dram_bytes" not in v or "algorithmic_bytes" not in v:
            continue
        if v["algorithmic_bytes"] == launch_bytes:
            return v["dram_bytes"]
        gap = abs(v["algorithmic_bytes"] - launch_bytes)
        if best is None or gap < best[0]:
            best = (gap, v["dram_bytes"] / v["algorithmic_bytes"])
    return None if best is None else int(round(best[1] * launch_bytes))


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------ merge workload
MERGE_CONFIGS = {
    # name: (description, source seeds, weights, kernel label)
    "c2": ("3-way vicuna-7B-shaped merge video=0.333,audio=0.333,vision=0.333 (C2)", (1000, 1001, 1002), (0.333, 0.333, 0.333)),
    "n4": ("4-way vicuna-7B-shaped merge video,audio,vision,point at 0.25 each (MCUB-4 checkpoints of C5)", (1000, 1001, 1002, 1003),
           (0.25, 0.25, 0.25, 0.25)),
    "c2b": ("base + 3 materialised blend W_eff = (1 - 0.999) W_base + 0.333 (video + audio + vision) (C2b, "
            "convert_to_multimodal.py:111-113 form)", (1003, 1000, 1001, 1002), (1.0 - 0.999, 0.333, 0.333, 0.333)),
}
SPLIT_ROWS_ABOVE = 1 << 26   # elements: embed_tokens / lm_head (131 M each) are cut into row slices when sharding


def merge_items(world: int):
    """The job's work items: (name, shape) of every tensor; for world > 1 the two 131 M-element matrices are cut into `world`
    row slices first (the merge is elementwise, so any row range is an independent item) — by whole tensors the heaviest
    rank carried 6.80 GB against 6.74 ideal at 8 ranks."""
    from modelcompose_b200 import synthetic as syn
    items = []
    for name, shape in syn.dense_7b_tensor_shapes():
        n = int(torch.Size(shape).numel())
        if world > 1 and n > SPLIT_ROWS_ABOVE and len(shape) == 2:
            rows = shape[0]
            for k in range(world):
                r0, r1 = rows * k // world, rows * (k + 1) // world
                items.append((f"{name}[{r0}:{r1}]", (r1 - r0, shape[1])))
        else:
            items.append((name, tuple(shape)))
    return items


def merge_shard(world: int, rank: int):
    from modelcompose_b200 import synthetic as syn
    shapes = merge_items(world)
    sizes = [int(torch.Size(s).numel()) for _, s in shapes]
    mine = syn.shard_tensors_greedy(sizes, world)[rank]
    return shapes, sizes, mine


def make_device_source(shapes, mine, device, seed):
    """This rank's tensors of one checkpoint: bf16 N(0, 0.02), one allocation per tensor (as a loaded checkpoint has)."""
    g = torch.Generator(device=device).manual_seed(seed)
    lst = []
    for i in mine:
        t = torch.empty(shapes[i][1], dtype=torch.bfloat16, device=device)
        t.normal_(0.0, 0.02, generator=g)
        lst.append(t)
    return lst


def cpu_sample_tensors(n_src: int = 3):
    """One decoder layer of each source (202,383,360 elements: q,k,v,o,gate,up,down + 2 norms), seeded on CPU."""
    from modelcompose_b200 import synthetic as syn
    shapes = [s for n, s in syn.dense_7b_tensor_shapes() if n.startswith("model.layers.0.")]
    out = []
    for shp in shapes:
        ts = []
        for seed in range(1000, 1000 + n_src):
            g = torch.Generator().manual_seed(seed + 7)
            # cheap deterministic fill (randn over 200M elements x3 would dominate the bench's wall time)
            base = torch.randn(4096, generator=g) * 0.02
            t = base.repeat((int(torch.Size(shp).numel()) + 4095) // 4096)[: int(torch.Size(shp).numel())]
            ts.append(t.to(torch.bfloat16).reshape(shp))
        out.append(ts)
    return out, "one vicuna-7B decoder layer x 3 sources (202,383,360 elements per source, 1.62 GB algorithmic)"


CPU_FORM = "mean"   # the reference arm's headline form


def CPU_FORMS(MO, w):
    """The reference's CPU merge arithmetic, as callables over the list of one tensor per source:
    `mean`     — the reference CLI's own equal-weight merge (merge_unimodal_modelcompose.py:109-112: Python sum() of the bf16
                 tensors, every add rounded to bf16, then `/ len`): for three sources the closest thing the reference has to
                 the 0.333 / 0.333 / 0.333 blend, and the form the CPU arm is quoted on;
    `sum`      — :105-108, the same without the division;
    `weighted` — the materialised online-merge-reset blend itself (multimodal_llama.py:130-149 coefficients applied to full
                 weights, fp32 temporaries), oracle/merge_oracle.weighted_merge: bit for bit what the GPU arm's headline computes."""
    return (("mean", lambda ts: MO.ref_sum(ts) / len(ts)), ("sum", lambda ts: MO.ref_sum(ts)),
            ("weighted", lambda ts: MO.weighted_merge(ts, w)))


def cpu_merge_rates(sample_tensors, min_seconds: float, max_passes: int = 50):
    """Every form of CPU_FORMS on the host's cores, as GB/s of (3 reads + 1 write) x 2 B."""
    from oracle import merge_oracle as MO
    w = MERGE_CONFIGS["c2"][2]
    nbytes = sum(t[0].numel() for t in sample_tensors) * 2 * (len(w) + 1)
    out = {}
    for name, fn in CPU_FORMS(MO, w):
        fn(sample_tensors[0])  # warm the allocator / thread pool
        t0 = time.perf_counter()
        passes = 0
        while True:
            for ts in sample_tensors:
                fn(ts)
            passes += 1
            dt = time.perf_counter() - t0
            if dt >= min_seconds or passes >= max_passes:
                break
        out[name] = (nbytes * passes / dt / 1e9, dt, passes)
    return out


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)  # torchrun pins OMP_NUM_THREADS=1; rank 0 alone works here, on every host core
    sample, desc = cpu_sample_tensors()
    from oracle import merge_oracle as MO
    w = MERGE_CONFIGS["c2"][2]
    nbytes = sum(t[0].numel() for t in sample) * 2 * (len(w) + 1)
    rates = {}
    for name, fn in CPU_FORMS(MO, w):
        for _ in range(max(args.warmup, 1)):
            for ts in sample:
                fn(ts)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for ts in sample:
                fn(ts)
        rates[name] = (time.perf_counter() - t0) / args.steps
    best = CPU_FORM
    dt = rates[best]
    gbs = nbytes / dt / 1e9
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "3x7B merge GB/s", "value": round(gbs, 3), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": MERGE_CONFIGS["c2"][0], "step": "bounded sample: " + desc, "l2": "sample larger than L2/LLC",
                   "cpu_form": best},
        "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": "port", "sample": desc,
                         "host_cpus": os.cpu_count(), "form": best,
                         "forms_GBps": {k: round(nbytes / v / 1e9, 3) for k, v in rates.items()},
                         "note": "the reference is pure Python and /root/reference does not travel to the GPU box: "
                                 "oracle/merge_oracle.py restates merge_unimodal_modelcompose.py:105-112 (`sum`, `mean`) and the materialised blend"},
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.workload in ("all", "prefill"):
        pre = cpu_prefill_layer_rate(args.prefill_config)
        line["prefill"] = {"impl": "reference", "metric": "composed-prefill tokens/s", "value": pre["value"], "unit": "tokens/s",
                           "config": {"workload": PREFILL_CONFIGS[args.prefill_config][0]}, "cpu_baseline": pre,
                           "e2e": {"value": pre["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def setup_dist(args):
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # one process per GPU: stage host buffers on the GPU's own NUMA node (at N = 1 the rank keeps every core: it also
        # runs the CPU baseline)
        from modelcompose_b200 import _cabi
        _cabi.bind_host_thread_to_gpu(local_rank)
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    return dist, rank, world, device, barrier


class MergeJob:
    """Device-resident sources of this rank's shard (generated once, shared by the C2 / 4-way / base+N lines), the output
    tensors, and — for the e2e legs — one pinned host ARENA per checkpoint holding the same tensors back to back."""

    def __init__(self, args, world, rank, device):
        if args.emulate_world > 1:  # profiling aid: rank 0's shard of a K-way job in one process (never a bench value)
            self.shapes, self.sizes, self.mine = merge_shard(args.emulate_world, 0)
        else:
            self.shapes, self.sizes, self.mine = merge_shard(world, rank)
        self.device, self.world, self.rank = device, world, rank
        self.src = {}
        self.outs = [torch.empty(self.shapes[i][1], dtype=torch.bfloat16, device=device) for i in self.mine]
        self.h_src, self.h_out, self.pick = {}, None, None
        self.total_elems = sum(self.sizes)
        self.emulated = args.emulate_world > 1

    def source(self, seed):
        if seed not in self.src:
            self.src[seed] = make_device_source(self.shapes, self.mine, self.device, seed)
        return self.src[seed]

    # ---- pinned host copies for the e2e leg
    def host_pick(self, n_src):
        """Every k-th tensor of the shard when host RAM is short (keeps pinned memory below 45 % of what the host has)."""
        if self.pick is None:
            import psutil
            need = sum(self.sizes[i] for i in self.mine) * 2 * (n_src + 1)
            stride = 1
            while need / stride > 0.45 * psutil.virtual_memory().available / self.world and stride < 64:
                stride *= 2
            self.pick, self.stride = list(range(0, len(self.mine), stride)), stride
        return self.pick

    def _arena(self, numels):
        """One pinned allocation, carved into per-tensor views at 256-byte aligned offsets."""
        offs, total = [], 0
        for n in numels:
            offs.append(total)
            total += (n * 2 + 255) // 256 * 256
        arena = torch.empty(max(total, 256), dtype=torch.uint8).pin_memory()
        return arena, [arena[o:o + n * 2].view(torch.bfloat16) for o, n in zip(offs, numels)]

    def host_source(self, seed):
        if seed not in self.h_src:
            dev = self.source(seed)
            arena, views = self._arena([dev[j].numel() for j in self.pick])
            for v, j in zip(views, self.pick):
                v.copy_(dev[j].view(-1))
            self.h_src[seed] = (arena, views)
        return self.h_src[seed][1]

    def host_out(self):
        if self.h_out is None:
            self.h_out = self._arena([self.outs[j].numel() for j in self.pick])
        return self.h_out[1]


def run_merge(args, dist, rank, world, device, barrier, job: MergeJob, cfg_name: str, with_cpu: bool):
    from modelcompose_b200 import merge as M
    desc, seeds, weights = MERGE_CONFIGS[cfg_name]
    n_src = len(seeds)
    local_rank = device.index
    srcs = [job.source(sd) for sd in seeds]
    outs, shapes, sizes, mine = job.outs, job.shapes, job.sizes, job.mine
    plan = M.MergePlan(srcs, outs, tuning=args.tuning)
    my_bytes = plan.algorithmic_bytes
    total_bytes = job.total_elems * 2 * (n_src + 1)
    if job.emulated:
        total_bytes = my_bytes

    for _ in range(args.warmup):
        plan.run(weights)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        t_start.record()
        for a, b in ev:
            a.record()
            plan.run(weights)
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
        want = MO.weighted_merge([srcs[s][j].cpu() for s in range(n_src)], weights)
        if not torch.equal(outs[j].cpu().view(torch.int16), want.view(torch.int16)):
            raise SystemExit(f"PARITY FAILURE ({cfg_name}) on tensor {shapes[mine[j]][0]}")

    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"profiling_only": True, "config": cfg_name, "ms_per_step": round(ms_per_step, 4), "GBps": round(value, 1),
                              "launch_ms": round(launch_ms, 4), "emulate_world": args.emulate_world}), flush=True)
        plan.close()
        return None
    # ---- e2e: same merge through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timing)
    e2e = run_merge_e2e(args, job, seeds, weights, device, world, barrier, dist, probe=(cfg_name == "c2"))

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = my_bytes / (launch_ms * 1e-3) / 1e9
        cpu_baseline = None  # reported at N=1 only, on the primary line
        if with_cpu and world == 1:
            cpu_sample, sdesc = cpu_sample_tensors()
            rates = cpu_merge_rates(cpu_sample, min_seconds=4.0)
            best = CPU_FORM
            cpu_baseline = {"value": round(rates[best][0], 3), "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "port",
                            "form": best, "forms_GBps": {k: round(v[0], 3) for k, v in rates.items()},
                            "sample": f"{sdesc}; `mean` = the reference CLI's own equal-weight merge (merge_unimodal_modelcompose.py:109-112), "
                                      f"`weighted` = the materialised blend the GPU arm computes; {rates[best][2]} passes in {rates[best][1]:.1f} s",
                            "host_cpus": os.cpu_count()}
        kern = f"mc::merge_kernel<{n_src},bf16,bf16>"
        line = {
            "metric": "3x7B merge GB/s" if cfg_name == "c2" else f"{n_src}x7B merge GB/s", "value": round(value, 2), "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "tensors": len(sizes), "elements_per_source": job.total_elems, "weights": list(weights),
                       "sharding": f"greedy by work item x{world} (tensors; embed_tokens / lm_head cut into {world} row slices)" if world > 1
                       else "one GPU holds every tensor",
                       "algorithmic_bytes": total_bytes, "l2": "inputs larger than L2 (%.2f GB per GPU vs 0.13 GB)" % (my_bytes / 1e9),
                       "tuning": args.tuning},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": measured_traffic(kern, world, my_bytes),
                         "traffic_unit": "bytes per launch: DRAM read+write bytes of the committed ncu --set full capture of this kernel "
                                         "(profiles/traffic.json), scaled by this launch's algorithmic bytes when the capture ran another shard",
                         "peak_source": peak_src, "kernel": kern, "launch_ms": round(launch_ms, 4),
                         "frac_of_8TBps_nominal": round(achieved / 8000.0, 4)},
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": args.steps * world,
            "clocks": clocks.summary(),
        }
    plan.close()
    return line


def run_merge_e2e(args, job: MergeJob, seeds, weights, device, world, barrier, dist, probe: bool):
    """Whole-shard merge through ``mc_merge_host`` from pinned host arenas; value = all ranks' bytes / max time."""
    from modelcompose_b200 import _cabi
    n_src = len(seeds)
    # the staging buffers belong on the GPU's own NUMA node (the rank is bound for this leg only: at N = 1 it also runs the
    # CPU baseline on every core afterwards)
    saved_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_bound = _cabi.bind_host_thread_to_gpu(device.index)
    pick = job.host_pick(max(n_src, 4))
    h_src = [job.host_source(sd) for sd in seeds]
    h_out = job.host_out()
    srcs = [job.source(sd) for sd in seeds]
    outs = job.outs
    elems = sum(o.numel() for o in h_out)
    h2d, d2h = elems * 2 * n_src, elems * 2
    lib = _cabi.lib()
    sp = _cabi.ptr_array([h_src[s][k].data_ptr() for s in range(n_src) for k in range(len(pick))])
    dp = _cabi.ptr_array([o.data_ptr() for o in h_out])
    ne = _cabi.i64_array([o.numel() for o in h_out])
    w = _cabi.f32_array(weights)

    def step():
        _cabi.check(lib.mc_merge_host(len(pick), n_src, sp, dp, ne, w, _cabi.MC_MERGE_WEIGHTED, _cabi.MC_BF16,
                                      _cabi.MC_BF16, 0), "mc_merge_host")

    def pcie_probe():
        """Plain pinned-memory copy rates of this box (what bounds the e2e number): 1 GiB each way, CUDA events; then both
        directions at once in the merge's 3 : 1 byte ratio, and one cold pass over a whole pinned arena.  With N ranks the
        probes of all ranks run at the same time (barrier first), so the numbers are what N concurrent links give."""
        big = max(range(len(pick)), key=lambda k: h_src[0][k].numel())
        hbuf, dbuf = h_src[0][big].view(-1), srcs[0][pick[big]].view(-1).clone()
        reps = max(1, (1 << 30) // (hbuf.numel() * 2))
        out = {}
        barrier()
        for name, (dst, src) in (("h2d_GBps", (dbuf, hbuf)), ("d2h_GBps", (h_out[big].view(-1), dbuf))):
            dst.copy_(src, non_blocking=True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
            b.record()
            torch.cuda.synchronize()
            out[name] = round(reps * hbuf.numel() * 2 / (a.elapsed_time(b) * 1e-3) / 1e9, 1)
        try:
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            dbuf2 = dbuf.clone()
            barrier()
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            s_in.wait_event(a)
            s_out.wait_event(a)
            with torch.cuda.stream(s_in):
                for _ in range(3 * reps):
                    dbuf.copy_(hbuf, non_blocking=True)
                b.record(s_in)
            with torch.cuda.stream(s_out):
                for _ in range(reps):
                    h_out[big].view(-1).copy_(dbuf2, non_blocking=True)
                c.record(s_out)
            torch.cuda.synchronize()
            out["h2d_GBps_while_d2h"] = round(3 * reps * hbuf.numel() * 2 / (a.elapsed_time(b) * 1e-3) / 1e9, 1)
            out["d2h_GBps_while_h2d"] = round(reps * hbuf.numel() * 2 / (a.elapsed_time(c) * 1e-3) / 1e9, 1)
            # the whole arena of source 0 once, as one copy: a cold stream (what the merge does) rather than a warm 1 GiB loop
            arena = job.h_src[seeds[0]][0]
            scratch = torch.empty(arena.numel(), dtype=torch.uint8, device=device)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            scratch.copy_(arena, non_blocking=True)
            b.record()
            torch.cuda.synchronize()
            out["h2d_GBps_cold_stream"] = round(arena.numel() / (a.elapsed_time(b) * 1e-3) / 1e9, 1)
            del scratch
        except Exception as exc:  # a probe must never cost the bench line
            out["bidirectional_probe_error"] = str(exc)[:80]
        return out
    probe_res = pcie_probe() if probe else None
    if probe:
        job.probe = probe_res
    steps = max(1, min(args.steps, 5 if probe else 3))
    # warm-up: the first passes over freshly pinned arenas run 20 - 35 % below the steady rate on some boxes (the 4-source pass that
    # follows the 3-source one in the same run, over the same arenas, is then FASTER although it moves more bytes): at least 3
    # passes, then until two consecutive passes agree within 3 %, at most 8 (every rank runs the same count: the stop is agreed)
    prev, warm = None, 0
    while warm < 8:
        tw = time.perf_counter()
        step()
        dt_w = time.perf_counter() - tw
        warm += 1
        if args.e2e_trace:
            print(f"e2e warm-up pass {warm}: {dt_w:.3f} s", file=sys.stderr, flush=True)
        settled = warm >= 3 and prev is not None and abs(dt_w - prev) <= 0.03 * prev
        if world > 1:
            flag = torch.tensor([1.0 if settled else 0.0], dtype=torch.float64, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            settled = bool(flag.item() > 0.5)
        if settled:
            break
        prev = dt_w
    barrier()
    pass_s = []
    t0 = time.perf_counter()
    for _ in range(steps):
        tw = time.perf_counter()
        step()  # returns after the last D2H byte landed (synchronous contract)
        pass_s.append(time.perf_counter() - tw)
        if args.e2e_trace:
            print(f"e2e timed pass: {pass_s[-1]:.3f} s", file=sys.stderr, flush=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=device)
    b = torch.tensor([float(elems * 2 * (n_src + 1))], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
    ok = all(torch.equal(h_out[k].view(torch.int16), outs[j].cpu().view(-1).view(torch.int16)) for k, j in
             list(zip(range(len(pick)), pick))[:3])
    if saved_affinity is not None and world == 1:
        os.sched_setaffinity(0, saved_affinity)
    if not ok:
        raise SystemExit("PARITY FAILURE: host-streamed merge differs from the device-resident merge")
    pr = getattr(job, "probe", None) or {}
    res = {"value": round(float(b.item()) / (float(t.item()) / steps) / 1e9, 2), "unit": "GB/s",
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": steps,
           "api": "mc_merge_host (one pinned host arena per checkpoint, H2D/kernel/D2H pipelined, returns after last D2H)",
           "sample": "whole shard" if job.stride == 1 else f"every {job.stride}th tensor of the shard (host RAM bound)",
           "timer": "host wall clock around the synchronous call, max over ranks", "warmup_passes": warm,
           "pass_seconds": [round(x, 3) for x in pass_s],
           "pass_note": "value = bytes of all timed passes / their total time; single passes vary on shared hosts (rank 0's passes listed)",
           "host_numa_bound": bool(numa_bound)}
    if pr:
        res["pcie_probe"] = pr
        res["pcie_probe_note"] = f"{world} rank(s) probing at the same time: rates per GPU"
        h2d_rate = pr.get("h2d_GBps_while_d2h", pr["h2d_GBps"])
        res["pcie_bound_GBps"] = round(float(elems * 2 * (n_src + 1)) / (h2d / (h2d_rate * 1e9)) / 1e9 * world, 1)
        res["frac_of_pcie_bound"] = round(res["value"] / res["pcie_bound_GBps"], 3)
    return res


# ------------------------------------------------------------------------------------------------ TIES workload
def run_ties(device, steps: int = 10, K: int = 20, func: str = "mean", e2e: bool = True):
    """TIES merge (`--strategy ties-mean -K 20`, reference ties_merging.py:161-179) of the shared `default` adapters of
    three vicuna-7B DAMC checkpoints: 3 sources x 448 LoRA tensors = 319,815,680 bf16 elements per source, N(0, 0.02) /
    U(+-1/sqrt(in)) random init on the device.  One GPU (the trim threshold and the majority sign are global statistics).
    Parity outside the timed region: exact rank of every threshold (torch counts on the device) and three tensors
    re-derived by the CPU oracle from those statistics."""
    from modelcompose_b200 import merge as M
    from modelcompose_b200 import synthetic as syn
    from oracle import ties_oracle as TO
    llama, r = syn.VICUNA_7B, 128
    shapes = []
    for _ in range(llama["num_hidden_layers"]):
        for name in syn.LINEAR_NAMES:
            out_f, in_f = syn.linear_shape(llama, name)
            shapes += [(r, in_f), (out_f, r)]
    srcs = []
    for s in range(3):
        g = torch.Generator(device=device).manual_seed(3000 + s)
        lst = []
        for (a, b) in shapes:
            if a == r:   # lora_A ~ U(+-1/sqrt(in)) (peft init)
                t = (torch.rand((a, b), generator=g, device=device) * 2 - 1) / (b ** 0.5)
            else:        # lora_B ~ N(0, 0.02)
                t = torch.randn((a, b), generator=g, device=device) * 0.02
            lst.append(t.to(torch.bfloat16))
        srcs.append(lst)
    outs = [torch.empty(sh, dtype=torch.float32 if func == "mean" else torch.bfloat16, device=device) for sh in shapes]
    plan = M.TiesPlan(srcs, outs)
    for _ in range(3):
        plan.run(K, func)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        plan.run(K, func)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    st = plan.stats()
    d, k = plan.elements, M.ties_kth_rank(plan.elements, K)
    for s in range(3):
        thr = st["thresholds"][s]
        below = sum(int((t.abs().float() < thr).sum()) for t in srcs[s])
        upto = sum(int((t.abs().float() <= thr).sum()) for t in srcs[s])
        if not below < k <= upto:
            raise SystemExit(f"PARITY FAILURE: TIES threshold of source {s} is not the k-th smallest magnitude")
    if st["n_pos"] + st["n_neg"] + st["n_zero"] + st["n_ambiguous"] != d or st["majority"] != (st["n_pos"] > st["n_neg"]) - (st["n_pos"] < st["n_neg"]):
        raise SystemExit("PARITY FAILURE: TIES sign census inconsistent")
    for j in (0, 1, len(shapes) - 1):
        flat = torch.stack([srcs[s][j].cpu().reshape(-1) for s in range(3)])
        want = TO.merge_given_statistics(flat, st["thresholds"], st["majority"], func)
        got = outs[j].cpu().reshape(-1)
        if got.dtype != want.dtype or not torch.equal(got.view(torch.int32 if got.dtype == torch.float32 else torch.int16),
                                                     want.view(torch.int32 if want.dtype == torch.float32 else torch.int16)):
            raise SystemExit(f"PARITY FAILURE: TIES output tensor {j} differs from the oracle")
    peak, peak_src = measured_peaks()
    gbs = plan.algorithmic_bytes / (ms * 1e-3) / 1e9
    res = {"metric": "TIES merge GB/s (ties-%s, K=%d)" % (func, K), "value": round(gbs, 1), "unit": "GB/s", "n_gpus": 1, "steps": steps,
           "ms_per_step": round(ms, 4), "dtype": "bf16", "data": "synthetic",
           "config": {"workload": "ties-%s of the `default` LoRA adapters of 3 vicuna-7B DAMC checkpoints" % func, "tensors": len(shapes),
                      "elements_per_source": d, "algorithmic_bytes": plan.algorithmic_bytes,
                      "passes": "1 sampling pass over 1/32 of the data + 1 counting pass + 1 merge pass over 3 sources, fp32 output"},
           "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4),
                        "traffic": measured_traffic("mc_ties_plan_run", 1, plan.algorithmic_bytes),
                        "traffic_unit": "DRAM read+write bytes of one plan run summed over its kernels (ncu capture, profiles/traffic.json), scaled by "
                                        "this run's algorithmic bytes",
                        "peak_source": peak_src, "kernel": "whole mc_ties_plan_run (sample, bracket, count, select, merge, fix-up)"},
           "stats": st, "gpu_launches": 6 * steps}  # sample + bracket, count + window select, full-range pass (no-op), merge, fix-up, re-merge (no-op)
    if e2e:
        # the same merge through the host-buffer entry point the CLI uses (mc_ties_host: sources in pinned host memory -> device
        # slabs -> plan -> outputs back in pinned host memory); every call allocates and frees its device memory, as the CLI's does
        import time
        def arena(tensors, dtype):
            flat = torch.empty(sum(t.numel() for t in tensors), dtype=dtype, pin_memory=True)
            views, pos = [], 0
            for t in tensors:
                views.append(flat[pos:pos + t.numel()].view(t.shape))
                pos += t.numel()
            return views
        h_srcs = []
        for lst in srcs:
            views = arena(lst, torch.bfloat16)
            for v, t in zip(views, lst):
                v.copy_(t)
            h_srcs.append(views)
        h_outs = arena(outs, outs[0].dtype)
        torch.cuda.synchronize()
        times = []
        for it in range(4):   # first call: context / allocator warm-up
            t0 = time.perf_counter()
            _, hst = M.ties_merge_host_tensors(h_srcs, K, func, outputs=h_outs)
            times.append(time.perf_counter() - t0)
        for j in (0, len(shapes) // 2, len(shapes) - 1):
            if not torch.equal(h_outs[j].view(torch.int32 if h_outs[j].dtype == torch.float32 else torch.int16),
                               outs[j].cpu().view(torch.int32 if outs[j].dtype == torch.float32 else torch.int16)):
                raise SystemExit(f"PARITY FAILURE: host-buffer TIES output tensor {j} differs from the device-resident run")
        if hst["thresholds"] != st["thresholds"]:
            raise SystemExit("PARITY FAILURE: host-buffer TIES thresholds differ from the device-resident run")
        sec = sum(times[1:]) / len(times[1:])
        res["e2e"] = {"value": round(plan.algorithmic_bytes / sec / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": 3 * d * 2,
                      "d2h_bytes_per_step": d * outs[0].element_size(), "ms_per_step": round(sec * 1e3, 2),
                      "what": "merge.ties_merge_host_tensors (mc_ties_host) on pinned host tensors: H2D of the 3 sources, plan, kernels, D2H of the "
                              "outputs, device memory allocated and freed inside the call; wall clock around the synchronous call, 3 calls"}
        del h_srcs, h_outs
    del plan, srcs, outs
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------ prefill workload
PREFILL_CONFIGS = {
    # name: (BASELINE config, requests per GPU, merged adapters (infer_modals order), modalities in every request, text tokens)
    "c3": ("composed image+audio prefill, vicuna-7B shape, batch 32, 576 image + 256 audio + 128 text tokens (C3)",
           32, ["audio", "vision", "video"], ["vision", "audio"], 128),
    "c4": ("composed video+image+audio prefill (MUSIC-AVQA shape), 8 requests per GPU (C4)",
           8, ["audio", "vision", "video"], ["video", "vision", "audio"], 128),
    "c5": ("4-modality MCUB-4 composition prefill, 16 requests per GPU (C5)",
           16, ["audio", "vision", "video", "point"], ["vision", "audio", "video", "point"], 128),
}


def bf16_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["bf16_tflops_sustained"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


def build_prefill(cfg_name: str, device, rank: int, layers=None, materialize=None, decode_dense=None):
    """Random-init composed model + one batch of requests (ids on host and device, encoder features on host and device)."""
    from modelcompose_b200 import model as MD
    from modelcompose_b200 import splice as SP
    from modelcompose_b200 import synthetic as syn
    desc, batch, merged, present, n_text_total = PREFILL_CONFIGS[cfg_name]
    coeff = 0.25 if len(merged) == 4 else 0.333
    cfg, base, adapters = syn.make_composed_on_device(merged, device, torch.bfloat16, coeff=coeff, seed=1, layers=layers)
    model = MD.MultimodalLlamaForCausalLM(MD.MultimodalConfig.from_dict(cfg), base, adapters, device=device, dtype=torch.bfloat16,
                                          materialize=materialize, decode_dense=decode_dense)
    # layer 0 as the checkpoint has it (unpacked adapters): the untimed oracle spot-check evaluates the reference schedule on it
    model._bench_layer0 = ({k: v for k, v in base.items() if k.startswith("model.layers.0.")},
                           {k: v for k, v in adapters.items() if k.startswith("model.layers.0.")}, cfg)
    del base, adapters
    n_head = 36
    n_text = n_text_total - n_head - 2 * len(present)
    # request 0 is identical on every rank (cross-rank verification probe), the others are rank-specific
    ids0 = syn.make_prompt_ids(1, present, n_text, cfg["vocab_size"], 3, SP.MODAL_TOKEN_INDEXES, n_head)
    ids = torch.cat([ids0, syn.make_prompt_ids(batch - 1, present, n_text, cfg["vocab_size"], 30 + rank, SP.MODAL_TOKEN_INDEXES, n_head)])
    feats = {}
    for m in present:
        g = torch.Generator().manual_seed(2000 + len(m))
        probe = torch.randn(1, syn.MODAL_TOKENS[m], syn.MODAL_FEATURE_DIM[m], generator=g)
        g2 = torch.Generator().manual_seed(2100 + len(m) + 10 * rank)
        # cheap deterministic fill for the rank-specific requests (randn over 100+ MB on the host would dominate set-up)
        rest = torch.randn(1, syn.MODAL_TOKENS[m], syn.MODAL_FEATURE_DIM[m], generator=g2).repeat(batch - 1, 1, 1)
        rest = rest * torch.linspace(0.5, 1.5, batch - 1).view(-1, 1, 1)
        feats[m] = torch.cat([probe, rest]).to(torch.bfloat16).pin_memory()
    ids_h = ids.pin_memory()
    mask_h = torch.ones_like(ids).pin_memory()
    flops = syn.prefill_flops(cfg, {m: syn.MODAL_TOKENS[m] for m in present}, n_text_total, len(merged), cfg["lora_r"])
    return model, desc, batch, ids_h, mask_h, feats, flops


def prefill_spot_check(model, cfg_name: str, device):
    """Untimed parity check of the bench's own full-width model, failing the run on mismatch: one short probe request with
    every modality of the config (reduced block lengths, so the CPU oracle's dense-then-mask schedule finishes in seconds)
    goes through the public forward; the hidden state after decoder layer 0 (RMSNorm, routed q/k/v + RoPE, causal attention,
    routed o / gate / up / down, residuals — every routing group of the config) is compared with oracle/model_oracle.py
    evaluated on the same spliced embeddings and modality masks.  Bar as tests/test_prefill_gpu.py: max-abs <= 2^-5 of the
    tensor's scale, cosine >= 0.9995 (bf16)."""
    from modelcompose_b200 import splice as SP
    from modelcompose_b200 import synthetic as syn
    from oracle import merge_oracle as MO
    from oracle import model_oracle as XO
    _, _, merged, present, _ = PREFILL_CONFIGS[cfg_name]
    base0, ad0, cfg = model._bench_layer0
    g = torch.Generator().manual_seed(77)
    ids = syn.make_prompt_ids(1, present, 24, cfg["vocab_size"], 78, SP.MODAL_TOKEN_INDEXES, 8)
    feats = {}
    for i, m in enumerate(present):
        n = 20 + 7 * i
        shape = (1, 2, n // 2, syn.MODAL_FEATURE_DIM[m]) if m == "video" else (1, n, syn.MODAL_FEATURE_DIM[m])
        feats[m] = torch.randn(shape, generator=g).to(torch.bfloat16).to(device)
    out = model.forward(ids.to(device), torch.ones_like(ids).to(device), modal_inputs=feats, output_hidden_states=True)
    torch.cuda.synchronize()
    x0, x1 = out.hidden_states[0].cpu(), out.hidden_states[1].float().cpu()
    names = model.modal_names
    mid = out.modal_id.cpu()
    masks = {m: (mid == i) for i, m in enumerate(names)}
    _, scaling, dnames = MO.effective_scaling(names, cfg["lora_r"], cfg["lora_alpha"], cfg["reset_scaling_weights"])
    cpu = lambda t: t.detach().to("cpu", torch.bfloat16)
    layer = {"input_layernorm": cpu(base0["model.layers.0.input_layernorm.weight"]),
             "post_attention_layernorm": cpu(base0["model.layers.0.post_attention_layernorm.weight"])}
    for ln in syn.LINEAR_NAMES:
        pfx = f"model.layers.0.{ln}."
        A = {k[len(pfx) + 7:-7]: cpu(v) for k, v in ad0.items() if k.startswith(pfx + "lora_A.")}
        Bm = {k[len(pfx) + 7:-7]: cpu(v) for k, v in ad0.items() if k.startswith(pfx + "lora_B.")}
        layer[ln.split(".")[1]] = XO.LinearParams(cpu(base0[pfx + "weight"]), A, Bm, scaling, dnames)
    S = x0.shape[1]
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = XO.decoder_layer_forward(x0, layer, masks, names, cfg["num_attention_heads"], torch.arange(S)[None],
                                       XO.causal_additive_mask(1, S, torch.bfloat16), cfg["rms_norm_eps"]).float()
    err, scale = (x1 - ref).abs().max().item(), ref.abs().max().item()
    cos = torch.nn.functional.cosine_similarity(x1.flatten(), ref.flatten(), dim=0).item()
    groups = {m: int(masks[m].sum()) for m in names}
    ok = err <= 2.0 ** -5 * scale and cos >= 0.9995 and all(groups[m] > 0 for m in ["default"] + list(present))
    res = {"what": "hidden state after decoder layer 0 of a short probe request vs oracle/model_oracle.decoder_layer_forward",
           "tokens": S, "rows_per_routing_group": groups, "max_abs": round(err, 5), "scale": round(scale, 4), "cosine": round(cos, 7),
           "bar": "max_abs <= 2^-5 * scale and cosine >= 0.9995", "ok": bool(ok), "oracle_seconds": round(time.perf_counter() - t0, 1)}
    if not ok:
        raise SystemExit(f"PARITY FAILURE (prefill {cfg_name}): {json.dumps(res)}")
    return res


def run_prefill(args, device, rank, world, dist, barrier, cfg_name: str, materialize=None, with_decode: bool = False):
    """Composed prefill, batch-sharded (every rank holds a full replica and its own requests; no collective on the timed
    path).  Returns the result dict (rank 0) with tokens/s, roofline of the routed-linear kernel, e2e and verification."""
    from modelcompose_b200 import _cabi
    from modelcompose_b200 import linear as LN
    # a branch-form model that is also asked for decode steps carries the dense text-group weights as well (model.py: DECODE_DENSE),
    # so that the decode steps can be timed in both forms after the same branch-form prefill
    both_decode_forms = with_decode and not materialize and os.environ.get("MC_BENCH_DECODE_DENSE", "1") != "0"
    model, desc, batch, ids_h, mask_h, feats_h, flops = build_prefill(cfg_name, device, rank, layers=args.prefill_layers,
                                                                      materialize=materialize, decode_dense=True if both_decode_forms else None)
    dense_requested = model.decode_dense
    if both_decode_forms:
        model.decode_dense = False   # the prefill never reads it; the first decode pass below is the branch form
    ids_d, mask_d = ids_h.to(device), mask_h.to(device)
    feats_d = {m: v.to(device) for m, v in feats_h.items()}
    S_out = int(flops["seq_len"])
    steps, warmup = max(1, args.prefill_steps), max(3, min(args.warmup, 3))

    def step_resident():
        return model.forward(ids_d, mask_d, modal_inputs=feats_d)

    for _ in range(warmup):
        out = step_resident()
    assert out.logits.shape[1] == S_out, (out.logits.shape, S_out)
    barrier()
    n0 = _cabi.LAUNCHES
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device.index) as clocks:
        t0.record()
        for _ in range(steps):
            out = step_resident()
        t1.record()
        barrier()
    launches = _cabi.LAUNCHES - n0
    t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    tokens = batch * S_out * world
    value = tokens / (ms_per_step * 1e-3)

    if os.environ.get("MC_BENCH_NVTX"):  # profiling aid: one step inside an NVTX range (ncu --nvtx --nvtx-include "mc_prefill_step/")
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("mc_prefill_step")
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
    # ---- dominant kernel: the routed/grouped linear launches, timed with CUDA events on the launching stream
    LN.start_profile()
    for _ in range(2):
        step_resident()
    lin_ms, lin_n = LN.stop_profile()
    lin_ms /= 2
    # FLOPs of the linears as EXECUTED: the materialised form has no low-rank branches (their work went into W_eff at load)
    lin_flops = (flops["linears"] - (flops["lora"] if model.materialize else 0.0)) * batch
    step_flops = (flops["total"] - (flops["lora"] if model.materialize else 0.0)) * batch
    achieved = lin_flops / (lin_ms * 1e-3) / 1e12

    # ---- e2e through the public forward API, fed the way the reference's eval loader feeds its model (collator -> batch ->
    # forward, multimodal_dataset.py:141-214): per-request instances -> data.bucket_by_length (equal spliced length: the
    # reference's batched inference raises on ragged batches) -> data.FeatureCollator -> data.PinnedBatch (pinned staging);
    # every timed step copies the staged batch host -> device, runs forward and reads the last-position logits back
    from modelcompose_b200 import data as DT
    from modelcompose_b200 import synthetic as syn
    instances = [{"input_ids": ids_h[i], "modal_inputs": {m: v[i] for m, v in feats_h.items()}} for i in range(batch)]
    rows_per_block = {m: syn.MODAL_TOKENS[m] + 10 for m in feats_h}
    groups = list(DT.bucket_by_length(instances, batch, rows_per_block))
    assert len(groups) == 1 and len(groups[0]) == batch, "synthetic requests share one layout: one bucket expected"
    collated = DT.FeatureCollator(pad_token_id=0, model_max_length=ids_h.shape[1])([instances[i] for i in groups[0]])
    staged = DT.PinnedBatch().load(collated)
    last = torch.empty((batch, model.config.vocab_size), dtype=torch.bfloat16).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in staged._batch.values()) + \
        sum(v.numel() * v.element_size() for v in staged._modal.values())
    d2h = last.numel() * 2

    def step_e2e():
        b = staged.to(device)
        o = model.forward(b["input_ids"], b["attention_mask"], modal_inputs=b["modal_inputs"])
        last.copy_(o.logits[:, -1, :], non_blocking=True)
        torch.cuda.synchronize()
    step_e2e()
    barrier()
    e_steps = max(1, min(steps, 3))
    w0 = time.perf_counter()
    for _ in range(e_steps):
        step_e2e()
    dt = time.perf_counter() - w0
    te = torch.tensor([dt], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = tokens / (float(te.item()) / e_steps)

    # ---- verification: request 0 is the same on every rank; NCCL all-gather of its last-position logits (untimed)
    probe = out.logits[0, -1, :].float().contiguous()
    finite = bool(torch.isfinite(out.logits[:, -1, :]).all())
    verified = None
    if world > 1:
        gathered = [torch.empty_like(probe) for _ in range(world)]
        dist.all_gather(gathered, probe)
        verified = all(torch.equal(g, gathered[0]) for g in gathered)
    peak_sus, peak_burst, peak_src = bf16_peaks()
    ws0 = next(ws for key, ws in model._ws.items() if key[0] == batch)
    up_tuning = ws0.up_tuning & 0xff
    up_kernel = {3: "mc::linear2_kernel<4> (512x256 CTA-pair tiles, cta_group::2), segmented by routing group" if model.materialize else
                 "mc::linear2_kernel<4> (512x256 CTA-pair tiles, cta_group::2) for the base + LoRA-up launches, mc::linear_kernel<128,6> for LoRA-down",
                 4: "mc::linear3_kernel<6> (256x256 CTA-pair tiles) for the base + LoRA-up launches, mc::linear_kernel<128,6> for LoRA-down"
                 }.get(up_tuning, "mc::linear_kernel<256,4>")
    if getattr(ws0, "up_mixed", False):
        up_kernel += "; up_proj(+SiLU*mul) and o_proj on mc::linear_kernel<256,4> (single-CTA, overlapped epilogue)"
    spot = prefill_spot_check(model, cfg_name, device) if rank == 0 else None  # replaces the batch's workspace: keep it last
    res = {
        "metric": "composed-prefill tokens/s", "value": round(value, 1), "unit": "tokens/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak", "dtype": "bf16",
        "data": "synthetic", "vs_baseline": None,
        "config": {"workload": desc, "requests_per_gpu": batch, "seq_len_after_splice": S_out, "prefix_suffix": "5+5",
                   "layers": model.config.num_hidden_layers, "sharding": f"by request batch x{world}, full replica per GPU",
                   "l2": "weights 13.5 GB + activations per step, far larger than L2",
                   "linear_form": "materialised W_eff per routing group (grouped GEMM, weights blended at load by the merge kernel)"
                   if model.materialize else "base weight + low-rank branch of the token's group (the reference's form)",
                   "algorithmic_tflop_per_step_per_gpu": round(step_flops / 1e12, 2)},
        "roofline": {"bound": "tensor", "achieved": round(achieved, 1), "peak": peak_sus, "unit": "TFLOP/s",
                     "frac": round(achieved / peak_sus, 4), "traffic": step_traffic(cfg_name, model.materialize),
                     "traffic_unit": "DRAM read+write bytes of ALL kernels of one step of this config (NVTX-scoped ncu capture, "
                                     "profiles/r02_prefill_traffic.txt); null for configs without a capture",
                     "peak_source": peak_src,
                     "kernel": up_kernel + " (routed LoRA linears, projector, lm_head)",
                     "launches_per_step": lin_n // 2, "kernel_ms_per_step": round(lin_ms, 3),
                     "kernel_share_of_step": round(lin_ms / ms_per_step, 4),
                     "algorithmic_tflop_per_step": round(lin_flops / 1e12, 2), "frac_of_burst_peak": round(achieved / peak_burst, 4),
                     "whole_step_tflops": round(step_flops / (ms_per_step * 1e-3) / 1e12, 1)},
        "e2e": {"value": round(e2e_value, 1), "unit": "tokens/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e_steps, "api": "data.bucket_by_length -> data.FeatureCollator -> data.PinnedBatch (pinned host staging) -> "
                "MultimodalLlamaForCausalLM.forward(input_ids, attention_mask, modal_inputs=...); last-position logits copied back", "timer": "host wall clock incl. synchronize, max over ranks"},
        "gpu_launches": int(launches) * world,
        "attention": "mc::attention3_kernel (tcgen05, P / O in TMEM)" if __import__("modelcompose_b200.model", fromlist=["x"]).ATTENTION_NATIVE
                     else "library call (cuDNN via torch SDPA)",
        "verification": {"finite_logits": finite, "oracle_spot_check": spot, "probe_request_identical_across_ranks": verified,
                         "collective": "ncclAllGather of request 0 last-position logits, outside the timed region" if world > 1 else None},
        "clocks": clocks.summary(),
    }
    if with_decode:
        res["decode"] = run_decode(args, model, cfg_name, ids_d, mask_d, feats_d, device, rank, world, dist, barrier)
        if both_decode_forms and dense_requested:
            model.decode_dense = True
            model._dws = None        # the decode workspace is rebuilt on the dense text-group weights
            res["decode_dense"] = run_decode(args, model, cfg_name, ids_d, mask_d, feats_d, device, rank, world, dist, barrier)
    del model
    torch.cuda.empty_cache()
    return res


def step_traffic(cfg_name: str, materialized: bool):
    """DRAM bytes of one whole prefill step from the committed NVTX-scoped ncu capture of that config, or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        d = json.load(f)
    v = d.get(f"prefill step {cfg_name} {'materialized' if materialized else 'branch'}")
    return None if v is None else int(v["dram_bytes"])


def hbm_peak():
    return measured_peaks()


def run_decode(args, model, cfg_name, ids_d, mask_d, feats_d, device, rank, world, dist, barrier, steps: int = 64):
    """Decode steps after the config's prefill (generation loop of modelcompose/eval/model_multimodal_qa_loader.py:93-102 over
    multimodal_arch.py:290-293 / multimodal_llama.py:436-438): one new token per request per step against the key/value cache,
    default adapter on every row.  A step = one replay of the captured graph (embedding gather, 32 layers of stream-K skinny
    linears + split-KV attention, lm_head, greedy argmax).  HBM-bound: the roofline is weight + cache bytes over copy bandwidth."""
    from modelcompose_b200 import _cabi
    from modelcompose_b200 import decode as DC
    batch = ids_d.shape[0]
    warmup = max(3, min(args.warmup, 5))
    out = model.forward(ids_d, mask_d, modal_inputs=feats_d, use_cache=True, last_logits_only=True, cache_extra=2 * (steps + warmup) + 16)
    cache = out.past_key_values
    S0 = cache.length
    cache.prefill_mask = None
    model._rope_tables(cache.capacity)
    dws = model._decode_workspace(cache)
    first = torch.empty(batch, dtype=torch.int64, device=device)
    DC.argmax_rows(out.logits[:, -1, :].contiguous(), None, first)
    dws.ids.copy_(first)
    dws.pos.fill_(cache.length)
    for _ in range(warmup):   # eager step, capture, replays
        dws.run()
        cache.length += 1
    assert dws.graph is not None, "the decode step must run as a captured graph"
    barrier()
    n0 = _cabi.LAUNCHES
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    len0 = cache.length
    with ClockSampler(device.index) as clocks:
        t0.record()
        for _ in range(steps):
            dws.run()
        t1.record()
        barrier()
    cache.length += steps
    launches = _cabi.LAUNCHES - n0
    t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    value = batch * world / (ms_per_step * 1e-3)
    avg_len = len0 + (steps + 1) / 2.0
    w_bytes, kv_bytes = dws.weight_bytes, dws.cache_bytes(1) * avg_len
    if os.environ.get("MC_BENCH_NVTX"):  # profiling aid: one decode step inside an NVTX range (ncu --nvtx --nvtx-include "mc_decode_step/")
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("mc_decode_step")
        dws.run()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        cache.length += 1
    # the dominant kernel alone: the skinny linears launched one by one with CUDA events around each (same buffers, no graph)
    ev = []
    dws_e = DC.DecodeWorkspace(model, cache, None, use_graph=False)
    dws_e.ids.copy_(dws.ids)
    dws_e.pos.fill_(cache.length)
    orig = DC.SkinnyLaunch.run

    def timed_run(self):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig(self)
        b.record()
        ev.append((a, b))
    DC.SkinnyLaunch.run = timed_run
    try:
        for _ in range(2):
            dws_e.run()
            dws_e.pos.fill_(cache.length)   # same position again: the cache entry is simply rewritten
    finally:
        DC.SkinnyLaunch.run = orig
    torch.cuda.synchronize()
    lin_ms = sum(a.elapsed_time(b) for a, b in ev) / 2
    del dws_e
    # ---- e2e: a streaming server's step — the new token ids come from pinned host memory, the step's greedy tokens go back
    ids_host = torch.empty((batch, 1), dtype=torch.int64).pin_memory()
    tok_host = torch.empty(batch, dtype=torch.int64).pin_memory()
    ids_host.copy_(dws.next64.cpu()[:, None])
    e_steps = min(steps, 32)
    torch.cuda.synchronize()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e_steps):
        o = model.forward(ids_host.to(device, non_blocking=True), None, past_key_values=cache, use_cache=True)
        tok_host.copy_(dws.next64, non_blocking=True)
        torch.cuda.synchronize()
        ids_host[:, 0] = tok_host
    dt = time.perf_counter() - w0
    te = torch.tensor([dt], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = batch * world / (float(te.item()) / e_steps)
    finite = bool(torch.isfinite(o.logits).all())
    # request 0 is the same on every rank and every kernel is deterministic: its decode logits must agree bit for bit (untimed)
    verified = None
    if world > 1:
        probe0 = o.logits[0, 0, :].float().contiguous()
        gathered = [torch.empty_like(probe0) for _ in range(world)]
        dist.all_gather(gathered, probe0)
        verified = all(torch.equal(g_, gathered[0]) for g_ in gathered)
    # ---- untimed parity check on the bench's own weights: decode steps of a short probe batch vs the prefill of the extended
    # sequence (the prefill path itself is checked against the oracle in this run).  The bar (that of tests/test_decode_gpu.py)
    # is applied to a 2-layer view of the model (same full-width weights): the two paths round at the same points but sum in a
    # different order, and 32 random-init layers amplify those last-bit differences, so the full-depth numbers are reported
    # next to it with a sanity bar only.
    check = None
    if rank == 0:
        import copy
        from modelcompose_b200 import splice as SP
        from modelcompose_b200 import synthetic as syn
        _, _, merged, present, _ = PREFILL_CONFIGS[cfg_name]
        g = torch.Generator().manual_seed(91)
        pids = syn.make_prompt_ids(2, present, 16, model.config.vocab_size, 92, SP.MODAL_TOKEN_INDEXES, 6).to(device)
        pf = {m: torch.randn((2, 12 + 5 * i, syn.MODAL_FEATURE_DIM[m]), generator=g).to(torch.bfloat16).to(device) for i, m in enumerate(present)}

        def probe(mdl):
            gen = mdl.generate(pids, modal_inputs=pf, max_new_tokens=4, do_sample=False)
            full = mdl.forward(gen, torch.ones_like(gen), modal_inputs=pf)
            Sp = full.logits.shape[1]
            o1 = mdl.forward(pids, torch.ones_like(pids), modal_inputs=pf, use_cache=True, cache_extra=8)
            c2 = o1.past_key_values
            worst, cosm, scale = 0.0, 1.0, 0.0
            for i in range(3):
                tok = gen[:, pids.shape[1] + i:pids.shape[1] + i + 1]
                od = mdl.forward(tok, None, past_key_values=c2, modal_inputs=pf)
                a, b = od.logits[:, 0].float().flatten(), full.logits[:, Sp - 4 + i, :].float().flatten()
                worst, scale = max(worst, (a - b).abs().max().item()), max(scale, b.abs().max().item())
                cosm = min(cosm, torch.nn.functional.cosine_similarity(a, b, dim=0).item())
            return worst, scale, cosm
        shallow = copy.copy(model)
        shallow.layers, shallow._ws, shallow._dws, shallow._proj_cache = model.layers[:2], {}, None, {}
        w2, s2, c2_ = probe(shallow)
        del shallow
        w32, s32, c32 = probe(model)
        ok = w2 <= 2.0 ** -5 * s2 and c2_ >= 0.9995 and c32 >= 0.99
        check = {"what": "logits of 3 decode steps of a 2-request probe vs the prefill of the extended sequence (same weights)",
                 "layers_2": {"max_abs": round(w2, 5), "scale": round(s2, 4), "cosine": round(c2_, 7)},
                 "all_layers": {"max_abs": round(w32, 5), "scale": round(s32, 4), "cosine": round(c32, 7)},
                 "bar": "2-layer view: max_abs <= 2^-5 * scale and cosine >= 0.9995; full depth: cosine >= 0.99 (last-bit "
                        "differences amplified by 32 random-init layers)", "ok": bool(ok)}
        if not ok:
            raise SystemExit(f"PARITY FAILURE (decode {cfg_name}): {json.dumps(check)}")
    peak, peak_src = hbm_peak()
    achieved = (w_bytes + kv_bytes) / (ms_per_step * 1e-3) / 1e9
    lin_achieved = w_bytes / (lin_ms * 1e-3) / 1e9
    return {
        "metric": "composed-decode tokens/s", "value": round(value, 1), "unit": "tokens/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "dtype": "bf16",
        "data": "synthetic", "vs_baseline": None,
        "config": {"workload": PREFILL_CONFIGS[cfg_name][0] + " — decode steps after that prefill", "requests_per_gpu": batch,
                   "context_tokens": [int(len0), int(len0 + steps)], "prompt_tokens_after_splice": int(S0),
                   "linear_form": "materialised W_eff of the default group" if model.materialize else
                   "dense W_eff of the default group for the decode steps only, branch-form prefill (decode_dense)" if model.decode_dense else
                   "base weight + low-rank branch of the default group (rank %d)" % (dws.t[0].shape[1] if dws.t else 0),
                   "l2": "weights %.1f GB + cache %.1f GB per step, far larger than L2" % (w_bytes / 1e9, kv_bytes / 1e9),
                   "step": "one CUDA-graph replay: embedding gather, 32 layers, final norm, lm_head, greedy argmax"},
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": (lambda a, b: None if a is None or b is None else int(a + b))(
                         measured_traffic("mc::skinny_streamk_kernel", world, int(w_bytes)), measured_traffic("mc::decode_attention_kernel", world, int(kv_bytes))),
                     "traffic_unit": "DRAM bytes per step: weight bytes x (traffic / algorithmic) of the ncu capture of the stream-K kernel + cache bytes x "
                                     "that of the attention kernel (profiles/traffic.json)",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_step": int(w_bytes + kv_bytes), "weight_bytes": int(w_bytes), "kv_cache_bytes": int(kv_bytes),
                     "kernel": "mc::skinny_streamk_kernel (persistent stream-K skinny linears: TMA box ring -> ldmatrix -> mma.sync.m16n8k16); attention: "
                               "mc::decode_attention_kernel (split-KV, 16 lanes per key)",
                     "kernel_ms_per_step": round(lin_ms, 4), "kernel_share_of_step": round(lin_ms / ms_per_step, 4),
                     "kernel_achieved_GBps_on_weight_bytes": round(lin_achieved, 1), "kernel_frac": round(lin_achieved / peak, 4),
                     "kernel_timing": "event pairs around every skinny-linear launch of an ungraphed step (includes launch gaps)"},
        "e2e": {"value": round(e2e_value, 1), "unit": "tokens/s", "h2d_bytes_per_step": int(batch * 8), "d2h_bytes_per_step": int(batch * 8),
                "steps": e_steps, "api": "MultimodalLlamaForCausalLM.forward(input_ids [B,1] from pinned host memory, past_key_values=cache); "
                "greedy tokens copied back and synchronised every step", "timer": "host wall clock incl. synchronize, max over ranks"},
        "gpu_launches": int(launches) * world, "launches_per_step": dws.launches_per_step(),
        "verification": {"finite_logits": finite, "decode_vs_prefill_check": check, "probe_request_identical_across_ranks": verified,
                         "collective": "ncclAllGather of request 0's decode logits after %d steps, outside the timed region" % (warmup + steps + e_steps)
                         if world > 1 else None},
        "clocks": clocks.summary(),
    }


def cpu_prefill_layer_rate(cfg_name: str, max_seconds: float = 25.0):
    """Reference schedule on host cores (oracle port of LocalLoraAttention / LocalLoraMLP: every adapter on every token,
    then mask-and-sum) for ONE full-width decoder layer, batch 1, bf16; extrapolated x layers to tokens/s."""
    from modelcompose_b200 import synthetic as syn
    from oracle import merge_oracle as MO
    from oracle import model_oracle as XO
    desc, batch, merged, present, n_text = PREFILL_CONFIGS[cfg_name]
    llama = syn.VICUNA_7B
    H, I, r = llama["hidden_size"], llama["intermediate_size"], 128
    S = n_text + sum(syn.MODAL_TOKENS[m] + 10 for m in present)
    g = torch.Generator().manual_seed(0)
    names = ["default"] + merged
    _, scaling, dnames = MO.effective_scaling(names, r, 256, ",".join(f"default-{m}=0.333" for m in merged))
    layer = {"input_layernorm": torch.ones(H, dtype=torch.bfloat16), "post_attention_layernorm": torch.ones(H, dtype=torch.bfloat16)}
    for ln in syn.LINEAR_NAMES:
        o, i = syn.linear_shape(llama, ln)
        W = (torch.randn(o, i, generator=g) * 0.02).to(torch.bfloat16)
        A = {a: (torch.randn(r, i, generator=g) * 0.01).to(torch.bfloat16) for a in merged + dnames}
        Bm = {a: (torch.randn(o, r, generator=g) * 0.02).to(torch.bfloat16) for a in merged + dnames}
        layer[ln.split(".")[1]] = XO.LinearParams(W, A, Bm, scaling, dnames)
    x = (torch.randn(1, S, H, generator=g) * 0.5).to(torch.bfloat16)
    seg = torch.zeros(S, dtype=torch.long)
    o = 40
    for m in present:
        n = syn.MODAL_TOKENS[m] + 10
        seg[o:o + n] = names.index(m)
        o += n + 3
    masks = {m: (seg == names.index(m))[None] for m in merged}
    masks["default"] = (seg == 0)[None]
    ordered = {m: masks[m] for m in names}
    pos = torch.arange(S)[None]
    add = XO.causal_additive_mask(1, S, torch.bfloat16)
    t0 = time.perf_counter()
    reps = 0
    with torch.no_grad():
        while True:
            XO.decoder_layer_forward(x, layer, ordered, names, llama["num_attention_heads"], pos, add, 1e-5)
            reps += 1
            if time.perf_counter() - t0 > max_seconds or reps >= 3:
                break
    per_layer = (time.perf_counter() - t0) / reps
    tok_s = S / (per_layer * llama["num_hidden_layers"])
    return {"value": round(tok_s, 2), "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"one full-width decoder layer (reference dense-then-mask schedule), batch 1, S'={S}, bf16, {reps} rep(s), "
                      f"{per_layer:.2f} s/layer; extrapolated x{llama['num_hidden_layers']} layers (projector, splice, lm_head excluded)",
            "host_cpus": os.cpu_count()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "merge", "prefill", "ties", "decode"],
                    help="all (default): the merge line (BASELINE config 2) carrying the prefill result under \"prefill\"; "
                         "merge / prefill: that workload alone as the primary line")
    ap.add_argument("--prefill-config", default="c3", choices=sorted(PREFILL_CONFIGS))
    ap.add_argument("--prefill-steps", type=int, default=5)
    ap.add_argument("--materialize", action="store_true", default=os.environ.get("MC_MATERIALIZE", "0") != "0",
                    help="evaluate the routed linears in the materialised form (dense W_eff per routing group) on the primary prefill lines")
    ap.add_argument("--prefill-layers", type=int, default=None, help="development aid: fewer decoder layers (never a bench value)")
    ap.add_argument("--merge-config", default="c2", choices=sorted(MERGE_CONFIGS))
    ap.add_argument("--merge-all", action="store_true", help="with --workload merge: also nest the other merge configs")
    ap.add_argument("--quick", action="store_true", help="development aid: primary merge line + one prefill config only (no nested lines)")
    ap.add_argument("--e2e-trace", action="store_true", help="development aid: per-pass times of the host-buffer merge on stderr")
    ap.add_argument("--tuning", type=int, default=0)
    ap.add_argument("--emulate-world", type=int, default=1,
                    help="profiling aid: run rank 0's shard of a K-way job on one GPU (ncu captures)")
    ap.add_argument("--no-e2e", action="store_true", help="profiling aid: skip the host-buffer e2e and CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="tuning aid: skip the CPU baseline of the prefill line (never a bench line)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE line, the JSON: anything native libraries print there (NCCL's version banner under
    # NCCL_DEBUG) goes to stderr instead — fd 1 is pointed at fd 2 while the job runs and the line is written to the saved fd
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    dist, rank, world, device, barrier = setup_dist(args)
    line = None
    if args.workload == "ties":
        if rank == 0:
            emit(run_ties(device, e2e=not args.no_e2e))
        if world > 1:
            barrier()
            dist.destroy_process_group()
        return
    nested = args.workload == "all" and not args.no_e2e and not args.quick
    if args.workload in ("all", "merge"):
        job = MergeJob(args, world, rank, device)
        line = run_merge(args, dist, rank, world, device, barrier, job, args.merge_config, with_cpu=True)
        if nested or args.workload == "merge" and args.merge_all:
            # the other merge shapes BASELINE names: 4 checkpoints (config 5) and base + N (the materialised blend)
            for name in ("n4", "c2b"):
                if name == args.merge_config:
                    continue
                sub = run_merge(args, dist, rank, world, device, barrier, job, name, with_cpu=False)
                if line is not None and sub is not None:
                    line["merge_" + name] = sub
        del job
        torch.cuda.empty_cache()
        if line is not None and world == 1 and not args.no_e2e:
            line["ties"] = run_ties(device)  # the other merge strategy family of the CLI, one GPU
    if args.workload in ("all", "prefill", "decode") and not args.no_e2e:
        configs = [(args.prefill_config, args.materialize)]
        if nested:
            configs += [(c, args.materialize) for c in ("c4", "c5") if c != args.prefill_config]
            configs += [(args.prefill_config, not args.materialize)]   # the other evaluation form of the linears, same config
        for i, (cfg_name, mat) in enumerate(configs):
            # decode steps after the prefill of the first config (both evaluation forms of it when nested)
            with_decode = args.workload == "decode" or (args.workload == "all" and cfg_name == args.prefill_config)
            pre = run_prefill(args, device, rank, world, dist, barrier, cfg_name, materialize=mat, with_decode=with_decode)
            if rank == 0:
                # CPU baseline on rank 0 at N=1 only (under torchrun the other ranks would wait on it), first config only
                pre["cpu_baseline"] = (cpu_prefill_layer_rate(cfg_name, max_seconds=15.0)
                                       if world == 1 and not args.no_cpu_baseline and i == 0 else None)
                dec = pre.pop("decode", None)
                dec_dense = pre.pop("decode_dense", None)
                suffix = ""
                if mat != args.materialize:
                    suffix = "_materialized" if mat else "_branch"
                if args.workload == "decode" and i == 0:
                    line = dec
                    line["prefill"] = pre
                    if dec_dense is not None:
                        line["decode_dense"] = dec_dense
                elif args.workload == "prefill" and i == 0:
                    line = pre
                elif line is not None:
                    key = "prefill" if i == 0 and args.workload == "all" else "prefill_" + cfg_name
                    line[key + suffix] = pre
                    if dec is not None:
                        line["decode" + ("" if i == 0 else "_" + cfg_name) + suffix] = dec
                    if dec_dense is not None:   # branch-form prefill, dense text-group weights for the decode steps
                        line["decode_dense" + ("" if i == 0 else "_" + cfg_name)] = dec_dense
    if rank == 0 and line is not None:
        emit(line)
    if world > 1:
        barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
