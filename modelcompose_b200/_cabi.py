"""ctypes binding of ``include/modelcompose_b200.h`` — the only way Python reaches the CUDA kernels.

The signatures carry plain pointers and sizes (no torch types): tensors are passed as
``tensor.data_ptr()`` and the stream as ``torch.cuda.current_stream().cuda_stream``.
There is no CPU fallback: a missing library or a failing call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_LIB: Optional[C.CDLL] = None

MC_OK = 0
MC_F32, MC_F16, MC_BF16 = 0, 1, 2
MC_MERGE_WEIGHTED, MC_MERGE_REF_SUM, MC_MERGE_REF_MEAN = 0, 1, 2
MC_MERGE_MAX_SRC = 8
MC_TIES_SUM, MC_TIES_MEAN, MC_TIES_MAX = 0, 1, 2

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libmodelcompose_b200.so")

_vp, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); must list every MC_API symbol of include/modelcompose_b200.h
SIGNATURES = {
    "mc_abi_version": (_i, []),
    "mc_last_error": (C.c_char_p, []),
    "mc_device_info": (_i, [C.c_char_p, _sz, C.POINTER(_i), C.POINTER(_i)]),
    "mc_merge_plan_create": (_i, [C.POINTER(_vp), _i, _i, _pp, _pp, C.POINTER(_i64), _i, _i, _i]),
    "mc_merge_plan_run": (_i, [_vp, C.POINTER(C.c_float), _i, _vp]),
    "mc_merge_plan_bytes": (_i64, [_vp]),
    "mc_merge_plan_destroy": (_i, [_vp]),
    "mc_merge_tensors": (_i, [_i, _i, _pp, _pp, C.POINTER(_i64), C.POINTER(C.c_float), _i, _i, _i, _vp]),
    "mc_merge_host": (_i, [_i, _i, _pp, _pp, C.POINTER(_i64), C.POINTER(C.c_float), _i, _i, _i, _sz]),
    "mc_ties_plan_create": (_i, [C.POINTER(_vp), _i, _i, _pp, _pp, C.POINTER(_i64), _i, _i]),
    "mc_ties_plan_run": (_i, [_vp, _i64, _i, _vp]),
    "mc_ties_plan_stats": (_i, [_vp, _vp, _vp]),
    "mc_ties_plan_bytes": (_i64, [_vp]),
    "mc_ties_plan_elements": (_i64, [_vp]),
    "mc_ties_plan_destroy": (_i, [_vp]),
    "mc_ties_plan_metrics": (_i, [_vp, _i64, _vp, _vp]),
    "mc_interference_host": (_i, [_i, _i, _pp, C.POINTER(_i64), _i64, _i, _vp]),
    "mc_ties_host": (_i, [_i, _i, _pp, _pp, C.POINTER(_i64), _i64, _i, _i, _vp]),
    # structs are passed as void* (modelcompose_b200.splice defines the ctypes.Structure mirrors)
    "mc_splice_plan_create": (_i, [C.POINTER(_vp), _i, _i, _i, _vp, _i]),
    "mc_splice_plan_scan": (_i, [_vp, _vp, _vp]),
    "mc_splice_plan_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), _vp, _vp]),
    "mc_splice_plan_bytes": (_i64, [_vp, _i]),
    "mc_splice_run": (_i, [_vp, _vp, _vp, _vp]),
    "mc_splice_plan_destroy": (_i, [_vp]),
    "mc_linear_plan_create": (_i, [C.POINTER(_vp), _vp, _i, _i, _i]),
    "mc_linear_plan_run": (_i, [_vp, _vp]),
    "mc_linear_plan_flops": (C.c_double, [_vp]),
    "mc_linear_plan_destroy": (_i, [_vp]),
    "mc_route_tile_masks": (_i, [_vp, _i, _vp, _vp]),
    "mc_route_tile_masks_coarse": (_i, [_vp, _i, _vp, _i, _vp]),
    "mc_route_permutation": (_i, [_vp, _i, C.c_char_p, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mc_silu_mul": (_i, [_vp, _vp, _vp, _i64, _i, _i64, _i64, _i64, _i, _vp]),
    "mc_attention_causal": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _i, _i, _i, _i, C.c_float, _i, _vp]),
    "mc_attention_causal_tuned": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _i, _i, _i, _i, C.c_float, _i, _i, _vp]),
    "mc_gather_rows": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i, _vp]),
    "mc_rmsnorm": (_i, [_vp, _vp, _vp, _i64, _i, _i64, _i64, C.c_float, _i, _vp]),
    "mc_rope": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i64, _i64, _i, _vp]),
    "mc_skinny_workspace_bytes": (_sz, []),
    "mc_skinny_plan_create": (_i, [C.POINTER(_vp), _vp, _i, _i, _i]),
    "mc_skinny_plan_run": (_i, [_vp, _vp, _sz, _vp]),
    "mc_skinny_plan_bytes": (_i64, [_vp]),
    "mc_skinny_plan_set_norm": (_i, [_vp, _vp, _i64, _vp, _vp, _i64, _i, _i, C.c_float]),
    "mc_skinny_plan_destroy": (_i, [_vp]),
    "mc_decode_rope_append": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mc_decode_attention": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _i, C.c_float, _i, _vp, _vp, _i, _vp]),
    "mc_decode_attention_fused": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, C.c_float, _i, _vp, _vp,
                                       _i, _vp]),
    "mc_argmax_rows": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "mc_set_launch_mode": (_i, [_i]),
}


class TiesStats(C.Structure):
    """mirror of mc_ties_stats_t"""
    _fields_ = [("threshold", C.c_float * MC_MERGE_MAX_SRC), ("n_pos", C.c_int64), ("n_neg", C.c_int64), ("n_zero", C.c_int64),
                ("n_ambiguous", C.c_int64), ("majority", C.c_int32), ("full_select_ran", C.c_int32), ("fix_pass_ran", C.c_int32)]


class InterferenceMetrics(C.Structure):
    """mirror of mc_interference_metrics_t"""
    _fields_ = [("l2", C.c_double), ("cosine", C.c_double), ("ssd", C.c_double), ("tssd", C.c_double),
                ("ssd_elements", C.c_int64), ("tssd_elements", C.c_int64), ("threshold", C.c_float * MC_MERGE_MAX_SRC)]


class McError(RuntimeError):
    pass


LAUNCHES = 0  # kernels of this library launched through the Python wrappers (bench.py's `gpu_launches`)


def count_launch(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def lib() -> C.CDLL:
    """Load the CUDA library (built in-tree by ``modelcompose_b200.build``).  Fails loudly."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise McError(
                f"{LIB_PATH} is missing: build it with `python -m modelcompose_b200.build` "
                "(modelcompose_b200 has no CPU or PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.mc_abi_version() != 3:
            raise McError("libmodelcompose_b200.so ABI version mismatch; rebuild")
        _LIB = handle
    return _LIB


def check(rc: int, what: str) -> None:
    if rc != MC_OK:
        msg = lib().mc_last_error()
        raise McError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def dtype_code(torch_dtype) -> int:
    import torch
    table = {torch.float32: MC_F32, torch.float16: MC_F16, torch.bfloat16: MC_BF16}
    if torch_dtype not in table:
        raise McError(f"unsupported dtype {torch_dtype} (float32 / float16 / bfloat16 only)")
    return table[torch_dtype]


def ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def i64_array(vals):
    arr = (C.c_int64 * len(vals))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr


def f32_array(vals):
    arr = (C.c_float * len(vals))()
    for i, v in enumerate(vals):
        arr[i] = float(v)
    return arr


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def bind_host_thread_to_gpu(device_index: int) -> bool:
    """Pin the calling thread to the CPUs next to GPU ``device_index`` (NVML's ideal affinity), so that the pinned host
    buffers it allocates afterwards live on that GPU's NUMA node.  One process per GPU: without this, eight ranks staging
    checkpoints from whatever node they woke up on share the socket interconnect (the host-buffer merge at N = 4 / 8 is
    bound by exactly that).  Returns False (and changes nothing) when NVML is not available."""
    try:
        import os
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = getattr(torch.cuda.get_device_properties(device_index), "uuid", None)
        if uuid is not None:
            name = str(uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID((name if name.startswith("GPU-") else "GPU-" + name).encode())
        else:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[device_index]) if vis and vis.split(",")[device_index].isdigit() else device_index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False
