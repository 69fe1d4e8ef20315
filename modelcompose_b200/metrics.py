"""Parameter-interference metrics of a merged checkpoint — drop-in for the reference's
``scripts/model_composition/calculate_metrics.py`` (same function names, same ``merge_metrics.txt``).

The reference loads every input checkpoint named in ``merge_info.txt``, casts to float32, flattens the keys all inputs
share and computes L2 / cosine distance between the first two inputs plus the soft sign dissimilarity before and after a
top-k trim with torch ops on the CPU.  Here the tensors go through ``mc_interference_host`` (one exact top-k select + one
fused reduction pass on the GPU, include/modelcompose_b200.h); there is no CPU arithmetic path.
"""
from __future__ import annotations

import argparse
import os
import re
from collections import defaultdict
from pathlib import Path

import torch

from .merge import convert_delta_to_ft, interference_metrics_host


def parse_merge_info(file):
    """reference calculate_metrics.py:14-23."""
    pattern = r"Inputs:\n(.*?)\n\nOutput\((.*?)\):(.*?)$"
    match = re.search(pattern, open(file).read().strip(), re.DOTALL)
    if match:
        return match.group(1).split("\n"), match.group(2), match.group(3)
    return None, None, None


def calculate_metrics(merged_ckpt, reset_thresh=50):
    """reference calculate_metrics.py:41-74 — writes ``merge_metrics.txt`` next to the merged checkpoint, prints the four
    metrics, and (extra) returns them.  The values print in the reference's format: L2 / SSD / TSSD are 0-dim float32
    tensors there (``tensor(1.2345)``), Cosine a Python float."""
    filepaths, _, _ = parse_merge_info(Path(merged_ckpt) / "merge_info.txt")
    weights_to_merge = defaultdict(list)
    for filepath in filepaths:
        adapter_path = os.path.join(filepath, "adapter_model.bin")
        if not os.path.exists(adapter_path):
            adapter_path = os.path.join(filepath, "mm_projector.bin")
        adapter_weights = torch.load(adapter_path, map_location=torch.device("cpu"))
        for key in adapter_weights:
            weights_to_merge[key].append(adapter_weights[key])  # the reference's .float() is exact: the kernels widen per element
    ft_checks, _ = convert_delta_to_ft(weights_to_merge)
    keys = sorted(ft_checks[0])
    m = interference_metrics_host([[check[k] for k in keys] for check in ft_checks], reset_thresh)
    l2, ssd, tssd = (torch.tensor(m[k], dtype=torch.float32) for k in ("L2", "SSD", "TSSD"))
    cosine_sim = float(torch.tensor(m["Cosine"], dtype=torch.float32))
    with open(Path(merged_ckpt) / "merge_metrics.txt", "w") as fout:
        fout.write(f"L2: {l2}\n")
        fout.write(f"Cosine: {cosine_sim}\n")
        fout.write(f"SSD: {ssd}\n")
        fout.write(f"TSSD: {tssd}\n")
    print(f"L2: {l2}\n")
    print(f"Cosine: {cosine_sim}\n")
    print(f"SSD: {ssd}\n")
    print(f"TSSD: {tssd}\n")
    return m


def main(argv=None):
    parser = argparse.ArgumentParser(description="Calculate parameter interference metrics")
    parser.add_argument("merged_ckpt", help="Path to the merged checkpoint")
    args = parser.parse_args(argv)
    calculate_metrics(args.merged_ckpt)


if __name__ == "__main__":
    main()
