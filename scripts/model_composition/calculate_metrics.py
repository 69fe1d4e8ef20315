#!/usr/bin/env python
"""Drop-in for the reference CLI of the same path:

    python scripts/model_composition/calculate_metrics.py MERGED_CKPT_DIR

Reads MERGED_CKPT_DIR/merge_info.txt, loads the input checkpoints it names and writes merge_metrics.txt (L2, Cosine,
SSD, TSSD); the arithmetic runs in the CUDA library behind the C ABI (modelcompose_b200/metrics.py).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from modelcompose_b200.metrics import main  # noqa: E402

if __name__ == "__main__":
    main()
