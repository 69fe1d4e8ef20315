#!/usr/bin/env python
"""Drop-in for the reference CLI of the same path:

    python scripts/model_composition/merge_unimodal_modelcompose.py CKPT_DIR... -o OUT_DIR \
        --strategy online-merge-reset-default-video=0.333,default-audio=0.333,default-vision=0.333

Same flags, same output files (adapter_model.bin, config.json, merge_info.txt); tensor arithmetic
(`sum` / `mean`) runs in the CUDA merge kernel behind the C ABI.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from modelcompose_b200.merge import main  # noqa: E402

if __name__ == "__main__":
    main()
