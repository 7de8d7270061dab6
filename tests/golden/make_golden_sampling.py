"""Golden vectors for the frame-sampling rule of the eval scripts: run the UNMODIFIED reference's get_index
(/root/reference/internvl/train/internvl/eval/stage1_eval.py:429-441) on a grid of (bound, fps, frame count, segments).

    python tests/golden/make_golden_sampling.py      (build container only: needs /root/reference)

get_index is a method that does not touch `self`; the module is imported with the same stubs as make_golden.py (decord,
imageio, peft, pytorchvideo, timm are absent in this image). Output: tests/golden/golden_sampling.json.
"""
import inspect
import json
import os
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = "/root/reference/internvl/train/internvl/eval/stage1_eval.py"


def reference_get_index():
    """Extract the reference's own source of get_index and compile it (importing stage1_eval.py drags in deepspeed and a
    dozen absent packages; the function itself is self-contained numpy)."""
    src = open(REF_FILE).read().splitlines()
    start = next(i for i, l in enumerate(src) if l.strip().startswith("def get_index("))
    end = next(i for i in range(start + 1, len(src)) if src[i].strip().startswith("def "))
    body = textwrap.dedent("\n".join(src[start:end]))
    ns = {"np": np}
    exec(compile(body, REF_FILE, "exec"), ns)
    return ns["get_index"]


def main():
    get_index = reference_get_index()
    cases = []
    for bound in (None, (0.0, 4.0), (1.5, 9.25), (2.0, 2.5)):
        for fps in (8.0, 23.976, 24.0, 25.0, 29.97, 30.0, 60.0):
            for n_frames in (9, 16, 33, 81, 120, 121, 241, 1000):
                for seg in (8, 16, 32, 64):
                    idx = get_index(None, bound, fps, n_frames - 1, first_idx=0, num_segments=seg)
                    cases.append({"bound": bound, "fps": fps, "max_frame": n_frames - 1, "num_segments": seg,
                                  "indices": [int(v) for v in idx]})
    with open(os.path.join(HERE, "golden_sampling.json"), "w") as f:
        json.dump({"source": "stage1_eval.py:429-441 get_index, executed from the reference file", "cases": cases}, f)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
