"""Checkpoint ingest for the path: read ``vision_model.*`` / ``mlp1.*`` tensors straight from
safetensors shards -- the format ``InternVLChatModel.save_pretrained`` / ``from_pretrained`` uses in the reference's
eval scripts (stage2_eval.py:764-780) -- without constructing the PyTorch model, and hand them to ``aigv_set_weight``.

The safetensors container is: u64 little-endian header length, a JSON header {name: {dtype, shape, data_offsets}} and the
raw little-endian tensor bytes. Only the tensors of the path are touched (the LLM's 7 B parameters are skipped by offset),
and LoRA pairs of a peft-wrapped tower (``...base_layer.weight`` + ``lora_A/lora_B``, eval2 :275-283) are merged like
``tools/merge_lora.py:18-21`` does (``VisionEngine.load_state_dict`` -> ``merge_lora_names``).
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, Iterable

import numpy as np

_DTYPES = {"F32": (np.float32, 4), "F16": (np.float16, 2), "BF16": (np.uint16, 2), "F64": (np.float64, 8)}


def _wanted(name: str) -> bool:
    n = name.replace("base_model.model.", "")
    return n.startswith("vision_model.") or n.startswith("mlp1.")


def read_safetensors(path: str, wanted=_wanted) -> Dict[str, np.ndarray]:
    """{name: fp32 ndarray} of the tensors for which wanted(name) holds. bf16 / fp16 are widened to fp32 exactly."""
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen).decode("utf-8"))
        base = 8 + hlen
        for name, info in header.items():
            if name == "__metadata__" or not wanted(name):
                continue
            if info["dtype"] not in _DTYPES:
                raise ValueError(f"{path}: tensor {name} has unsupported dtype {info['dtype']}")
            dt, esz = _DTYPES[info["dtype"]]
            lo, hi = info["data_offsets"]
            count = int(np.prod(info["shape"], dtype=np.int64)) if info["shape"] else 1
            if hi - lo != count * esz:
                raise ValueError(f"{path}: tensor {name} has {hi - lo} bytes, expected {count * esz}")
            f.seek(base + lo)
            raw = np.frombuffer(f.read(hi - lo), dtype=dt)
            if info["dtype"] == "BF16":
                arr = (raw.astype(np.uint32) << 16).view(np.float32)
            else:
                arr = raw.astype(np.float32)
            out[name] = arr.reshape(info["shape"])
    return out


def write_safetensors(path: str, tensors: Dict[str, np.ndarray], bf16_names: Iterable[str] = ()) -> None:
    """Minimal writer (tests / tools): fp32 tensors, or bf16 for the names listed (values must be bf16-representable)."""
    bf16_names = set(bf16_names)
    header, blobs, off = {}, [], 0
    for name, arr in tensors.items():
        a = np.ascontiguousarray(arr, dtype=np.float32)
        if name in bf16_names:
            blob = (a.view(np.uint32) >> 16).astype(np.uint16).tobytes()
            dt = "BF16"
        else:
            blob = a.tobytes()
            dt = "F32"
        header[name] = {"dtype": dt, "shape": list(a.shape), "data_offsets": [off, off + len(blob)]}
        blobs.append(blob)
        off += len(blob)
    hj = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hj += b" " * (-len(hj) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


def load_checkpoint_dir(path: str) -> Dict[str, np.ndarray]:
    """All path tensors of a checkpoint directory: ``model.safetensors`` or the shards of ``model.safetensors.index.json``."""
    if os.path.isfile(path):
        return read_safetensors(path)
    idx = os.path.join(path, "model.safetensors.index.json")
    if os.path.exists(idx):
        with open(idx) as f:
            wm = json.load(f)["weight_map"]
        files = sorted({v for k, v in wm.items() if _wanted(k)})
    else:
        files = sorted(fn for fn in os.listdir(path) if fn.endswith(".safetensors"))
    if not files:
        raise FileNotFoundError(f"no safetensors files with vision_model.* / mlp1.* tensors under {path}")
    sd: Dict[str, np.ndarray] = {}
    for fn in files:
        sd.update(read_safetensors(os.path.join(path, fn)))
    return sd
