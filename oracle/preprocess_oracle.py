"""ORACLE (test infrastructure, NOT product code): numpy restatement of the reference's frame front end
(uint8 frame -> 448x448 tiles -> ToTensor/Normalize -> bf16).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.

DS = /root/reference/internvl/train/dataset.py. The tile-grid choice and cropping follow DS:685-738 and the
transform follows DS:268-274. The resampling arithmetic itself is NOT in the reference tree: ``image.resize``
(DS:722,736) is Pillow's ``Image.resize`` with its default filter, BICUBIC (third-party, un-vendored and unpinned in
/root/reference/requirements.txt; this image has Pillow 12.2.0). ``resize_bicubic_u8`` restates Pillow's published
two-pass fixed-point algorithm (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc, ImagingResample) and is pinned bit-exactly against the
installed Pillow in tests/test_preprocess_oracle.py (Pillow is in the image on the GPU box too).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c
IMAGENET_MEAN = (0.485, 0.456, 0.406)  # /root/reference/internvl/train/constants.py:10
IMAGENET_STD = (0.229, 0.224, 0.225)   # /root/reference/internvl/train/constants.py:11


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, in0: float, in1: float, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (support 2.0).
    Returns (ksize, bounds[out,2] int, kk[out,ksize] int32 fixed point)."""
    support_base = 2.0
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support_base * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ws = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in ws:
            ww += w
        for x in range(xmax):
            w = ws[x] / ww if ww != 0.0 else ws[x]
            if w < 0:
                kk[xx, x] = int(-0.5 + w * (1 << PRECISION_BITS))
            else:
                kk[xx, x] = int(0.5 + w * (1 << PRECISION_BITS))
        bounds[xx, 0] = xmin
        bounds[xx, 1] = xmax
    return ksize, bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """Apply the fixed-point taps along axis 0 of img [in, ...] uint8 -> [out, ...] uint8 (clip8 of Resample.c)."""
    ksize = kk.shape[1]
    acc = np.full((out_size,) + img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
    src = img.astype(np.int64)
    for t in range(ksize):
        idx = np.minimum(bounds[:, 0] + t, img.shape[0] - 1)
        w = np.where(t < bounds[:, 1], kk[:, t], 0)
        acc += src[idx] * w.reshape((-1,) + (1,) * (img.ndim - 1))
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Pillow ``Image.resize((out_w,out_h))`` (default BICUBIC, full box) on a uint8 HWC image."""
    in_h, in_w = img.shape[:2]
    if (in_w, in_h) == (out_w, out_h):
        return img.copy()  # Image.resize returns self.copy()
    need_h = out_w != in_w
    need_v = out_h != in_h
    _, bh, kh = precompute_coeffs(in_w, 0.0, float(in_w), out_w)
    _, bv, kv = precompute_coeffs(in_h, 0.0, float(in_h), out_h)
    ybox_first = int(bv[0, 0])
    ybox_last = int(bv[out_h - 1, 0] + bv[out_h - 1, 1])
    cur = img
    if need_h:
        bv = bv.copy()
        bv[:, 0] -= ybox_first
        rows = cur[ybox_first:ybox_last]
        cur = _resample_axis0(rows.transpose(1, 0, 2), out_w, bh, kh).transpose(1, 0, 2)
    if need_v:
        cur = _resample_axis0(cur, out_h, bv, kv)
    return np.ascontiguousarray(cur)


def find_closest_aspect_ratio(aspect_ratio, target_ratios, width, height, image_size):
    """DS:685-699."""
    best_ratio_diff = float("inf")
    best_ratio = (1, 1)
    area = width * height
    for ratio in target_ratios:
        target_aspect_ratio = ratio[0] / ratio[1]
        ratio_diff = abs(aspect_ratio - target_aspect_ratio)
        if ratio_diff < best_ratio_diff:
            best_ratio_diff = ratio_diff
            best_ratio = ratio
        elif ratio_diff == best_ratio_diff:
            if area > 0.5 * image_size * image_size * ratio[0] * ratio[1]:
                best_ratio = ratio
    return best_ratio


def tile_grid(width, height, min_num=1, max_num=6, image_size=448):
    """Grid choice of dynamic_preprocess (DS:703-713): returns (i, j) = tiles across, tiles down."""
    aspect_ratio = width / height
    target_ratios = set((i, j) for n in range(min_num, max_num + 1) for i in range(1, n + 1)
                        for j in range(1, n + 1) if min_num <= i * j <= max_num)
    target_ratios = sorted(target_ratios, key=lambda x: x[0] * x[1])
    return find_closest_aspect_ratio(aspect_ratio, target_ratios, width, height, image_size)


def dynamic_preprocess(img: np.ndarray, min_num=1, max_num=6, image_size=448, use_thumbnail=False):
    """DS:702-738 on a uint8 HWC array: resize to the chosen grid, crop tiles row-major, optional thumbnail."""
    h, w = img.shape[:2]
    gi, gj = tile_grid(w, h, min_num, max_num, image_size)
    tw, th = image_size * gi, image_size * gj
    resized = resize_bicubic_u8(img, tw, th)
    tiles = []
    for t in range(gi * gj):
        x0 = (t % gi) * image_size
        y0 = (t // gi) * image_size
        tiles.append(resized[y0:y0 + image_size, x0:x0 + image_size].copy())
    if use_thumbnail and len(tiles) != 1:
        tiles.append(resize_bicubic_u8(img, image_size, image_size))
    return tiles


def transform(tile_u8: np.ndarray) -> np.ndarray:
    """build_transform eval branch (DS:268-274) on a 448x448 uint8 HWC tile: Resize is a no-op at 448,
    ToTensor (/255 in fp32, CHW), Normalize ((x-mean)/std in fp32). Returns fp32 [3,448,448]."""
    x = tile_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    mean = np.asarray(IMAGENET_MEAN, dtype=np.float32).reshape(3, 1, 1)
    std = np.asarray(IMAGENET_STD, dtype=np.float32).reshape(3, 1, 1)
    return ((x - mean) / std).astype(np.float32)


def load_video_frames(frames_u8: np.ndarray, max_num=1, use_thumbnail=True, image_size=448):
    """Call shape of load_video (/root/reference/internvl/train/internvl/eval/stage1_eval.py:452-456,484-485):
    per frame dynamic_preprocess(...); video path keeps tile [0] only when max_num == 1. Returns
    (tiles_u8 [T,448,448,3], pixel_values fp32 [T,3,448,448], tiles_per_frame)."""
    tiles, counts = [], []
    for f in frames_u8:
        t = dynamic_preprocess(f, max_num=max_num, use_thumbnail=use_thumbnail, image_size=image_size)
        if max_num == 1:
            t = t[:1]
        tiles.extend(t)
        counts.append(len(t))
    tiles_u8 = np.stack(tiles)
    pv = np.stack([transform(t) for t in tiles])
    return tiles_u8, pv, counts
