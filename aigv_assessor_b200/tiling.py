"""Host logic of the frame front end: which tile grid a frame is resized to.

Mirrors ``find_closest_aspect_ratio`` and the grid choice of ``dynamic_preprocess``
(/root/reference/internvl/train/dataset.py:685-713) expression for expression — including CPython's iteration order of
the candidate set, which decides ties between equal-area grids — so the CUDA front end lands on the same tiles.
"""
from __future__ import annotations

from functools import lru_cache


def find_closest_aspect_ratio(aspect_ratio, target_ratios, width, height, image_size):
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


@lru_cache(maxsize=None)
def _target_ratios(min_num: int, max_num: int):
    ratios = set((i, j) for n in range(min_num, max_num + 1) for i in range(1, n + 1) for j in range(1, n + 1)
                 if i * j <= max_num and i * j >= min_num)
    return tuple(sorted(ratios, key=lambda x: x[0] * x[1]))


def tile_grid(width: int, height: int, min_num: int = 1, max_num: int = 6, image_size: int = 448):
    """(tiles across, tiles down) for a width x height frame."""
    if width <= 0 or height <= 0:
        raise ValueError("empty frame")
    return find_closest_aspect_ratio(width / height, _target_ratios(min_num, max_num), width, height, image_size)


def tiles_per_frame(width: int, height: int, max_num: int = 6, use_thumbnail: bool = True, image_size: int = 448):
    gi, gj = tile_grid(width, height, 1, max_num, image_size)
    blocks = gi * gj
    return blocks + (1 if (use_thumbnail and blocks != 1) else 0)
