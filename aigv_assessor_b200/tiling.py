"""Host logic of the frame front end: which tile grid a frame is resized to.

The grid choice of ``dynamic_preprocess`` (/root/reference/internvl/train/dataset.py:685-713: candidate grids i x j with
min_num <= i*j <= max_num ordered by area, the one whose aspect ratio is closest to the frame's, ties resolved towards
the larger grid when the frame has more than half its pixels) is implemented once, in the library
(``aigv_tile_grid``, csrc/api.cu); this module is the Python face of it. tests/test_host_cpu.py holds it to the
oracle restatement and, in the build container, to the reference function over ~20 000 frame sizes incl. exact ties.
"""
from __future__ import annotations

import ctypes as C

from . import _lib


def tile_grid(width: int, height: int, min_num: int = 1, max_num: int = 6, image_size: int = 448):
    """(tiles across, tiles down) for a width x height frame."""
    if width <= 0 or height <= 0:
        raise ValueError("empty frame")
    gi, gj = C.c_int(), C.c_int()
    st = _lib.load().aigv_tile_grid(int(width), int(height), int(min_num), int(max_num), int(image_size),
                                    C.byref(gi), C.byref(gj))
    _lib.check(st, None, "aigv_tile_grid")
    return gi.value, gj.value


def tiles_per_frame(width: int, height: int, max_num: int = 6, use_thumbnail: bool = True, image_size: int = 448):
    gi, gj = tile_grid(width, height, 1, max_num, image_size)
    blocks = gi * gj
    return blocks + (1 if (use_thumbnail and blocks != 1) else 0)
