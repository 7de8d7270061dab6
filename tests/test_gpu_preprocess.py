"""CUDA frame front end vs Pillow/torchvision semantics: bit-exact uint8 tiles and bit-exact bf16 pixel_values
against the reference goldens, the oracle and the installed Pillow."""
import hashlib

import numpy as np
import pytest
import torch

from aigv_assessor_b200 import synth
from aigv_assessor_b200.config import tiny
from aigv_assessor_b200.engine import VisionEngine
from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return VisionEngine(tiny(), max_tiles=1)


def test_front_end_matches_reference_goldens(eng, golden_meta):
    for case in golden_meta["preprocess"]["cases"]:
        frame = synth.make_frames(1, case["h"], case["w"], seed=case["seed"], kind=case["kind"])
        pv, counts, u8 = eng.preprocess_frames(torch.from_numpy(frame), max_num=case["max_num"], use_thumbnail=True,
                                               return_u8=True)
        torch.cuda.synchronize()
        assert counts == [case["n_tiles"]]
        assert hashlib.sha256(u8.cpu().numpy().tobytes()).hexdigest() == case["tiles_sha256"], case
        bits = pv.view(torch.int16).cpu().numpy().view(np.uint16)
        assert hashlib.sha256(bits.tobytes()).hexdigest() == case["pv_bf16_sha256"], case


@pytest.mark.parametrize("h,w,max_num,kind", [(720, 1280, 1, "uniform"), (300, 200, 6, "smooth"), (449, 447, 1, "uniform"),
                                              (96, 1000, 6, "uniform"), (2160, 3840, 1, "smooth")])
def test_front_end_matches_oracle_and_pillow(eng, h, w, max_num, kind):
    from PIL import Image
    frames = synth.make_frames(2, h, w, seed=h + w, kind=kind)
    pv, counts, u8 = eng.preprocess_frames(torch.from_numpy(frames), max_num=max_num, return_u8=True)
    torch.cuda.synchronize()
    tiles, ref_pv, ref_counts = po.load_video_frames(frames, max_num=max_num, use_thumbnail=True)
    assert counts == ref_counts
    assert np.array_equal(u8.cpu().numpy(), tiles)
    assert np.array_equal(pv.view(torch.int16).cpu().numpy().view(np.uint16), synth.to_bf16_bits(ref_pv))
    if max_num == 1:
        pil = np.asarray(Image.fromarray(frames[1]).resize((448, 448)))
        assert np.array_equal(u8[1].cpu().numpy(), pil)


def test_front_end_rejects_bad_input(eng):
    with pytest.raises(ValueError):
        eng.preprocess_frames(torch.zeros((1, 10, 10, 4), dtype=torch.uint8))
    with pytest.raises(ValueError):
        eng.preprocess_frames(torch.zeros((1, 10, 10, 3), dtype=torch.float32))
    pv, counts = eng.preprocess_frames(torch.zeros((0, 64, 64, 3), dtype=torch.uint8))
    assert pv.shape == (0, 3, 448, 448) and counts == []
