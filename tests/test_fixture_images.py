"""Parity on the reference's own fixture images (images/*.png: configs[0] resume_example, configs[3] the FUNSD page,
the table, the rotated sample).  tests/golden/fixture_images.npz holds the decoded pages, 8-bit ink-derived score maps
(random-init CRAFT gives near-constant maps: SURVEY 8d) and what the oracle returned when the fixture was made here
(tests/golden/make_golden_images.py).  CPU: the oracle still reproduces the fixture.  GPU: the CUDA path through the
C ABI returns the same boxes, in the same order, and bit-identical PARSeq input crops."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import imagemaps, tuatara_ref as R

NAMES = ["resume_example", "funsd_0001129658", "funsd_91372360", "table_english", "rotated_text"]


@pytest.fixture(scope="module")
def fx():
    return np.load(Path(__file__).parent / "golden" / "fixture_images.npz")


def _oracle(img, maps_u8):
    m = imagemaps.maps_f32(maps_u8)
    craft = lambda x: (torch.zeros(1, x.shape[2] // 2, x.shape[3] // 2, 2), None)  # noqa: E731
    parseq = lambda x: torch.zeros(x.shape[0], 26, 95)  # noqa: E731
    st = R.Stages()
    items = R.image_to_data(img.copy(), craft, parseq, score_override=(m[..., 0], m[..., 1]), stages=st)
    return items, st


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_fixture(fx, name):
    items, st = _oracle(fx[f"{name}.img"], fx[f"{name}.maps_u8"])
    assert np.array_equal(np.array([it["bbox"] for it in items], np.float32).reshape(-1, 4), fx[f"{name}.bbox"])
    assert np.array_equal(np.array(st.crop_rects, np.int32).reshape(-1, 4), fx[f"{name}.crop_rects"])
    assert np.array_equal(st.crops_u8[:8], fx[f"{name}.crops"])
    # the maps can be regenerated from the page itself (8-bit quantisation makes them machine independent)
    assert np.array_equal(imagemaps.ink_maps_u8(st.craft_input_u8), fx[f"{name}.maps_u8"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_matches_fixture(engine, fx, name):
    import tuatara_b200 as tb

    img = np.ascontiguousarray(fx[f"{name}.img"])
    got = engine.ocr_pages([img], score_override=[imagemaps.maps_f32(fx[f"{name}.maps_u8"])])[0]
    assert np.array_equal(np.array([g["bbox"] for g in got], np.float32).reshape(-1, 4), fx[f"{name}.bbox"])
    rects = fx[f"{name}.crop_rects"][:8]
    crops = tb.crop_resize(img, rects)
    assert np.array_equal(crops, fx[f"{name}.crops"])
