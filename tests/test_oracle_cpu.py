"""The oracle against its pins: the committed golden vectors (cv2 4.13 / torch 2.11 outputs at the
reference's call sites), cv2 itself for the numpy restatement of cv::resize, and internal
consistency of the restated pipeline.  (The reference ships no vectors of its own: SURVEY.md 4.)"""
from pathlib import Path

import cv2
import numpy as np
import torch

from oracle import tuatara_ref as R
from oracle.cvmath import resize_linear_u8
from tuatara_b200 import synth

GOLD = Path(__file__).resolve().parent / "golden"


def _rects(det):
    return np.array([[r[0][0], r[0][1], r[1][0], r[1][1], r[2]] for r in det], np.float32).reshape(-1, 5)


def test_postprocess_golden():
    g = np.load(GOLD / "postprocess.npz")
    for name in ("synth", "blobs"):
        maps = g[name + "_maps"]
        det, dbg = R.get_detected_boxes(torch.from_numpy(maps[..., 0].copy()), torch.from_numpy(maps[..., 1].copy()), 0.7, 0.4, 0.4)
        assert np.array_equal(dbg.labels, g[name + "_labels"])
        assert np.array_equal(dbg.stats[:, :5], g[name + "_stats"])
        assert np.array_equal(_rects(det).view(np.uint32), g[name + "_rects"].view(np.uint32))
        assert list(dbg.mapper) == list(g[name + "_mapper"])


def test_resize_golden_and_numpy_restatement():
    g = np.load(GOLD / "resize.npz")
    for key, pre in (("img", "pre_small"), ("big", "pre_big")):
        img = g[key]
        got, ratio, _ = R.resize_aspect_ratio(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), 1024, cv2.INTER_LINEAR)
        assert np.array_equal(got, g[pre])
        th, tw, h32, w32, _ = R.resize_target(img.shape[0], img.shape[1], 1024)
        mine = np.zeros_like(got)
        mine[:th, :tw] = resize_linear_u8(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), tw, th)
        assert np.array_equal(mine, g[pre])
    assert max(g["pre_big"].shape) == 1024 and R.resize_target(1128, 300, 1024)[0] == 1023
    sw = cv2.cvtColor(g["img"], cv2.COLOR_BGR2RGB)
    for r, c in zip(g["rects"], g["crops"]):
        x, y, w, h = (int(v) for v in r)
        assert np.array_equal(R.crop_to_parseq_u8(sw, (x, y, w, h)), c)
        assert np.array_equal(cv2.cvtColor(resize_linear_u8(sw[y:y + h, x:x + w], 128, 32), cv2.COLOR_BGR2RGB), c)


def test_numpy_resize_matches_cv2_fuzz():
    rng = np.random.default_rng(0)
    for t in range(120):
        h, w = int(rng.integers(1, 200)), int(rng.integers(1, 300))
        dw, dh = (128, 32) if t % 3 == 0 else (int(rng.integers(1, 300)), int(rng.integers(1, 200)))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(resize_linear_u8(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


def test_oracle_nets_golden():
    from oracle.models import make_craft, make_parseq

    g = np.load(GOLD / "nets.npz")
    craft, parseq = make_craft(0), make_parseq("base", 0)
    x = torch.from_numpy(g["craft_in"])[None].permute(0, 3, 1, 2).float().div(255.0)
    with torch.no_grad():
        maps = craft(x)[0][0].numpy()
    assert np.allclose(maps, g["craft_maps"], atol=2e-4, rtol=1e-4)
    logits = parseq(torch.from_numpy(g["crops_u8"]).permute(0, 3, 1, 2).float().div(255.0)).numpy()
    assert np.allclose(logits, g["logits"], atol=2e-3, rtol=1e-3)
    assert sum(p.numel() for p in craft.parameters()) == 20770466
    assert sum(p.numel() for p in parseq.parameters()) == 23832671


def test_craft_skip_tensors_are_relu_aliased():
    """Upstream's in-place ReLU at the head of slice2..4 aliases the skip tensors (oracle/models.py)."""
    from oracle.models import make_craft

    taps = {}
    with torch.no_grad():
        make_craft(0)(torch.rand(1, 3, 64, 64), taps)
    assert float(taps["relu2_2"].min()) >= 0 and float(taps["relu3_2"].min()) >= 0 and float(taps["relu4_3"].min()) >= 0
    assert float(taps["relu5_3"].min()) < 0  # slice5 starts with a MaxPool: stays pre-ReLU


def test_synthetic_page_gives_300_boxes_inside_the_image():
    maps = synth.synth_score_maps(5)
    det, dbg = R.get_detected_boxes(torch.from_numpy(maps[..., 0].copy()), torch.from_numpy(maps[..., 1].copy()), 0.7, 0.4, 0.4)
    assert dbg.n_labels == 301 and len(det) == 300
    inv = np.float32(1) / np.float32(0.8)
    boxes = R.adjust_result_coordinates(det, inv, inv)
    assert not any(R.crop_rect((1280, 1280, 3), b)[4] for b in boxes)


def test_text_threshold_is_compared_as_float_like_the_reference():
    """tuatara.cpp:154 `if (maxVal < text_threshold)`: double against the float parameter (0.7f -> 0.699999988...).  A
    component whose maximum is exactly float32(0.7) is kept; the oracle used to compare against the Python double 0.7
    and dropped it (round-1 verdict)."""
    import numpy as np
    import torch
    t = np.float32(0.7)
    m = np.zeros((48, 96, 2), np.float32)
    m[4:12, 4:24, 0] = 0.5;  m[6, 10, 0] = t
    m[20:28, 4:24, 0] = 0.5; m[22, 10, 0] = np.nextafter(t, np.float32(0))
    m[4:12, 60:80, 0] = 1.0
    m[0, 0, 1], m[0, 1, 1] = 0.0, 1.0
    det, dbg = R.get_detected_boxes(torch.from_numpy(m[..., 0].copy()), torch.from_numpy(m[..., 1].copy()), 0.7, 0.4, 0.4)
    kept_rows = sorted(int(round(r[0][1])) for r in det)
    assert len(det) == 2 and kept_rows[0] < 12, kept_rows   # the float32(0.7) blob (rows 4..11) and the 1.0 blob survive
    assert float(t) < 0.7   # the whole point: float32(0.7) is below the double 0.7


def test_parseq_upstream_early_exit_keeps_the_outputs():
    """Upstream PARSeq leaves the AR loop once every sequence of the batch has an EOS; the oracle applies it when
    `early_exit` is set (bench.py's CPU baseline, 4 crops per forward like tuatara.cpp:452-475).  Every key the shorter
    refinement drops sits behind an EOS and would be masked, so the (N, 26, C) logits and the strings do not change."""
    from oracle.models import make_parseq

    m = make_parseq("base", 0).eval()
    x = torch.rand(8, 3, 32, 128, generator=torch.Generator().manual_seed(3))
    tok = R.Tokenizer()
    for chunk in (x[:4], x[4:]):
        m.early_exit = False
        full = m(chunk)
        m.early_exit = True
        short = m(chunk)
        m.early_exit = False
        assert short.shape == full.shape
        assert torch.allclose(short, full, atol=1e-5, rtol=0)
        assert tok.decode(short) == tok.decode(full)
