"""End to end through the public API (image_to_data / Engine.ocr_pages) vs the oracle pipeline."""
import numpy as np
import pytest
import torch

import tuatara_b200 as tb
from oracle import tuatara_ref as R
from tuatara_b200 import synth

pytestmark = pytest.mark.gpu


def test_synth_page_boxes_and_order(engine, oracle_models):
    """Score-map override: boxes, order and bboxes must be identical to the oracle's; texts equal
    wherever both decoders see the same crop (strings are checked strictly in test_models_gpu)."""
    craft, parseq = oracle_models
    pages = [synth.synth_page(i) for i in range(2)]
    maps = [synth.synth_score_maps(i) for i in range(2)]
    got = engine.ocr_pages(pages, score_override=maps)
    for i in range(2):
        st = R.Stages()
        dummy_parseq = lambda x: torch.zeros(x.shape[0], 26, 95)  # noqa: E731  (bboxes do not depend on it)
        ref = R.image_to_data(pages[i].copy(), craft, dummy_parseq,
                              score_override=(maps[i][..., 0], maps[i][..., 1]), stages=st)
        assert len(got[i]) == len(ref) == 300
        assert [g["bbox"] for g in got[i]] == [r["bbox"] for r in ref]


def test_random_weights_page_runs(engine):
    """Honest random-weights run (no override): near-constant maps -> very few boxes, must not crash."""
    out = engine.ocr_pages([synth.synth_page(3)])
    assert len(out) == 1
    for item in out[0]:
        assert len(item["bbox"]) == 4 and isinstance(item["text"], str)


def test_image_to_data_api(weights_dir, native_lib):
    img = synth.synth_page(4)[:640, :800].copy()
    res = tb.image_to_data(img, weights_dir, "outputs")
    assert isinstance(res, list)
    assert tb.image_to_data(img, "", "outputs") == []
    with pytest.raises(RuntimeError):
        tb.image_to_data(img[..., 0], weights_dir, "outputs")


def test_mixed_sizes_and_empty(engine):
    a = synth.synth_page(5)
    b = np.full((300, 500, 3), 255, np.uint8)
    out = engine.ocr_pages([a, b, a[:700]])
    assert len(out) == 3


def test_config_struct_native_canvas_and_thresholds(weights_dir, oracle_models):
    """tt_config (the reference's TODO at tuatara.cpp:396): a native-resolution canvas (1280 -> CRAFT input 1280^2,
    maps 640^2) and non-default thresholds / min_area give the oracle's boxes when it is run with the same values."""
    import cv2

    craft, _ = oracle_models
    cfg = tb.default_config()
    cfg.canvas_size, cfg.text_threshold, cfg.link_threshold, cfg.low_text, cfg.min_area = 1280.0, 0.6, 0.3, 0.35, 40
    eng = tb.Engine(weights_dir, cfg=cfg)
    try:
        page = synth.synth_page(7)
        m512 = synth.synth_score_maps(7)
        m640 = np.ascontiguousarray(cv2.resize(m512, (640, 640), interpolation=cv2.INTER_LINEAR))  # maps of the 1280 canvas
        got = eng.ocr_pages([page], score_override=[m640])[0]
        dummy_parseq = lambda x: torch.zeros(x.shape[0], 26, 95)  # noqa: E731
        ref = R.image_to_data(page.copy(), craft, dummy_parseq, score_override=(m640[..., 0], m640[..., 1]), canvas_size=1280,
                              text_threshold=0.6, link_threshold=0.3, low_text=0.35, min_area=40)
        ref_default = R.image_to_data(page.copy(), craft, dummy_parseq, score_override=(m512[..., 0], m512[..., 1]))
        assert len(ref) > 0 and [g["bbox"] for g in got] == [r["bbox"] for r in ref]
        assert [r["bbox"] for r in ref] != [r["bbox"] for r in ref_default]  # the settings do change the result
    finally:
        eng.close()


def test_two_gpu_engine_shards_pages(weights_dir):
    """One engine over two devices: page i runs on device i mod 2, results gathered in page order."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    pages = [synth.synth_page(i) for i in range(5)]
    maps = [synth.synth_score_maps(i) for i in range(5)]
    one = tb.Engine(weights_dir, devices=[0])
    two = tb.Engine(weights_dir, devices=[0, 1])
    try:
        assert two.ocr_pages(pages, score_override=maps) == one.ocr_pages(pages, score_override=maps)
    finally:
        one.close()
        two.close()


def test_pages_are_independent_units_at_full_size(engine):
    """configs[4] shape (1280x1280 pages, 300 words each, 2400 crops in one PARSeq batch): each page's items must be
    the same whether it is processed alone or inside a group -- the property the data-parallel sharding rests on."""
    pages = [synth.synth_page(i) for i in range(8)]
    maps = [synth.synth_score_maps(i) for i in range(8)]
    together = engine.ocr_pages(pages, score_override=maps)
    assert [len(p) for p in together] == [300] * 8
    for i in (0, 3, 7):
        alone = engine.ocr_pages([pages[i]], score_override=[maps[i]])[0]
        assert alone == together[i]


def test_concurrent_callers_are_serialised_safely(engine):
    """image_to_data is re-entrant in the reference (no globals); here concurrent callers of one engine share its
    execution slots behind per-slot mutexes and must each get the single-caller result."""
    from concurrent.futures import ThreadPoolExecutor

    pages = [np.ascontiguousarray(synth.synth_page(i)[:640, :768]) for i in range(4)]
    maps = [np.ascontiguousarray(synth.synth_score_maps(i)[:320, :384]) for i in range(4)]
    serial = [engine.ocr_pages([p], score_override=[m])[0] for p, m in zip(pages, maps)]
    with ThreadPoolExecutor(4) as ex:
        for _ in range(3):
            got = list(ex.map(lambda pm: engine.ocr_pages([pm[0]], score_override=[pm[1]])[0], zip(pages, maps)))
            assert got == serial


def test_two_slots_small_pages_regression(native_lib):
    """Round-1 hang (profiles/r2_hang_root_cause.md): two execution slots running SMALL pages concurrently wedged one CTA
    pair in `tcgen05.alloc.cta_group::2` -- the allocation was issued before the peer CTA had started.  The probe runs
    the regime that stalled within 0-60 iterations (TMA epilogue forced on small launches, 16-warp GELU epilogue, both
    slots busy) for 300 iterations per thread; its own watchdog exits with status 3 on a stall."""
    import os
    import subprocess
    import sys as _sys

    from conftest import ROOT
    env = dict(os.environ, TT_GEMM_TE="2", TT_GEMM_EW="16", PROBE_STALL_S="8", PROBE_TAG="pytest")
    r = subprocess.run([_sys.executable, str(ROOT / "tools" / "concurrency_probe.py"), "host", "640", "300"], env=env,
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "concurrent ok [300, 300]" in r.stdout


def test_mixed_page_sizes_share_one_recognition_batch(engine):
    """Pages of different sizes in one request: detection runs per size bucket, the crops of all of them are recognised in
    one PARSeq batch; every page must equal its single-page result (and the oracle-checked uniform case above)."""
    shapes = [(640, 768), (1280, 1280), (512, 512), (640, 768), (1000, 754), (512, 512), (1280, 1280), (763, 607)]
    pages, maps = [], []
    for i, (h, w) in enumerate(shapes):
        pages.append(np.ascontiguousarray(synth.synth_page(i)[:h, :w]))
        _, _, h32, w32, _ = tb.resize_plan(h, w)
        full = synth.synth_score_maps(i)
        if max(h, w) > 1024:
            m = full
        else:
            m = np.zeros((h32 // 2, w32 // 2, 2), np.float32)
            hh, ww = min(h32 // 2, full.shape[0]), min(w32 // 2, full.shape[1])
            m[:hh, :ww] = full[:hh, :ww]
        maps.append(np.ascontiguousarray(m))
    together = engine.ocr_pages(pages, score_override=maps)
    assert sum(len(p) for p in together) > 500
    for i in range(len(shapes)):
        alone = engine.ocr_pages([pages[i]], score_override=[maps[i]])[0]
        assert alone == together[i], i


def test_pytuatara_module_against_the_oracle(engine, weights_dir, oracle_models, native_lib):
    """The reference's Python module (bindings/python.cpp:54-58) with real weight files on the GPU.  The module has no
    score-map override, and random-init CRAFT maps are near constant (min-max normalisation then amplifies bf16 noise),
    so the oracle is fed the raw maps the CUDA CRAFT produced for the same page: from there on (normalise, CCL, boxes,
    crops, PARSeq, decode, formatting) pytuatara's result must equal the oracle's."""
    import sys

    from tuatara_b200 import _native
    sys.path.insert(0, str(_native.LIB_PATH.parent))
    import pytuatara

    craft, parseq = oracle_models
    n_boxes = 0
    for seed, (h, w) in ((0, (607, 763)), (1, (512, 640)), (2, (754, 1000))):
        page = np.ascontiguousarray(synth.synth_page(seed)[:h, :w])
        got = pytuatara.image_to_data(image=page.copy(), weights_dir=weights_dir, outputs_dir="outputs")
        assert isinstance(got, list) and all(set(g) == {"text", "bbox"} and len(g["bbox"]) == 4 for g in got)
        craft_in, _ = tb.preprocess(page)
        maps = engine.craft_forward(craft_in)
        ref = R.image_to_data(page.copy(), craft, parseq, score_override=(maps[..., 0], maps[..., 1]))
        assert [g["bbox"] for g in got] == [r["bbox"] for r in ref], (seed, len(got), len(ref))
        same = sum(g["text"] == r["text"] for g, r in zip(got, ref))
        assert same >= 0.8 * len(ref), f"page {seed}: {same}/{len(ref)} strings equal"
        assert got == tb.image_to_data(page, weights_dir, "outputs")   # the ctypes twin takes the same path
        n_boxes += len(got)
    assert n_boxes > 0


def test_rectify_option_on_the_rotated_fixture(weights_dir, oracle_models):
    """tt_config.rectify = 1 (opt-in; default 0 keeps the reference's axis-aligned crop): boxes / bboxes do not change,
    the recogniser is fed the perspective-warped quads.  Oracle: the reference algorithm up to the final boxes, then
    cv2.getPerspectiveTransform + warpPerspective per box and the fp32 PARSeq."""
    from pathlib import Path

    from oracle import imagemaps
    from tuatara_b200 import _native

    fx = np.load(Path(__file__).parent / "golden" / "fixture_images.npz")
    craft, parseq = oracle_models
    cfg = _native.tt_config()
    tb.lib().tt_config_default(cfg)
    cfg.rectify = 1
    eng_r = tb.Engine(weights_dir, devices=[0], cfg=cfg)
    eng_0 = tb.Engine(weights_dir, devices=[0])
    try:
        for name in ("rotated_text", "table_english"):
            img = fx[f"{name}.img"]
            maps = imagemaps.maps_f32(fx[f"{name}.maps_u8"])
            got_r = eng_r.ocr_pages([img], score_override=[maps])[0]
            got_0 = eng_0.ocr_pages([img], score_override=[maps])[0]
            assert [g["bbox"] for g in got_r] == [g["bbox"] for g in got_0] and len(got_r) > 0
            st = R.Stages()
            R.image_to_data(img.copy(), craft, lambda x: torch.zeros(x.shape[0], 26, 95),
                            score_override=(maps[..., 0], maps[..., 1]), stages=st)
            crops = np.stack([R.rectified_crop(img, R.rect_to_quad(b)) for b in st.boxes])
            texts = [R.truncate_at_eos(t) for t in R.Tokenizer().decode(torch.softmax(R.run_parseq(parseq, crops), -1))]
            same = sum(g["text"] == t for g, t in zip(got_r, texts))
            print(name, "rectified strings equal", same, "of", len(texts))
            assert same >= 0.8 * len(texts)
    finally:
        eng_r.close()
        eng_0.close()
