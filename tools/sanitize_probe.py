#!/usr/bin/env python
"""Small end-to-end + PARSeq runs for compute-sanitizer (memcheck / racecheck): exercises the round-2 kernels
(k_dec_dense, LayerNorm-fused GEMMs, fused max-pool, lookup-table self attention, crop warp) at ragged sizes."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tuatara_b200 as tb  # noqa: E402
from tuatara_b200 import _native, synth, weights  # noqa: E402


def main():
    wdir = weights.export_random(ROOT / "tests" / "_cache" / "weights_bench_seed0")
    cfg = _native.tt_config()
    tb.lib().tt_config_default(cfg)
    cfg.rectify = 1
    eng = tb.Engine(wdir, devices=[0], cfg=cfg)
    rng = np.random.default_rng(0)
    crops = rng.integers(0, 256, (77, 32, 128, 3), dtype=np.uint8)
    l, ids = eng.parseq_forward(crops)
    assert np.isfinite(l).all()
    page = np.ascontiguousarray(synth.synth_page(0)[:352, :416])
    maps = np.ascontiguousarray(synth.synth_score_maps(0)[:176, :208])
    out = eng.ocr_pages([page, page], score_override=[maps, maps])
    print("ok", len(out[0]), ids.shape)
    eng.close()


if __name__ == "__main__":
    main()
