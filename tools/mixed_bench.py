#!/usr/bin/env python
"""Mixed-size request against a uniform one: 64 pages cycling through the reference's five fixture pages and a synthetic
1280 x 1280 page, and 64 synthetic 1280 x 1280 pages, both through tt_ocr_pages_ex from host buffers (development aid;
the engine cuts a request into detection units per size bucket and recognises the crops of all sizes in shared PARSeq
batches)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tuatara_b200 as tb  # noqa: E402
from tuatara_b200 import synth, weights  # noqa: E402


def main():
    wdir = weights.export_random(ROOT / "tests" / "_cache" / "weights_bench_seed0")
    eng = tb.Engine(wdir, devices=[0])
    fx = np.load(ROOT / "tests" / "golden" / "fixture_images.npz")
    names = ["resume_example", "funsd_0001129658", "funsd_91372360", "table_english", "rotated_text"]
    fixtures = [(np.ascontiguousarray(fx[f"{n}.img"]), np.ascontiguousarray(fx[f"{n}.maps_u8"].astype(np.float32) / np.float32(255.0)))
                for n in names]
    synth_pages = [(synth.synth_page(i), synth.synth_score_maps(i)) for i in range(64)]
    mixed = [fixtures[i % 6] if i % 6 < 5 else synth_pages[i] for i in range(64)]
    uniform = synth_pages

    def run(batch, reps=3):
        pages, maps = [b[0] for b in batch], [b[1] for b in batch]
        out = eng.ocr_pages(pages, score_override=maps)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = eng.ocr_pages(pages, score_override=maps)
        dt = (time.perf_counter() - t0) / reps
        words = sum(len(p) for p in out)
        px = sum(p.shape[0] * p.shape[1] for p in pages)
        return len(batch) / dt, words / dt, px / dt / 1e6, words

    for name, batch in (("uniform 64 x 1280x1280", uniform), ("mixed 64 (5 fixture pages + 1280x1280)", mixed)):
        pps, wps, mpx, words = run(batch)
        print(f"{name}: {pps:7.1f} pages/s  {wps:9.0f} words/s  {mpx:7.1f} Mpixel/s  ({words} words per request)", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
