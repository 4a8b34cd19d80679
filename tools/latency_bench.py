#!/usr/bin/env python
"""Single-page latency through the public API (configs[0]/[3] shapes: resume 763x607, FUNSD 1000x754, synthetic
1280x1280; ~300 words via the score-map override), host buffers in, items out.  Development aid / DESIGN numbers."""
import sys
import time
from pathlib import Path

import cv2
import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tuatara_b200 as tb  # noqa: E402
from tuatara_b200 import synth, weights  # noqa: E402

wdir = weights.export_random(ROOT / "tests" / "_cache" / "weights_bench_seed0")
eng = tb.Engine(wdir, devices=[0])
full, fmap = synth.synth_page(0), synth.synth_score_maps(0)
for name, (h, w) in {"resume 763x607": (763, 607), "funsd 1000x754": (1000, 754), "synthetic 1280x1280": (1280, 1280)}.items():
    page = np.ascontiguousarray(cv2.resize(full, (w, h), interpolation=cv2.INTER_AREA))
    _, _, h32, w32, _ = tb.resize_plan(h, w)
    m = np.ascontiguousarray(cv2.resize(fmap, (w32 // 2, h32 // 2), interpolation=cv2.INTER_LINEAR))
    for _ in range(3):
        out = eng.ocr_pages([page], score_override=[m])
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        out = eng.ocr_pages([page], score_override=[m])
        ts.append(time.perf_counter() - t0)
    print(f"{name}: {len(out[0])} items, median {np.median(ts) * 1e3:.2f} ms, min {min(ts) * 1e3:.2f} ms per page (1 page per call)")
eng.close()

# stage breakdown of the single-page call (CUDA-event intervals on the engine's stream)
import ctypes as C  # noqa: E402
lib = tb.lib()
eng = tb.Engine(wdir, devices=[0])
page, m = full, fmap
for _ in range(3):
    eng.ocr_pages([page], score_override=[m])
lib.tt_profile_enable(1)
N = 5
t0 = time.perf_counter()
for _ in range(N):
    eng.ocr_pages([page], score_override=[m])
wall = (time.perf_counter() - t0) / N
lib.tt_profile_enable(0)
buf = C.create_string_buffer(1 << 14)
lib.tt_profile_stages(buf, len(buf))
pm, pf, pl = C.c_double(), C.c_double(), C.c_ulonglong()
lib.tt_profile_collect(C.byref(pm), C.byref(pf), C.byref(pl))
print(f"single 1280x1280 page with profiling on: {wall * 1e3:.2f} ms wall; gemm launches {pl.value // N} taking {pm.value / N:.2f} ms")
for ln in buf.value.decode().splitlines():
    name, cnt, ms, fl, by = ln.split(",")
    print(f"  {name:16s} {float(ms) / N:7.3f} ms")
eng.close()
