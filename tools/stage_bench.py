#!/usr/bin/env python
"""Per-stage timings and a per-launch table of the tensor-core kernel (writes gpurun_out/*.csv).
Development aid: the judged numbers come from bench.py."""
import ctypes as C
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
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    lib = tb.lib()
    wdir = weights.export_random(ROOT / "tests" / "_cache" / "weights_bench_seed0")
    eng = tb.Engine(wdir, devices=[0])
    stream = torch.cuda.ExternalStream(lib.tt_engine_stream(eng._h, 0))
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    tag = sys.argv[2] if len(sys.argv) > 2 else "full"
    import os
    print(f"== {tag}: TT_GEMM_PAIR={os.environ.get('TT_GEMM_PAIR')} TT_ENC_CHUNK={os.environ.get('TT_ENC_CHUNK')}")
    pages = [torch.from_numpy(synth.synth_page(i)).cuda() for i in range(B)]
    maps = [torch.from_numpy(synth.synth_score_maps(i)).cuda() for i in range(B)]

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def dump(fn, path):
        Path(path).unlink(missing_ok=True)
        fn()
        torch.cuda.synchronize()
        lib.tt_profile_enable(1)
        fn()
        torch.cuda.synchronize()
        lib.tt_profile_enable(0)
        ms, fl, n = C.c_double(), C.c_double(), C.c_ulonglong()
        lib.tt_profile_dump(str(path).encode(), C.byref(ms), C.byref(fl), C.byref(n))
        return ms.value, fl.value, n.value

    full = lambda: eng.ocr_pages(pages, score_override=maps)  # noqa: E731
    t_full = timed(full)
    print(f"full pipeline, {B} pages: {t_full:.2f} ms  ({t_full / B:.3f} ms/page, {B / t_full * 1e3:.1f} pages/s)")
    ms, fl, n = dump(full, out / f"gemm_launches_{tag}.csv")
    print(f"  gemm kernel: {n} launches, {ms:.2f} ms, {fl / ms / 1e9:.1f} TFLOP/s")

    # CRAFT only / PARSeq only via the stage entry points (host buffers: includes copies, so only the GEMM
    # table is meaningful here)
    if tag != "full":
        eng.close()
        return
    crops = np.random.default_rng(0).integers(0, 256, (1024, 32, 128, 3), dtype=np.uint8)
    pq = lambda: eng.parseq_forward(crops)  # noqa: E731
    t0 = time.perf_counter(); pq(); torch.cuda.synchronize(); t1 = time.perf_counter()
    t0 = time.perf_counter(); pq(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"parseq_forward 1024 crops (host in/out): {(t1 - t0) * 1e3:.2f} ms -> {1024 / (t1 - t0):.0f} crops/s")
    ms, fl, n = dump(pq, out / "gemm_launches_parseq.csv")
    print(f"  gemm kernel: {n} launches, {ms:.2f} ms, {fl / ms / 1e9:.1f} TFLOP/s")
    ci, _ = tb.preprocess(synth.synth_page(0))
    cf = lambda: eng.craft_forward(ci)  # noqa: E731
    ms, fl, n = dump(cf, out / "gemm_launches_craft1.csv")
    print(f"craft 1 page gemm kernel: {n} launches, {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s")
    eng.close()


if __name__ == "__main__":
    main()
