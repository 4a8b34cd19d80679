#!/usr/bin/env python
"""PARSeq encoder / decoder stage times for a few batch sizes, fused vs unfused decoder (development aid)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tuatara_b200 as tb  # noqa: E402
from tuatara_b200 import weights  # noqa: E402


def stages(lib):
    buf = C.create_string_buffer(1 << 16)
    lib.tt_profile_stages(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms, fl, by = line.split(",")
        out[name] = float(ms) / max(1, int(cnt))
    return out


def main():
    lib = tb.lib()
    wdir = weights.export_random(ROOT / "tests" / "_cache" / "weights_bench_seed0")
    eng = tb.Engine(wdir, devices=[0])
    sizes = [int(a) for a in sys.argv[1:]] or [300, 2400, 9600]
    rng = np.random.default_rng(0)
    for n in sizes:
        crops = rng.integers(0, 256, (n, 32, 128, 3), dtype=np.uint8)
        for mode in ("1", "0"):
            os.environ["TT_DEC_FUSED"] = mode
            eng.parseq_forward(crops)
            torch.cuda.synchronize()
            lib.tt_profile_enable(1)
            for _ in range(3):
                eng.parseq_forward(crops)
            torch.cuda.synchronize()
            lib.tt_profile_enable(0)
            ms, fl, nl = C.c_double(), C.c_double(), C.c_ulonglong()
            lib.tt_profile_collect(C.byref(ms), C.byref(fl), C.byref(nl))
            st = stages(lib)
            print(f"n={n:5d} fused={mode}: encoder {st.get('parseq_encoder', 0):8.3f} ms  decoder {st.get('parseq_decoder', 0):8.3f} ms "
                  f"({st.get('parseq_decoder', 0) / 26 * 1e3:7.1f} us/AR step incl. refine)")
    os.environ.pop("TT_DEC_FUSED", None)
    eng.close()


if __name__ == "__main__":
    main()
