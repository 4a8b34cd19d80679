#!/usr/bin/env python
"""Fused encoder MLP (TT_ENC_MLPFUSE=1) against the two-GEMM path: logits rel-L2 under one forced AR context, ids."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tuatara_b200 as tb  # noqa: E402
from tuatara_b200 import weights  # noqa: E402


def main():
    wdir = weights.export_random(ROOT / "tests" / "_cache" / "weights_bench_seed0")
    eng = tb.Engine(wdir, devices=[0])
    rng = np.random.default_rng(0)
    for n in [int(a) for a in sys.argv[1:]] or [2, 3, 300]:
        crops = rng.integers(0, 256, (n, 32, 128, 3), dtype=np.uint8)
        os.environ["TT_ENC_MLPFUSE"] = "0"
        l0, i0 = eng.parseq_forward(crops)
        forced = np.ascontiguousarray(i0[:, :25]).astype(np.int32)
        l0, i0 = eng.parseq_forward(crops, forced)
        os.environ["TT_ENC_MLPFUSE"] = "1"
        os.environ["TT_ENC_PROJFUSE"] = os.environ.get("PROBE_PROJ", "0")
        l1, i1 = eng.parseq_forward(crops, forced)
        os.environ["TT_ENC_PROJFUSE"] = "0"
        os.environ["TT_ENC_MLPFUSE"] = "0"
        err = float(np.linalg.norm(l1.astype(np.float64) - l0) / np.linalg.norm(l0.astype(np.float64)))
        print(f"n={n}: finite {bool(np.isfinite(l1).all())} rel-L2 {err:.3e} ids equal {float((i0 == i1).mean()):.4f}", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
