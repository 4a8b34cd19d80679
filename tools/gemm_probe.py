#!/usr/bin/env python
"""Launches one GEMM shape through the C ABI (for ncu captures): K N act f32 BN resident reps."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tuatara_b200 as tb  # noqa: E402

lib = tb.lib()
M = 131072
K, N, act, f32, BN, res, reps = (int(v) for v in sys.argv[1:8])
A = (torch.randn(M, K) * 0.5).to(torch.bfloat16).cuda()
W = (torch.randn(N, K) * 0.1).to(torch.bfloat16).cuda()
b = torch.randn(N).float().cuda()
R = torch.randn(M, N).float().cuda() if f32 else None
out = torch.empty(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device="cuda")
for _ in range(reps):
    tb.check(lib.tt_linear_dev(A.data_ptr(), K, M, K, W.data_ptr(), N, b.data_ptr(), act, R.data_ptr() if f32 else None,
                               1, N, out.data_ptr(), f32, N, BN, res, None), "lin")
torch.cuda.synchronize()
print("done")
