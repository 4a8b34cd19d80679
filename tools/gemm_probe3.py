#!/usr/bin/env python
"""Times GEMM shapes / tile choices through the C ABI with CUDA events (development aid).
Each case: K N act f32 BN flags  (flags with BN given: bit0 weight-resident, bit1 CTA pair)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tuatara_b200 as tb  # noqa: E402

lib = tb.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 307200
cases = [(384, 1152, 0, 0, 192, 3), (384, 1024, 0, 0, 256, 3), (384, 128, 0, 0, 128, 3), (384, 128, 0, 0, 128, 1),
         (384, 1152, 0, 0, 128, 3), (384, 1536, 2, 0, 256, 3),
         (1536, 384, 0, 1, 192, 2), (1536, 256, 0, 1, 256, 2), (1536, 128, 0, 1, 128, 2), (1536, 384, 0, 1, 128, 2),
         (384, 384, 0, 1, 192, 3), (384, 256, 0, 1, 256, 3), (384, 128, 0, 1, 128, 3), (384, 384, 0, 1, 128, 3)]
for K, N, act, f32, BN, res in cases:
    A = (torch.randn(M, K) * 0.5).to(torch.bfloat16).cuda()
    W = (torch.randn(N, K) * 0.1).to(torch.bfloat16).cuda()
    b = torch.randn(N).float().cuda()
    R = torch.randn(M, N).float().cuda() if f32 else None
    out = torch.empty(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device="cuda")
    def run():
        tb.check(lib.tt_linear_dev(A.data_ptr(), K, M, K, W.data_ptr(), N, b.data_ptr(), act, R.data_ptr() if f32 else None,
                                   1, N, out.data_ptr(), f32, N, BN, res, None), "lin")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"M{M} K{K} N{N} act{act} f32{f32} BN{BN} flags{res}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TF", flush=True)
