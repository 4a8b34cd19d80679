#!/usr/bin/env python
"""Runs the fused LayerNorm GEMM pair (producer: x += A W1^T + b, bf16(x), row sums; consumer: act(Linear(LN(x))))
a few times for ncu captures.  Usage: ln_probe.py K1 N2 act [M] [reps]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tuatara_b200 as tb  # noqa: E402
from tuatara_b200._native import check  # noqa: E402


def main():
    K1, N2, act = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    M = int(sys.argv[4]) if len(sys.argv) > 4 else 307200
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    D = 384
    lib = tb.lib()
    g = torch.Generator(device="cpu").manual_seed(0)
    A = (torch.randn(M, K1, generator=g) * 0.5).to(torch.bfloat16).cuda()
    W1 = (torch.randn(D, K1, generator=g) * 0.05).to(torch.bfloat16).cuda()
    b1 = torch.zeros(D).cuda()
    X = torch.randn(M, D, generator=g).cuda()
    XH = X.to(torch.bfloat16)
    XL = (X - XH.float()).to(torch.bfloat16)
    stats = torch.zeros(M, 8, device="cuda")
    W2f = (torch.randn(N2, D, generator=g) * 0.05).to(torch.bfloat16).cuda()
    c0, c1 = torch.zeros(N2).cuda(), W2f.float().sum(1)
    out = torch.empty(M, N2, dtype=torch.bfloat16, device="cuda")
    for _ in range(reps):
        check(lib.tt_linear_ln_pair_dev(A.data_ptr(), M, K1, W1.data_ptr(), b1.data_ptr(), D, XH.data_ptr(), XL.data_ptr(),
                                        stats.data_ptr(), W2f.data_ptr(), c0.data_ptr(), c1.data_ptr(), N2, act, 1e-6,
                                        out.data_ptr(), None), "tt_linear_ln_pair_dev")
    torch.cuda.synchronize()
    print("ok", float(out.float().abs().mean()))


if __name__ == "__main__":
    main()
