#!/usr/bin/env python
"""Launches / times one 3x3 conv shape through the C ABI: B H W C0 Cout reps (development aid)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tuatara_b200 as tb  # noqa: E402

lib = tb.lib()
B, H, W, C0, Cout, reps = (int(v) for v in sys.argv[1:7])
x = torch.randn(B, H, W, C0).to(torch.bfloat16).cuda()
w = (torch.randn(Cout, 3, 3, C0) * 0.05).to(torch.bfloat16).cuda()
b = torch.randn(Cout).float().cuda()
out = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device="cuda")
def run():
    tb.check(lib.tt_conv_dev(x.data_ptr(), C0, None, 0, B, H, W, 9, 1, w.data_ptr(), b.data_ptr(), Cout, 1, out.data_ptr(), 0, 0, None), "conv")
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"conv {B}x{H}x{W} C{C0}->{Cout}: {ms*1e3:.1f} us  {2*B*H*W*Cout*9*C0/ms/1e9:.0f} TF")
