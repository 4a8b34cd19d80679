"""numpy restatements of the OpenCV arithmetic the reference relies on (TEST INFRASTRUCTURE ONLY).

These exist to *document and pin* the exact integer rules the CUDA kernels implement; each is
fuzzed against cv2 itself in tests/test_oracle_cpu.py.  None of it lives in /root/reference:
it is the behaviour of the third-party calls at tuatara.cpp:223 and :440 (cv::resize).
"""
from __future__ import annotations

import numpy as np


def _axis(dst_len: int, src_len: int, horizontal: bool, exact_inverse: bool = True):
    """cv::resize INTER_LINEAR coordinate/coefficient tables for one axis (OpenCV resize.cpp)."""
    d = np.arange(dst_len, dtype=np.float64)
    scale = 1.0 / (np.float64(dst_len) / np.float64(src_len)) if exact_inverse else np.float64(src_len) / dst_len
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if horizontal:
        lo = s < 0
        s[lo], f[lo] = 0, 0.0
        hi = s >= src_len - 1
        s[hi], f[hi] = src_len - 1, 0.0
    c1 = np.rint(f * np.float32(2048)).astype(np.int32)
    c0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int32)
    return s, c0, c1


def resize_linear_u8(src: np.ndarray, dst_w: int, dst_h: int, exact_inverse: bool = True) -> np.ndarray:
    """cv2.resize(src, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for uint8 HxWxC, bit-exact."""
    H, W = src.shape[:2]
    sx, a0, a1 = _axis(dst_w, W, True, exact_inverse)
    sy, b0, b1 = _axis(dst_h, H, False, exact_inverse)
    x1 = np.minimum(sx + 1, W - 1)
    y0 = np.clip(sy, 0, H - 1)
    y1 = np.clip(sy + 1, 0, H - 1)
    I = src.astype(np.int32)
    hz = I[:, sx] * a0[None, :, None] + I[:, x1] * a1[None, :, None]  # (H, dst_w, C) int32
    S0, S1 = hz[y0], hz[y1]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)
