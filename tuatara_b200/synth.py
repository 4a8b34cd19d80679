"""Deterministic synthetic inputs for the benchmark and the parity tests (numpy only).

Definitions follow SURVEY.md section 8d: a 1280x1280 white page with 300 "words" on a
20-line x 15-column grid, each word 2-6 dark glyph rectangles, ``default_rng(page_index)``;
and the matching 512x512 score maps (region Gaussians per glyph on channel 0, affinity
Gaussians between neighbouring glyphs on channel 1) that stand in for a trained CRAFT's
output -- random-init weights give near-constant maps and a single component per page.
"""
from __future__ import annotations

import numpy as np

PAGE = 1280
MAP = 512
ROWS, COLS = 20, 15
WORDS_PER_PAGE = ROWS * COLS


def word_layout(page_index: int):
    """-> list of (row, col, n_chars, [char centre x in map coords], centre y in map coords)."""
    rng = np.random.default_rng(page_index)
    words = []
    for r in range(ROWS):
        for c in range(COLS):
            nch = int(rng.integers(2, 7))
            cy = 14 + 25 * r
            x0 = 8 + 33 * c
            xs = [x0 + 2 + 4.2 * i for i in range(nch)]
            words.append((r, c, nch, xs, float(cy)))
    return words, rng


def synth_page(page_index: int, size: int = PAGE) -> np.ndarray:
    """uint8 (size,size,3) page: white background, dark glyph boxes at 2.5x the map layout."""
    words, rng = word_layout(page_index)
    s = size / MAP
    img = np.full((size, size, 3), 255, np.uint8)
    for (_r, _c, _n, xs, cy) in words:
        for cx in xs:
            val = rng.integers(0, 65, size=3)
            x_lo, x_hi = int((cx - 1.6) * s), int((cx + 1.6) * s)
            y_lo, y_hi = int((cy - 5.0) * s), int((cy + 5.0) * s)
            img[max(y_lo, 0):min(y_hi, size), max(x_lo, 0):min(x_hi, size)] = val.astype(np.uint8)
    return img


def synth_score_maps(page_index: int, size: int = MAP) -> np.ndarray:
    """float32 (size,size,2): ch0 region score, ch1 affinity score, peak 1.0 -- the layout CRAFT
    emits ((H/2, W/2, 2) channels-last, tuatara.cpp:377-394)."""
    words, _ = word_layout(page_index)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    text = np.zeros((size, size), np.float32)
    link = np.zeros((size, size), np.float32)
    sx, sy = 2.2, 3.5
    for (_r, _c, _n, xs, cy) in words:
        x_lo, x_hi = int(max(xs[0] - 12, 0)), int(min(xs[-1] + 12, size))
        y_lo, y_hi = int(max(cy - 13, 0)), int(min(cy + 13, size))
        X, Y = xx[y_lo:y_hi, x_lo:x_hi], yy[y_lo:y_hi, x_lo:x_hi]
        for cx in xs:
            g = np.exp(-0.5 * (((X - cx) / sx) ** 2 + ((Y - cy) / sy) ** 2))
            text[y_lo:y_hi, x_lo:x_hi] = np.maximum(text[y_lo:y_hi, x_lo:x_hi], g)
        for a, b in zip(xs[:-1], xs[1:]):
            mx = 0.5 * (a + b)
            g = np.exp(-0.5 * (((X - mx) / 1.6) ** 2 + ((Y - cy) / 2.6) ** 2))
            link[y_lo:y_hi, x_lo:x_hi] = np.maximum(link[y_lo:y_hi, x_lo:x_hi], g)
    return np.ascontiguousarray(np.stack([text, link], -1).astype(np.float32))


def random_blob_maps(seed: int, h: int, w: int, density: float = 0.5) -> np.ndarray:
    """Blurred-noise score maps with many irregular components (CCL / min-area-box stress)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((h, w, 2), np.float32)
    for ch in range(2):
        a = rng.random((h + 8, w + 8)).astype(np.float32)
        k = 3 + 2 * int(rng.integers(0, 3))
        c = np.cumsum(np.cumsum(a, 0), 1)
        b = (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k])[:h, :w] / (k * k)
        b = (b - b.min()) / (b.max() - b.min() + 1e-9)
        out[..., ch] = b ** (1.0 + 2.0 * (1.0 - density))
    return out
