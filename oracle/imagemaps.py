"""Deterministic score maps for REAL page images (test infrastructure, like everything under oracle/).

Random-init CRAFT weights give near-constant maps, so parity runs on the reference's fixture images
(images/*.png) override CRAFT's output with maps derived from the page's own ink: blurred darkness = region score,
its horizontal smear = affinity score.  The maps are quantised to 8 bits so that a committed fixture (uint8) and a
map recomputed on another machine are the same float32 values (u8 / 255)."""
import cv2
import numpy as np


def ink_maps_u8(craft_input_u8: np.ndarray) -> np.ndarray:
    """craft_input_u8: the (h32, w32, 3) uint8 CRAFT input of a page -> uint8 [h32/2, w32/2, 2] (region, affinity)."""
    h32, w32, _ = craft_input_u8.shape
    gray = cv2.cvtColor(craft_input_u8, cv2.COLOR_RGB2GRAY)
    ink = (255.0 - gray.astype(np.float32)) / 255.0
    ink[gray == 0] = 0.0  # the zero padding of resize_aspect_ratio is not ink
    small = cv2.resize(ink, (w32 // 2, h32 // 2), interpolation=cv2.INTER_AREA)
    region = np.minimum(cv2.GaussianBlur(small, (0, 0), 0.7) * 1.6, 1.0)  # saturating gain: most strokes reach the 0.7 peak
    link = np.minimum(cv2.blur(region, (3, 1)) * 0.8, 1.0)
    maps = np.stack([region, link], -1)
    return np.clip(np.rint(maps * 255.0), 0, 255).astype(np.uint8)


def maps_f32(maps_u8: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(maps_u8.astype(np.float32) / np.float32(255.0))
