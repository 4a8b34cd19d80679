"""Golden fixtures from the reference's own fixture images (/root/reference/images/*.png; SURVEY 8c pin iv).

Run here (the container that has /root/reference):  python tests/golden/make_golden_images.py
For each page: the decoded pixels exactly as the reference's C++ callers hand them over (cv::imread -> BGR,
examples/resume.cpp:9), 8-bit score maps derived from the page's ink (oracle/imagemaps.py), and what the oracle
restatement of tuatara.cpp returns for them: adjusted RotatedRects, Tesseract-style bboxes, crop rectangles and the
first PARSeq input crops.  tests/test_fixture_images.py checks the oracle (CPU) and the CUDA path (GPU) against it."""
import sys
from pathlib import Path

import cv2
import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import imagemaps, tuatara_ref as R  # noqa: E402

NAMES = ["resume_example", "funsd_0001129658", "funsd_91372360", "table_english", "rotated_text"]


def run_oracle(img_bgr, maps_u8):
    m = imagemaps.maps_f32(maps_u8)
    craft = lambda x: (torch.zeros(1, x.shape[2] // 2, x.shape[3] // 2, 2), None)  # noqa: E731  (output is overridden)
    parseq = lambda x: torch.zeros(x.shape[0], 26, 95)  # noqa: E731
    st = R.Stages()
    items = R.image_to_data(img_bgr.copy(), craft, parseq, score_override=(m[..., 0], m[..., 1]), stages=st)
    return items, st


def main():
    out = {}
    for name in NAMES:
        img = cv2.imread(f"/root/reference/images/{name}.png", cv2.IMREAD_COLOR)
        craft_in, *_ = R.resize_aspect_ratio(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), 1024, cv2.INTER_LINEAR, 1.0)
        maps_u8 = imagemaps.ink_maps_u8(craft_in)
        items, st = run_oracle(img, maps_u8)
        n = len(items)
        out[f"{name}.img"] = img
        out[f"{name}.maps_u8"] = maps_u8
        out[f"{name}.bbox"] = np.array([it["bbox"] for it in items], np.float32).reshape(n, 4)
        out[f"{name}.crop_rects"] = np.array(st.crop_rects, np.int32).reshape(n, 4)
        out[f"{name}.crops"] = st.crops_u8[:8]
        print(name, img.shape, "->", craft_in.shape, n, "boxes")
    np.savez_compressed(ROOT / "tests" / "golden" / "fixture_images.npz", **out)


if __name__ == "__main__":
    main()
