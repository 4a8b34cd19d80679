"""Generates the committed golden fixtures under tests/golden/ (run from the repo root:
``python tests/golden/make_golden.py``).

The reference ships no machine-readable vectors, so these are outputs of
the third-party code the reference delegates to -- cv2 4.13.0 (OpenCV) and torch 2.11 CPU (ATen)
-- captured at the reference's call sites through the oracle (oracle/tuatara_ref.py), plus the
tokenizer table printed by a g++-compiled restatement of tuatara.cpp:25-48.  They pin the oracle
against drift (another cv2/torch build) and give the GPU tests fixed vectors that do not depend
on cv2 being importable.
"""
import subprocess
import sys
from pathlib import Path

import cv2
import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import tuatara_ref as R  # noqa: E402
from oracle.models import make_craft, make_parseq  # noqa: E402
from tuatara_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def rects_array(det):
    return np.array([[r[0][0], r[0][1], r[1][0], r[1][1], r[2]] for r in det], np.float32).reshape(-1, 5)


def main():
    print("cv2", cv2.__version__, "torch", torch.__version__)
    # 1. tokenizer table from the compiled restatement of the reference constructor
    exe = ROOT / "oracle" / "_ref" / "tokenizer_kat"
    exe.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-O1", "-o", str(exe), str(ROOT / "oracle" / "tokenizer_kat.cpp")], check=True)
    (OUT / "tokenizer_kat.txt").write_text(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)

    # 2. post-processing: a crop of a synthetic score map and a blob map
    post = {}
    for name, maps in (("synth", synth.synth_score_maps(0)[:128, :160].copy()),
                       ("blobs", synth.random_blob_maps(11, 96, 120))):
        det, dbg = R.get_detected_boxes(torch.from_numpy(maps[..., 0].copy()), torch.from_numpy(maps[..., 1].copy()),
                                        0.7, 0.4, 0.4)
        post[name + "_maps"] = maps.astype(np.float32)
        post[name + "_labels"] = dbg.labels.astype(np.int32)
        post[name + "_stats"] = dbg.stats[:, :5].astype(np.int32)
        post[name + "_rects"] = rects_array(det)
        post[name + "_mapper"] = np.array(dbg.mapper, np.int32)
        inv = np.float32(1) / np.float32(0.8)
        adj = R.adjust_result_coordinates(det, inv, inv)
        post[name + "_adjusted"] = rects_array(adj)
        post[name + "_bbox"] = np.array([R.rotated_rect_to_tesseract_format(b) for b in adj], np.float32).reshape(-1, 4)
        post[name + "_bounding"] = np.array([R.rect_bounding(b) for b in adj], np.int32).reshape(-1, 4)
    np.savez_compressed(OUT / "postprocess.npz", **post)

    # 3. resize: page preprocess and crops
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (150, 131, 3), dtype=np.uint8)
    big = rng.integers(0, 256, (1128, 300, 3), dtype=np.uint8)  # long side 1128 -> 1023 (float32 quirk)
    pre_small, ratio_small, _ = R.resize_aspect_ratio(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), 1024, cv2.INTER_LINEAR)
    pre_big, ratio_big, _ = R.resize_aspect_ratio(cv2.cvtColor(big, cv2.COLOR_BGR2RGB), 1024, cv2.INTER_LINEAR)
    rects = np.array([(0, 0, 131, 150), (5, 7, 40, 13), (60, 20, 64, 16), (3, 100, 128, 32), (100, 100, 1, 1),
                      (10, 10, 100, 120)], np.int32)
    crops = np.stack([R.crop_to_parseq_u8(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), tuple(r)) for r in rects])
    np.savez_compressed(OUT / "resize.npz", img=img, big=big, pre_small=pre_small, pre_big=pre_big,
                        ratio_small=np.float32(ratio_small), ratio_big=np.float32(ratio_big), rects=rects, crops=crops)

    # 4. the fp32 oracle networks on tiny inputs (guards the oracle's graphs / seeded init)
    craft, parseq = make_craft(0), make_parseq("base", 0)
    x = torch.from_numpy(synth.synth_page(0)[:64, :96].copy())
    xin = x[None].permute(0, 3, 1, 2).float().div(255.0)
    with torch.no_grad():
        maps = craft(xin)[0][0].numpy()
    crops_u8 = np.stack([cv2.resize(synth.synth_page(1)[y:y + 30, xx:xx + 90], (128, 32)) for y, xx in ((20, 15), (300, 500))])
    logits = parseq(torch.from_numpy(crops_u8).permute(0, 3, 1, 2).float().div(255.0)).numpy()
    np.savez_compressed(OUT / "nets.npz", craft_in=x.numpy(), craft_maps=maps, crops_u8=crops_u8, logits=logits)
    print("written:", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    main()
