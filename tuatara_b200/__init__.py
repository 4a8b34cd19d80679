"""tuatara_b200: B200-native implementation of the Tuatara OCR hot path.

Python here is only the host-side mirror of the reference's interface
(``bindings/python.cpp:43-58``: ``image_to_data(image, weights_dir, outputs_dir)``) plus thin
stage-level wrappers used by the parity tests and the benchmark.  Every call goes through
the C ABI of ``libtuatara_b200.so``; nothing is computed in Python and nothing falls back
to the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native
from ._native import TuataraError, check, lib, ptr

__all__ = [
    "TuataraError", "Engine", "image_to_data", "images_to_data", "default_config", "resize_plan", "preprocess",
    "postprocess", "crop_resize", "decode_ids", "tokenizer_table", "min_area_rect", "convex_hull",
    "rect_points", "rect_bounding", "adjust_rect", "rect_to_bbox", "launch_count",
]


def default_config() -> _native.tt_config:
    cfg = _native.tt_config()
    lib().tt_config_default(C.byref(cfg))
    return cfg


def launch_count() -> int:
    return int(lib().tt_launch_count())


# --------------------------------------------------------------------------- host logic
def resize_plan(rows: int, cols: int, canvas_size: float = 1024.0, mag_ratio: float = 1.0):
    """-> (target_h, target_w, h32, w32, ratio)  (tuatara.cpp:211-226)."""
    th, tw, h32, w32, ratio = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_float()
    check(lib().tt_resize_plan(rows, cols, canvas_size, mag_ratio, C.byref(th), C.byref(tw), C.byref(h32),
                               C.byref(w32), C.byref(ratio)), "tt_resize_plan")
    return th.value, tw.value, h32.value, w32.value, np.float32(ratio.value)


def tokenizer_table():
    buf = C.create_string_buffer(128)
    eos, bos, pad = C.c_int(), C.c_int(), C.c_int()
    check(lib().tt_tokenizer_table(buf, C.byref(eos), C.byref(bos), C.byref(pad)), "tt_tokenizer_table")
    return buf.value.decode("latin-1"), eos.value, bos.value, pad.value


def decode_ids(ids: np.ndarray) -> list[str]:
    ids = np.ascontiguousarray(ids, np.int32)
    n, ln = ids.shape
    out = np.zeros((n, 64), np.uint8)
    check(lib().tt_decode(ptr(ids), n, ln, ptr(out), 64), "tt_decode")
    return [bytes(row).split(b"\0", 1)[0].decode("latin-1") for row in out]


def convex_hull(points: np.ndarray) -> np.ndarray:
    n = len(points)
    idx = np.zeros(max(n, 1), np.int32)
    n_out = C.c_int()
    if points.dtype.kind == "f":
        p = np.ascontiguousarray(points, np.float32)
        check(lib().tt_convex_hull_f32(ptr(p), n, ptr(idx), C.byref(n_out)), "tt_convex_hull_f32")
    else:
        p = np.ascontiguousarray(points, np.int32)
        check(lib().tt_convex_hull_i32(ptr(p), n, ptr(idx), C.byref(n_out)), "tt_convex_hull_i32")
    return idx[: n_out.value].copy()


def _rect_tuple(r: np.ndarray):
    return ((float(r[0]), float(r[1])), (float(r[2]), float(r[3])), float(r[4]))


def _rect_array(rect) -> np.ndarray:
    (cx, cy), (w, h), a = rect
    return np.array([cx, cy, w, h, a], np.float32)


def min_area_rect(points: np.ndarray):
    """cv2.minAreaRect-compatible: ((cx, cy), (w, h), angle)."""
    out = np.zeros(5, np.float32)
    if points.dtype.kind == "f":
        p = np.ascontiguousarray(points, np.float32).reshape(-1, 2)
        check(lib().tt_min_area_rect_f32(ptr(p), len(p), ptr(out)), "tt_min_area_rect_f32")
    else:
        p = np.ascontiguousarray(points, np.int32).reshape(-1, 2)
        check(lib().tt_min_area_rect_i32(ptr(p), len(p), ptr(out)), "tt_min_area_rect_i32")
    return _rect_tuple(out)


def rect_points(rect) -> np.ndarray:
    out = np.zeros((4, 2), np.float32)
    check(lib().tt_rect_points(ptr(_rect_array(rect)), ptr(out)), "tt_rect_points")
    return out


def rect_bounding(rect):
    out = np.zeros(4, np.int32)
    check(lib().tt_rect_bounding(ptr(_rect_array(rect)), ptr(out)), "tt_rect_bounding")
    return tuple(int(v) for v in out)


def adjust_rect(rect, ratio_w, ratio_h, ratio_net=2.0):
    out = np.zeros(5, np.float32)
    check(lib().tt_adjust_rect(ptr(_rect_array(rect)), float(ratio_w), float(ratio_h), float(ratio_net), ptr(out)),
          "tt_adjust_rect")
    return _rect_tuple(out)


def rect_to_bbox(rect) -> list[float]:
    out = np.zeros(4, np.float32)
    check(lib().tt_rect_to_bbox(ptr(_rect_array(rect)), ptr(out)), "tt_rect_to_bbox")
    return [float(v) for v in out]


# ------------------------------------------------------------------- GPU stages (host buffers)
def preprocess(image: np.ndarray, canvas_size: float = 1024.0, mag_ratio: float = 1.0):
    """Channel swap + resize + pad on the GPU -> (uint8 [h32, w32, 3], ratio)."""
    _th, _tw, h32, w32, ratio = resize_plan(image.shape[0], image.shape[1], canvas_size, mag_ratio)
    out = np.empty((h32, w32, 3), np.uint8)
    im = _native.image_struct(image)
    check(lib().tt_preprocess(C.byref(im), canvas_size, mag_ratio, ptr(out)), "tt_preprocess")
    return out, ratio


def postprocess(maps: np.ndarray, cfg: _native.tt_config | None = None, want_labels: bool = True):
    """get_detected_boxes on the GPU.  maps: float32 [H, W, 2].
    -> dict(n_labels, labels [H,W] int32, stats [n_labels,5] int32, rects [list of cv2-style tuples], rect_labels)."""
    maps = np.ascontiguousarray(maps, np.float32)
    H, W, _ = maps.shape
    cap = H * W // 2 + 2
    labels = np.empty((H, W), np.int32) if want_labels else None
    stats = np.zeros((cap, 5), np.int32)
    rects = np.zeros((cap, 5), np.float32)
    rlabels = np.zeros(cap, np.int32)
    n_labels, n_rects = C.c_int(), C.c_int()
    check(lib().tt_postprocess(ptr(maps), H, W, C.byref(cfg) if cfg is not None else None, ptr(labels), ptr(stats),
                               cap, C.byref(n_labels), ptr(rects), ptr(rlabels), cap, C.byref(n_rects)),
          "tt_postprocess")
    return dict(n_labels=n_labels.value, labels=labels, stats=stats[: n_labels.value].copy(),
                rects=[_rect_tuple(r) for r in rects[: n_rects.value]],
                rect_labels=rlabels[: n_rects.value].copy())


def crop_resize(image: np.ndarray, rects_xywh) -> np.ndarray:
    """Batched crop + 128x32 bilinear resize on the GPU -> uint8 [n, 32, 128, 3]."""
    r = np.ascontiguousarray(np.asarray(rects_xywh, np.int32).reshape(-1, 4))
    out = np.empty((len(r), 32, 128, 3), np.uint8)
    im = _native.image_struct(image)
    check(lib().tt_crop_resize(C.byref(im), ptr(r), len(r), ptr(out)), "tt_crop_resize")
    return out


def rect_to_quad(rect) -> np.ndarray:
    """The quadrilateral tt_config.rectify warps for a RotatedRect: float32 [4, 2] (tl, tr, br, bl)."""
    out = np.zeros(8, np.float32)
    check(lib().tt_rect_to_quad(ptr(_rect_array(rect)), ptr(out)), "tt_rect_to_quad")
    return out.reshape(4, 2)


def crop_warp(image: np.ndarray, quads) -> np.ndarray:
    """Rectified crops on the GPU (cv2.getPerspectiveTransform + cv2.warpPerspective semantics) -> uint8 [n, 32, 128, 3]."""
    q = np.ascontiguousarray(np.asarray(quads, np.float32).reshape(-1, 8))
    out = np.empty((len(q), 32, 128, 3), np.uint8)
    im = _native.image_struct(image)
    check(lib().tt_crop_warp(C.byref(im), ptr(q), len(q), ptr(out)), "tt_crop_warp")
    return out


# --------------------------------------------------------------------------------- engine
class Engine:
    """Weights resident on the GPU(s) + streams + workspaces (tt_engine_*)."""

    def __init__(self, weights_dir: str, devices=None, cfg: _native.tt_config | None = None):
        """devices: list of CUDA device indices, "all", or None (= the TT_DEVICES environment variable, else device 0)."""
        self._h = C.c_void_p()
        dev = None
        n = 0
        if isinstance(devices, str) and devices == "all":
            devices = list(range(lib().tt_device_count()))
        if devices is not None:
            n = len(devices)
            dev = (C.c_int * n)(*devices)
        check(lib().tt_engine_create(str(weights_dir).encode(), dev, n, C.byref(cfg) if cfg is not None else None,
                                     C.byref(self._h)), "tt_engine_create")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().tt_engine_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def ocr_pages(self, images: list, score_override: list | None = None, detect_only: bool = False) -> list[list[dict]]:
        """images: uint8 [H,W,3] numpy arrays (host) or torch CUDA tensors (already on the device).
        score_override: optional per-page float32 [h32/2, w32/2, 2] maps (numpy or CUDA tensors, or
        None entries) that replace CRAFT's output after CRAFT has run (benchmark / parity tests)."""
        n = len(images)
        on_dev = n > 0 and not isinstance(images[0], np.ndarray)
        if on_dev:
            structs = [_native.tt_image(t.data_ptr(), t.shape[0], t.shape[1], t.shape[2], t.stride(0)) for t in images]
        else:
            structs = [_native.image_struct(im) for im in images]
        arr = (_native.tt_image * n)(*structs)
        if n and any(isinstance(im, np.ndarray) == on_dev for im in images):
            raise TuataraError("ocr_pages: host and device pages cannot be mixed in one call")
        opt = _native.tt_ocr_options(int(on_dev), 0, None, int(detect_only))
        keep = []
        if score_override is not None:
            ov_dev = any(m is not None and not isinstance(m, np.ndarray) for m in score_override)
            ptrs = (C.c_void_p * n)()
            for i, m in enumerate(score_override):
                if m is None:
                    ptrs[i] = None
                    continue
                if isinstance(m, np.ndarray):
                    m = np.ascontiguousarray(m, np.float32)
                keep.append(m)
                ptrs[i] = ptr(m)
            opt.override_on_device = int(ov_dev)
            opt.score_override = ptrs
        res = C.POINTER(_native.tt_result)()
        check(lib().tt_ocr_pages_ex(self._h, arr, n, C.byref(opt), C.byref(res)), "tt_ocr_pages_ex")
        try:
            out = []
            for p in range(res.contents.n_pages):
                page = res.contents.pages[p]
                out.append([dict(text=page.items[i].text.decode("latin-1"), bbox=[float(v) for v in page.items[i].bbox])
                            for i in range(page.n_items)])
            return out
        finally:
            lib().tt_result_free(res)

    def craft_forward(self, craft_input: np.ndarray) -> np.ndarray:
        """craft_input: uint8 [h32, w32, 3] from preprocess() -> float32 [h32/2, w32/2, 2]."""
        x = np.ascontiguousarray(craft_input, np.uint8)
        h32, w32, _ = x.shape
        out = np.empty((h32 // 2, w32 // 2, 2), np.float32)
        check(lib().tt_craft_forward(self._h, ptr(x), h32, w32, ptr(out)), "tt_craft_forward")
        return out

    def craft_tap(self, name: str) -> np.ndarray:
        """A named activation of the last craft_forward() as float32 [H, W, C] (per-slice parity tests)."""
        dims = (C.c_int * 3)()
        check(lib().tt_craft_tap(self._h, name.encode(), None, 0, dims), "tt_craft_tap")
        out = np.empty((dims[0], dims[1], dims[2]), np.float32)
        check(lib().tt_craft_tap(self._h, name.encode(), ptr(out), out.size, dims), "tt_craft_tap")
        return out

    def parseq_forward(self, crops_u8: np.ndarray, forced_tokens: np.ndarray | None = None):
        """crops_u8: uint8 [n, 32, 128, 3] -> (logits float32 [n, 26, 95], ids int32 [n, 26])."""
        x = np.ascontiguousarray(crops_u8, np.uint8)
        n = x.shape[0]
        logits = np.empty((n, 26, 95), np.float32)
        ids = np.empty((n, 26), np.int32)
        ft = None if forced_tokens is None else np.ascontiguousarray(forced_tokens, np.int32)
        check(lib().tt_parseq_forward(self._h, ptr(x), n, ptr(ft), ptr(logits), ptr(ids)), "tt_parseq_forward")
        return logits, ids


_engines: dict = {}


def _engine_for(weights_dir: str) -> Engine:
    key = str(weights_dir)
    if key not in _engines:
        import os
        # like include/tuatara.h: TT_DEVICES when set, else every visible GPU
        _engines[key] = Engine(key, devices=None if os.environ.get("TT_DEVICES") else "all")
    return _engines[key]


def image_to_data(image: np.ndarray, weights_dir: str, outputs_dir: str) -> list[dict]:
    """Same call as the reference's ``pytuatara.image_to_data`` (bindings/python.cpp:43-57):
    uint8 [H, W, 3] in, list of ``dict(text=str, bbox=[4 floats])`` out.  Empty directory
    strings print to stderr and return [] like tuatara.cpp:315-323."""
    import sys

    if not weights_dir:
        print("Please provide a value for weights_dir", file=sys.stderr)
        return []
    if not outputs_dir:
        print("Please provide a value for outputs_dir", file=sys.stderr)
        return []
    if image.ndim != 3:
        raise RuntimeError("Input array should have 3 dimensions")
    return _engine_for(weights_dir).ocr_pages([image])[0]


def images_to_data(images: list[np.ndarray], weights_dir: str, outputs_dir: str = ".") -> list[list[dict]]:
    """Additive batch API: one call, many pages (sharded over the engine's GPUs)."""
    return _engine_for(weights_dir).ocr_pages(list(images))
