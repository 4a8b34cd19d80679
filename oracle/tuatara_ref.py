"""Line-by-line Python restatement of /root/reference/tuatara.cpp (the whole hot path).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the reference
lines it follows and calls the same third-party primitive (cv2 == OpenCV 4, torch CPU ==
ATen/LibTorch CPU) at the same call site, with the same literals.  float32 host
arithmetic is done with numpy float32 scalars so it rounds like the C++ ``float`` code.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import cv2
import numpy as np
import torch

f32 = np.float32


# ----------------------------------------------------------------------------- tokenizer
class Tokenizer:
    """tuatara.cpp:25-117.  The C++ literal at :32-34 is
    ``"...!\\"#$%&" "\\\\'()*+,-./:;<=>?@[\\\\]^_`{|}~"`` -- i.e. it contains a backslash
    *character* before the apostrophe (the author escaped a quote that needed no escape in
    a second literal starting with ``\\\\``), so the charset has 95 symbols, ``itos`` 98."""

    BOS, EOS, PAD = "[", "]", "P"  # tuatara.cpp:27-29

    def __init__(self):
        charset = ("0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!\"#$%&"
                   "\\'()*+,-./:;<=>?@[\\]^_`{|}~")  # :32-34 (same escapes as the C++ source)
        self.itos = self.EOS + charset + self.BOS + self.PAD  # :36-39
        self.stoi = {}
        for i, ch in enumerate(self.itos):  # :41-43 later duplicates win
            self.stoi[ch] = i
        self.eos_id = self.stoi[self.EOS]  # :45  -> 88 (the literal ']' inside the charset)
        self.bos_id = self.stoi[self.BOS]  # :46  -> 96
        self.pad_id = self.stoi[self.PAD]  # :47  -> 97

    def decode(self, token_dists: torch.Tensor) -> list[str]:
        """:61-78 with raw=false: max(-1) (:101-106), drop ids == eos_id (:108-116),
        ids -> chars (:93-99)."""
        out = []
        for i in range(token_dists.size(0)):
            _probs, ids = token_dists[i].max(-1)
            ids = ids[ids != self.eos_id]
            out.append("".join(self.itos[int(t)] for t in ids))
        return out


def truncate_at_eos(tokens: str, eos: str = "]") -> str:
    """tuatara.cpp:497-502: cut at the first EOS *character*."""
    k = tokens.find(eos)
    return tokens if k < 0 else tokens[:k]


# ------------------------------------------------------------------ score-map post-process
@dataclass
class DetDebug:
    labels: np.ndarray = None
    stats: np.ndarray = None
    n_labels: int = 0
    mapper: list = field(default_factory=list)
    text_norm: np.ndarray = None
    link_norm: np.ndarray = None
    points: list = field(default_factory=list)  # findNonZero output per kept component


def normalise_maps(textmap: torch.Tensor, linkmap: torch.Tensor):
    """tuatara.cpp:120-121 (ATen fp32 min/max/sub/div)."""
    t = (textmap - textmap.min()) / (textmap.max() - textmap.min())
    l = (linkmap - linkmap.min()) / (linkmap.max() - linkmap.min())
    return t, l


def get_detected_boxes(textmap: torch.Tensor, linkmap: torch.Tensor, text_threshold: float,
                       link_threshold: float, low_text: float, keep_points: bool = False, min_area: int = 10):
    """tuatara.cpp:119-204.  Returns (list of RotatedRect tuples ((cx,cy),(w,h),angle), DetDebug).
    The dead 'diamond'/start-corner code at :180-198 only mutates a local and is omitted."""
    tn, ln = normalise_maps(textmap.contiguous().float(), linkmap.contiguous().float())
    textmap_cv = np.ascontiguousarray(tn.numpy())  # :123-124
    linkmap_cv = np.ascontiguousarray(ln.numpy())
    img_h, img_w = textmap_cv.shape  # :126-127
    _, text_score = cv2.threshold(textmap_cv, low_text, 1, 0)  # :131
    _, link_score = cv2.threshold(linkmap_cv, link_threshold, 1, 0)  # :132
    comb = cv2.min(cv2.max(text_score + link_score, 0.0), 1.0)  # :136
    comb = comb.astype(np.uint8)  # :137 convertTo(CV_8U) of exact 0/1 values
    n_labels, labels, stats, _centroids = cv2.connectedComponentsWithStats(comb, connectivity=4)  # :142
    dbg = DetDebug(labels=labels, stats=stats, n_labels=n_labels, text_norm=textmap_cv, link_norm=linkmap_cv)
    det = []
    link_only = (link_score == 1) & (text_score == 0)  # :160
    for k in range(1, n_labels):  # :146
        size = int(stats[k, cv2.CC_STAT_AREA])
        if size < min_area:  # :147-148 (literal 10)
            continue
        mask = labels == k  # :150
        max_val = float(textmap_cv[mask].max())  # :151-152 minMaxLoc with mask
        # :154 `if (maxVal < text_threshold)`: double maxVal against the *float* parameter (0.7f, :397) promoted to
        # double = 0.699999988..., so a component whose maximum is exactly float32(0.7) is KEPT
        if max_val < float(np.float32(text_threshold)):
            continue
        segmap = np.zeros(textmap_cv.shape, np.uint8)  # :156
        segmap[mask] = 255  # :157
        dbg.mapper.append(k)  # :158
        segmap[link_only] = 0  # :160
        x, y = int(stats[k, cv2.CC_STAT_LEFT]), int(stats[k, cv2.CC_STAT_TOP])  # :162-163
        w, h = int(stats[k, cv2.CC_STAT_WIDTH]), int(stats[k, cv2.CC_STAT_HEIGHT])  # :164-165
        niter = int(math.sqrt((size * min(w, h)) // (w * h) * 2))  # :166 integer division
        sx, sy = max(0, x - niter), max(0, y - niter)  # :168-169
        ex, ey = min(img_w, x + w + niter + 1), min(img_h, y + h + niter + 1)  # :170-171
        kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (1 + niter, 1 + niter))  # :173
        roi = segmap[sy:ey, sx:ex]
        segmap[sy:ey, sx:ex] = cv2.dilate(roi, kernel)  # :174
        np_contours = cv2.findNonZero(segmap)  # :177-178
        rectangle = cv2.minAreaRect(np_contours)  # :179
        if keep_points:
            dbg.points.append(np_contours.reshape(-1, 2))
        det.append(rectangle)  # :200
    return det, dbg


# ------------------------------------------------------------------------- page preprocess
def resize_target(height: int, width: int, square_size: int, mag_ratio: float = 1.0):
    """The float32 size arithmetic of tuatara.cpp:211-220, 225-226."""
    target_size = f32(mag_ratio) * f32(max(height, width))  # :211
    if target_size > f32(square_size):  # :213-215
        target_size = f32(square_size)
    ratio = f32(target_size / f32(max(height, width)))  # :217
    target_h = int(f32(f32(height) * ratio))  # :219
    target_w = int(f32(f32(width) * ratio))  # :220
    th32 = target_h + (32 - target_h % 32) if target_h % 32 != 0 else target_h  # :225
    tw32 = target_w + (32 - target_w % 32) if target_w % 32 != 0 else target_w  # :226
    return target_h, target_w, th32, tw32, ratio


def resize_aspect_ratio(img: np.ndarray, square_size: int, interpolation: int, mag_ratio: float = 1.0):
    """tuatara.cpp:206-234."""
    height, width = img.shape[:2]
    target_h, target_w, th32, tw32, ratio = resize_target(height, width, square_size, mag_ratio)
    proc = cv2.resize(img, (target_w, target_h), interpolation=interpolation)  # :223
    resized = np.zeros((th32, tw32, img.shape[2]), img.dtype)  # :228
    resized[:target_h, :target_w] = proc  # :229
    return resized, ratio, (target_w // 2, target_h // 2)


# ----------------------------------------------------------------------------- box rescale
def rect_points(rect) -> np.ndarray:
    """cv::RotatedRect::points (tuatara.cpp:181,241,258) -> (4,2) float32."""
    (cx, cy), (w, h), a = rect
    return np.asarray(cv2.RotatedRect((float(cx), float(cy)), (float(w), float(h)), float(a)).points(), np.float32)


def rect_bounding(rect):
    """cv::RotatedRect::boundingRect (tuatara.cpp:416) -> (x, y, w, h) ints."""
    (cx, cy), (w, h), a = rect
    return tuple(int(v) for v in cv2.RotatedRect((float(cx), float(cy)), (float(w), float(h)), float(a)).boundingRect())


def adjust_result_coordinates(polys, ratio_w, ratio_h, ratio_net=2.0):
    """tuatara.cpp:236-253 (float32 multiplies, then minAreaRect of the 4 float corners)."""
    out = []
    sw = f32(f32(ratio_w) * f32(ratio_net))
    sh = f32(f32(ratio_h) * f32(ratio_net))
    for poly in polys:
        corners = rect_points(poly)  # :240-241
        corners[:, 0] = corners[:, 0] * sw  # :244
        corners[:, 1] = corners[:, 1] * sh  # :245
        out.append(cv2.minAreaRect(corners))  # :248
    return out


# ------------------------------------------------------------------------ output formatting
def _cround(v: float) -> float:
    """std::round: half away from zero."""
    return float(math.floor(abs(v) + 0.5) * (1.0 if v >= 0 else -1.0))


def rotated_rect_to_tesseract_format(rect) -> list[float]:
    """tuatara.cpp:256-274."""
    v = rect_points(rect)
    return [_cround(float(v[:, 0].min())), _cround(float(v[:, 1].min())),
            _cround(float(v[:, 0].max())), _cround(float(v[:, 1].max()))]


# --------------------------------------------------------------------------- crop + resize
def crop_rect(image_shape, rect, clamp: bool = True):
    """tuatara.cpp:416 ``image(box.boundingRect())``.  The reference throws when the rect
    leaves the image; with clamp=True the rect is intersected with the image instead (the
    one documented divergence).  Returns (x, y, w, h, clamped?)."""
    x, y, w, h = rect_bounding(rect)
    H, W = image_shape[:2]
    x0, y0, x1, y1 = max(x, 0), max(y, 0), min(x + w, W), min(y + h, H)
    clamped = (x0, y0, x1, y1) != (x, y, x + w, y + h)
    if clamped and not clamp:
        raise ValueError("boundingRect leaves the image (the reference would throw here)")
    return x0, y0, max(x1 - x0, 0), max(y1 - y0, 0), clamped


def crop_to_parseq_u8(image_swapped: np.ndarray, rect_xywh) -> np.ndarray:
    """tuatara.cpp:437-441: resize the ROI to 128x32 (INTER_LINEAR) and swap channels back."""
    x, y, w, h = rect_xywh
    if w <= 0 or h <= 0:  # rect entirely outside the image (the reference would throw): black crop
        return np.zeros((32, 128, 3), np.uint8)
    roi = image_swapped[y:y + h, x:x + w]
    r = cv2.resize(roi, (128, 32))  # :440
    return cv2.cvtColor(r, cv2.COLOR_BGR2RGB)  # :441


# ------------------------------------------------------------------------------ whole path
@dataclass
class Stages:
    craft_input_u8: np.ndarray = None  # padded, channel-swapped page (H32, W32, 3)
    ratio: float = 1.0
    score_text: np.ndarray = None
    score_link: np.ndarray = None
    det: list = None
    boxes: list = None
    crop_rects: list = None
    crops_u8: np.ndarray = None  # (N, 32, 128, 3) as fed to PARSeq (before /255)
    logits: np.ndarray = None
    det_debug: DetDebug = None


def run_parseq(parseq_model, crops_u8: np.ndarray, chunk_size: int = 4) -> torch.Tensor:
    """tuatara.cpp:443-486: tensorise (/255), chunks of <=4 (:452-459), forward per chunk
    (:307), re-order (:478), cat (:485).  The 6-thread pool (:461-475) only changes wall time."""
    t = torch.from_numpy(crops_u8).permute(0, 3, 1, 2).to(torch.float32).div(255.0)
    outs = []
    for i in range(0, t.shape[0], chunk_size):
        outs.append(parseq_model(t[i:i + chunk_size]))
    return torch.cat(outs, 0)


def image_to_data(image: np.ndarray, craft_model, parseq_model, score_override=None,
                  chunk_size: int = 4, stages: Stages | None = None, canvas_size: int = 1024, mag_ratio: float = 1.0,
                  text_threshold: float = 0.7, link_threshold: float = 0.4, low_text: float = 0.4, min_area: int = 10):
    """tuatara.cpp:314-512 with models passed in (the reference reloads them per call).

    ``image``: uint8 (H,W,3) exactly as the caller's buffer (BGR from cv::imread, RGB from
    bindings/run_ocr.py).  ``score_override``: optional (score_text, score_link) fp32 arrays
    that replace CRAFT's output (CRAFT still runs) -- needed because random-init weights give
    one giant component.  The keyword defaults are the reference's literals (:352-353, :397-399, :148);
    passing others is the config struct its TODO at :396 asks for.
    Returns list of dict(text=..., bbox=[4 floats])."""
    st = stages if stages is not None else Stages()
    image = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)  # :349 (in place in C++; swaps ch 0<->2)
    image_resized, target_ratio, _ = resize_aspect_ratio(image, canvas_size, cv2.INTER_LINEAR, mag_ratio)  # :352-358
    ratio_h = f32(1) / target_ratio  # :360
    ratio_w = f32(1) / target_ratio  # :361
    st.craft_input_u8, st.ratio = image_resized, float(target_ratio)
    x = torch.from_numpy(image_resized)[None].permute(0, 3, 1, 2).to(torch.float32).div(255.0)  # :363-370
    with torch.no_grad():
        y, _feature = craft_model(x)  # :376
    score_text = y[0, :, :, 0]  # :393
    score_link = y[0, :, :, 1]  # :394
    if score_override is not None:
        score_text = torch.from_numpy(np.ascontiguousarray(score_override[0]))
        score_link = torch.from_numpy(np.ascontiguousarray(score_override[1]))
    st.score_text, st.score_link = score_text.numpy().copy(), score_link.numpy().copy()
    det, st.det_debug = get_detected_boxes(score_text, score_link, text_threshold, link_threshold, low_text,
                                           min_area=min_area)  # :397-400
    boxes = adjust_result_coordinates(det, ratio_w, ratio_h)  # :406
    st.det, st.boxes = det, boxes
    if not boxes:  # the reference crashes in torch::cat({}) (:485); we return {}
        st.crop_rects, st.crops_u8 = [], np.zeros((0, 32, 128, 3), np.uint8)
        return []
    st.crop_rects = [crop_rect(image.shape, b)[:4] for b in boxes]  # :409-418
    st.crops_u8 = np.stack([crop_to_parseq_u8(image, r) for r in st.crop_rects])  # :437-441
    logits = run_parseq(parseq_model, st.crops_u8, chunk_size)  # :443-485
    st.logits = logits.numpy().copy()
    pred = torch.softmax(logits, -1)  # :486
    tok = Tokenizer()
    texts = [truncate_at_eos(t) for t in tok.decode(pred)]  # :492-505
    return [dict(text=t, bbox=rotated_rect_to_tesseract_format(b)) for t, b in zip(texts, boxes)]  # :511


# ------------------------------------------------------------------------------------------------
# Opt-in rectification (NOT reference behaviour: the reference crops the axis-aligned boundingRect, tuatara.cpp:416;
# its TODO at :411-415 asks for "perspective transform / rotated rectangle crop").  Test oracle of tt_config.rectify:
# the box's vertices in the order top-left, top-right, bottom-right, bottom-left, warped to 128 x 32 by OpenCV itself.
def rect_to_quad(rect) -> np.ndarray:
    pts = cv2.boxPoints(rect).astype(np.float32) if not hasattr(cv2, "RotatedRect") else np.asarray(
        cv2.RotatedRect(rect[0], rect[1], rect[2]).points(), np.float32)
    s = pts[:, 0] + pts[:, 1]
    d = pts[:, 1] - pts[:, 0]
    return np.stack([pts[np.argmin(s)], pts[np.argmin(d)], pts[np.argmax(s)], pts[np.argmax(d)]]).astype(np.float32)


def rectified_crop(image: np.ndarray, quad: np.ndarray) -> np.ndarray:
    dst = np.array([[0, 0], [127, 0], [127, 31], [0, 31]], np.float32)
    m = cv2.getPerspectiveTransform(np.asarray(quad, np.float32), dst)
    return cv2.warpPerspective(image, m, (128, 32), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
