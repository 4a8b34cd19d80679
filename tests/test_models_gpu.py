"""CRAFT / PARSeq on the GPU (bf16 operands, fp32 accumulation, BatchNorm folded) vs the fp32
torch-CPU oracle with the same seeded random-init weights.

Declared tolerances: relative L2 error <= 3e-2 on the RAW score maps and on the
logits (never on min-max-normalised maps); decoded ids must match wherever the oracle's top-1 /
top-2 logit margin exceeds 0.5 at that position and every earlier one (teacher-forced run) ."""
import cv2
import numpy as np
import pytest
import torch

import tuatara_b200 as tb
from oracle import tuatara_ref as R
from tuatara_b200 import synth

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("kind", ["noise_608x768", "page_1280"])
def test_craft_score_maps(engine, oracle_models, kind):
    craft, _ = oracle_models
    if kind == "page_1280":
        img = synth.synth_page(0)
    else:
        rng = np.random.default_rng(3)
        img = rng.integers(0, 256, (763, 607, 3), dtype=np.uint8)
    craft_in, _ratio = tb.preprocess(img)
    x = torch.from_numpy(craft_in)[None].permute(0, 3, 1, 2).float().div(255.0)
    with torch.no_grad():
        ref = craft(x)[0][0].numpy()
    got = engine.craft_forward(craft_in)
    assert got.shape == ref.shape
    assert np.isfinite(got).all()
    err = _rel_l2(got, ref)
    print(kind, "rel-L2", err, "max-abs", float(np.abs(got - ref).max()), "ref range", float(ref.min()), float(ref.max()))
    assert err <= 3e-2, f"raw score maps rel-L2 {err:.4f}"


def test_craft_per_slice_parity(engine, oracle_models):
    """every CRAFT slice against the fp32 oracle, so a regression names its layer.  Bars (bf16 operands,
    fp32 accumulate, BN folded): rel-L2 <= 1e-2 on the VGG slices and fc7, <= 2e-2 on the U-net stages (each stacks
    two more convolutions and a bilinear upsample on the slices), final maps <= 3e-2 (test_craft_score_maps)."""
    craft, _ = oracle_models
    craft_in, _ = tb.preprocess(synth.synth_page(3)[:640, :768])
    x = torch.from_numpy(craft_in)[None].permute(0, 3, 1, 2).float().div(255.0)
    taps = {}
    with torch.no_grad():
        craft(x, taps=taps)
    engine.craft_forward(craft_in)
    bars = {"relu2_2": 1e-2, "relu3_2": 1e-2, "relu4_3": 1e-2, "relu5_3": 1e-2, "fc7": 1e-2,
            "up1": 2e-2, "up2": 2e-2, "up3": 2e-2, "up4": 2e-2}
    worst = {}
    for name, bar in bars.items():
        ref = taps[name][0].permute(1, 2, 0).numpy()
        got = engine.craft_tap(name)
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        worst[name] = _rel_l2(got, ref)
    print("CRAFT per-slice rel-L2:", {k: round(v, 5) for k, v in worst.items()})
    for name, bar in bars.items():
        assert worst[name] <= bar, f"{name}: rel-L2 {worst[name]:.4f} > {bar}"


def test_craft_fused_maxpool_is_bit_identical(engine, monkeypatch):
    """The 2x2 max-pools run in the epilogues of conv1_2 / conv2_2 / conv3_3 / conv4_3 (halo and per-tap schedules, single
    CTAs and CTA pairs, pooled-only and pooled + skip outputs).  Pooling the bf16-rounded values is what the separate
    kernel did, so every activation and the maps must keep their bits; sizes with partial edge tiles included."""
    for h, w in ((640, 768), (1024, 1024), (608, 352)):
        craft_in, _ = tb.preprocess(synth.synth_page(6)[:h, :w])
        fused = engine.craft_forward(craft_in)
        taps_f = {k: engine.craft_tap(k) for k in ("relu2_2", "relu3_2", "relu4_3", "relu5_3")}
        monkeypatch.setenv("TT_CRAFT_POOLFUSE", "0")
        plain = engine.craft_forward(craft_in)
        taps_p = {k: engine.craft_tap(k) for k in taps_f}
        monkeypatch.delenv("TT_CRAFT_POOLFUSE")
        for k in taps_f:
            assert np.array_equal(taps_f[k], taps_p[k]), (h, w, k)
        assert np.array_equal(fused, plain), (h, w)


def test_craft_conv1_1_from_u8_matches_the_gemm_path(engine, monkeypatch):
    """conv1_1 straight from the u8 page (mma.sync, A fragments built from a smem halo) against the round-1 path (im2col
    tensor + K = 32 tcgen05 GEMM): the same bf16 products accumulated in fp32 in another order, so relu2_2 -- two
    convolutions downstream -- agrees to rounding noise, and the maps stay inside their oracle tolerance either way."""
    for h, w in ((640, 768), (608, 352)):
        craft_in, _ = tb.preprocess(synth.synth_page(8)[:h, :w])
        new = engine.craft_forward(craft_in)
        t_new = engine.craft_tap("relu2_2")
        monkeypatch.setenv("TT_CRAFT_C11", "0")
        old = engine.craft_forward(craft_in)
        t_old = engine.craft_tap("relu2_2")
        monkeypatch.delenv("TT_CRAFT_C11")
        e_tap, e_map = _rel_l2(t_new, t_old), _rel_l2(new, old)
        print("conv1_1 u8 vs gemm path: relu2_2 rel-L2", e_tap, "maps rel-L2", e_map)
        assert e_tap <= 2e-3 and e_map <= 1e-2


def _crops(n, seed=0):
    img = synth.synth_page(seed)
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        w, h = int(rng.integers(20, 200)), int(rng.integers(10, 60))
        x, y = int(rng.integers(0, 1280 - w)), int(rng.integers(0, 1280 - h))
        out.append(cv2.resize(img[y:y + h, x:x + w], (128, 32)))
    noise = rng.integers(0, 256, (n // 2, 32, 128, 3), dtype=np.uint8)
    return np.concatenate([np.stack(out), noise])


def test_parseq_logits_teacher_forced(engine, oracle_models, monkeypatch):
    _, parseq = oracle_models
    crops = _crops(32)
    x = torch.from_numpy(crops).permute(0, 3, 1, 2).float().div(255.0)
    taps = {}
    ref_free = parseq(x, taps=taps)
    forced = taps["ar_tokens"][:, 1:].clone()  # the oracle's own AR context
    ref = parseq(x, forced_tokens=forced).numpy()
    assert np.allclose(ref, ref_free.numpy(), atol=1e-5)  # forcing its own tokens changes nothing
    got, ids = engine.parseq_forward(crops, forced.numpy().astype(np.int32))
    assert np.isfinite(got).all()
    err = _rel_l2(got, ref)
    print("logits rel-L2", err, "max-abs", float(np.abs(got - ref).max()))
    assert err <= 3e-2, f"logits rel-L2 {err:.4f}"
    # ids must agree wherever the oracle's decision is clear
    top2 = np.sort(ref, -1)[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 0.5
    ref_ids = ref.argmax(-1)
    assert (ids[clear] == ref_ids[clear]).all(), f"{int((ids[clear] != ref_ids[clear]).sum())} clear positions differ"
    assert clear.mean() > 0.5, "margin test is vacuous"
    # the AR pass itself (fused decoder kernels) against the oracle's AR logits under the same forced context
    monkeypatch.setenv("TT_PARSEQ_AR_LOGITS", "1")
    got_ar, _ = engine.parseq_forward(crops, forced.numpy().astype(np.int32))
    monkeypatch.delenv("TT_PARSEQ_AR_LOGITS")
    ref_ar = taps["ar_logits"].numpy()
    err_ar = _rel_l2(got_ar, ref_ar)
    print("AR logits rel-L2", err_ar)
    assert err_ar <= 3e-2, f"AR logits rel-L2 {err_ar:.4f}"


def test_parseq_free_running_strings(engine, oracle_models):
    """No forcing: strings must match for crops whose every AR + refinement decision up to the end of the
    decoded string is clear (margin > 1.0 in the oracle); report the rest."""
    _, parseq = oracle_models
    crops = _crops(32, seed=1)
    x = torch.from_numpy(crops).permute(0, 3, 1, 2).float().div(255.0)
    taps = {}
    ref = parseq(x, taps=taps).numpy()
    ar = taps["ar_logits"].numpy()
    got, ids = engine.parseq_forward(crops)
    tok = R.Tokenizer()
    ref_txt = [R.truncate_at_eos(t) for t in tok.decode(torch.from_numpy(ref))]
    got_txt = tb.decode_ids(ids)

    def margins(l):
        t = np.sort(l, -1)[..., -2:]
        return t[..., 1] - t[..., 0]

    clear = (margins(ar) > 1.0).all(-1) & (margins(ref) > 1.0).all(-1)
    same = np.array([a == b for a, b in zip(ref_txt, got_txt)])
    print("clear crops", int(clear.sum()), "of", len(clear), "; strings equal", int(same.sum()))
    assert same[clear].all()
    assert same.mean() >= 0.8


def test_parseq_fused_decoder_matches_unfused(engine, monkeypatch):
    """The fused AR-step kernels (dec_fused.cu: the residual row in TMEM, LayerNorm / GELU in registers, chained
    tcgen05 GEMMs) against the same steps run as separate GEMM / LayerNorm launches (TT_DEC_FUSED=0): same bf16
    operands and fp32 accumulation, so logits agree to rounding noise and every clear decision is identical.
    Ragged sizes cover a partial last 128-crop tile and the single-tile case."""
    for n, seed in ((300, 2), (77, 3), (128, 4)):
        crops = _crops(n + (n & 1), seed=seed)[:n]
        _, id0 = engine.parseq_forward(crops)
        forced = np.ascontiguousarray(id0[:, :25]).astype(np.int32)  # one fixed AR context for both paths
        monkeypatch.setenv("TT_PARSEQ_AR_LOGITS", "1")  # compare the AR pass itself (the refinement only sees its tokens)
        monkeypatch.setenv("TT_DEC_FUSED", "1")
        lf, _ = engine.parseq_forward(crops, forced)
        monkeypatch.setenv("TT_DEC_FUSED", "0")
        lu, _ = engine.parseq_forward(crops, forced)
        monkeypatch.delenv("TT_DEC_FUSED")
        monkeypatch.delenv("TT_PARSEQ_AR_LOGITS")
        assert np.isfinite(lf).all()
        err = _rel_l2(lf, lu)
        top2 = np.sort(lu, -1)[..., -2:]
        clear = (top2[..., 1] - top2[..., 0]) > 0.25
        idf, idu = lf.argmax(-1), lu.argmax(-1)
        print(n, "fused vs unfused AR logits rel-L2", err, "max-abs", float(np.abs(lf - lu).max()), "ids equal", float((idf == idu).mean()))
        assert 0 < err <= 5e-3, f"fused decoder AR logits rel-L2 {err:.5f}"
        assert (idf[clear] == idu[clear]).all()
        # free running: both paths decode the same strings wherever every decision is clear
        monkeypatch.setenv("TT_DEC_FUSED", "0")
        _, id_u = engine.parseq_forward(crops)
        monkeypatch.delenv("TT_DEC_FUSED")
        same = (id0 == id_u).all(-1)
        print(n, "free-running id rows equal", int(same.sum()), "of", n)
        assert same.mean() >= 0.95


def test_parseq_early_exit_is_output_preserving(engine, monkeypatch):
    """Per-crop early exit of the AR loop (a crop leaves once its step produced EOS; upstream PARSeq breaks a batch when
    every sequence has one, and the reference feeds it 4 crops at a time, tuatara.cpp:452-475) against the full 26-step
    schedule (TT_DEC_EARLY_EXIT=0): nothing after a crop's first EOS reaches the refinement pass or the string, so
    the refinement logits and ids are bit-identical at EVERY position, and the AR logits up to the EOS step are too.
    Sizes: one tile, ragged, two half-batches on two streams (>= 512), and a batch where the lists shrink to nothing."""
    eos = 0
    for n, seed in ((128, 11), (77, 12), (1000, 13)):
        crops = _crops(n + (n & 1), seed=seed)[:n]
        if n == 1000:
            crops[500:] = np.random.default_rng(seed).integers(0, 256, crops[500:].shape, dtype=np.uint8)
        res = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("TT_DEC_EARLY_EXIT", mode)
            l, i = engine.parseq_forward(crops)
            monkeypatch.setenv("TT_PARSEQ_AR_LOGITS", "1")
            la, _ = engine.parseq_forward(crops)
            monkeypatch.delenv("TT_PARSEQ_AR_LOGITS")
            res[mode] = (l.copy(), i.copy(), la.copy())
        monkeypatch.delenv("TT_DEC_EARLY_EXIT")
        (l1, i1, a1), (l0, i0, a0) = res["1"], res["0"]
        assert np.array_equal(i1, i0)
        assert np.array_equal(l1, l0), float(np.abs(l1 - l0).max())
        # AR pass: step i writes position i; the crop is live up to and including the step whose argmax is EOS
        ar_tok = a0[..., :95].argmax(-1)                      # [n][26] tokens of the full schedule
        first = np.where((ar_tok == eos).any(-1), (ar_tok == eos).argmax(-1), ar_tok.shape[1] - 1)
        live = np.arange(ar_tok.shape[1])[None, :] <= first[:, None]
        assert np.array_equal(a1[live], a0[live])
        steps = (first + 1).clip(max=26)
        print(n, "mean AR steps per crop", float(steps.mean()), "of 26; crops running all 26:", int((steps == 26).sum()))


def test_parseq_encoder_layernorm_fusion_matches_unfused(engine, monkeypatch):
    """Encoder with LayerNorm folded into the GEMM epilogues (default) against the standalone LayerNorm kernel
    (TT_ENC_LNFUSE=0), same forced AR context: logits agree to bf16 noise, clear decisions are identical."""
    crops = _crops(200, seed=7)
    _, id0 = engine.parseq_forward(crops)
    forced = np.ascontiguousarray(id0[:, :25]).astype(np.int32)
    lf, idf = engine.parseq_forward(crops, forced)
    monkeypatch.setenv("TT_ENC_LNFUSE", "0")
    lu, idu = engine.parseq_forward(crops, forced)
    monkeypatch.delenv("TT_ENC_LNFUSE")
    err = _rel_l2(lf, lu)
    top2 = np.sort(lu, -1)[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 0.5
    print("LN-fused vs unfused encoder: logits rel-L2", err, "ids equal", float((idf == idu).mean()))
    assert 0 < err <= 1e-2, err
    assert (idf[clear] == idu[clear]).all()


def test_parseq_encoder_mlp_fusion_matches_unfused(engine, monkeypatch):
    """Encoder with the MLP block as one kernel (enc_mlp.cu, default) against the fc1 / fc2 GEMM launches
    (TT_ENC_MLPFUSE=0), same forced AR context: logits agree to bf16 noise, clear decisions are identical.  Sizes: an odd
    number of 128-row tiles per pair schedule (77 crops) and more tiles than CTA pairs (300 crops)."""
    for n, seed in ((77, 21), (300, 22)):
        crops = _crops(n + (n & 1), seed=seed)[:n]
        _, id0 = engine.parseq_forward(crops)
        forced = np.ascontiguousarray(id0[:, :25]).astype(np.int32)
        lf, idf = engine.parseq_forward(crops, forced)          # default: proj + MLP in one kernel
        monkeypatch.setenv("TT_ENC_PROJFUSE", "0")
        lm, idm = engine.parseq_forward(crops, forced)          # proj as a GEMM launch, the MLP fused
        monkeypatch.delenv("TT_ENC_PROJFUSE")
        monkeypatch.setenv("TT_ENC_MLPFUSE", "0")
        lu, idu = engine.parseq_forward(crops, forced)          # proj, fc1, fc2 as GEMM launches
        monkeypatch.delenv("TT_ENC_MLPFUSE")
        top2 = np.sort(lu, -1)[..., -2:]
        clear = (top2[..., 1] - top2[..., 0]) > 0.5
        for name, l, i in (("proj+MLP kernel", lf, idf), ("MLP kernel", lm, idm)):
            err = _rel_l2(l, lu)
            print(n, name, "vs GEMM launches: logits rel-L2", err, "ids equal", float((i == idu).mean()))
            assert np.isfinite(l).all()
            assert 0 < err <= 1e-2, err
            assert (i[clear] == idu[clear]).all()


def test_parseq_tiny_variant(oracle_models):
    """PARSeq-tiny (embed 192, 3 encoder / 6 decoder heads, MLP 768: the other variant the HuggingFace checkpoint may be,
    /root/reference/.gitignore:1): dims come from the weight file's meta tensor; logits against the fp32 oracle under
    the oracle's own AR context, ids wherever its decision is clear, both decoder paths."""
    import os

    from conftest import ROOT
    from oracle.models import make_parseq
    from tuatara_b200 import weights

    craft, _ = oracle_models
    tiny = make_parseq("tiny", 0)
    d = ROOT / "tests" / "_cache" / "weights_tiny_seed0"
    d.mkdir(parents=True, exist_ok=True)
    if not (d / "craft.ttw").exists():
        weights.export_craft(craft.state_dict(), d / "craft.ttw")
    if not (d / "parseq.ttw").exists():
        weights.export_parseq(tiny.state_dict(), d / "parseq.ttw")
    eng = tb.Engine(str(d), devices=[0])
    try:
        crops = _crops(40, seed=9)
        x = torch.from_numpy(crops).permute(0, 3, 1, 2).float().div(255.0)
        taps = {}
        tiny(x, taps=taps)
        forced = taps["ar_tokens"][:, 1:].clone()
        ref = tiny(x, forced_tokens=forced).numpy()
        for fused in ("1", "0"):
            os.environ["TT_DEC_FUSED"] = fused
            try:
                got, ids = eng.parseq_forward(crops, forced.numpy().astype(np.int32))
            finally:
                os.environ.pop("TT_DEC_FUSED")
            err = _rel_l2(got, ref)
            print("tiny logits rel-L2", err, "(fused decoder" if fused == "1" else "(unfused decoder", ")")
            assert np.isfinite(got).all() and err <= 3e-2, err
            top2 = np.sort(ref, -1)[..., -2:]
            clear = (top2[..., 1] - top2[..., 0]) > 0.5
            assert (ids[clear] == ref.argmax(-1)[clear]).all()
        # whole path with the tiny recogniser: same boxes as with any recogniser, strings decode
        out = eng.ocr_pages([synth.synth_page(0)], score_override=[synth.synth_score_maps(0)])[0]
        assert len(out) == 300 and all(isinstance(o["text"], str) for o in out)
    finally:
        eng.close()


def test_converted_torchscript_weights_run_identically(engine, oracle_models, tmp_path):
    """the reference's two weight files (TorchScript archives, tuatara.cpp:333 / :423) through
    tuatara_b200.convert, loaded by a second engine: logits and ids equal those of the directly exported weights."""
    from tuatara_b200 import convert

    craft, parseq = oracle_models
    torch.jit.trace(craft, torch.zeros(1, 3, 64, 64), check_trace=False).save(str(tmp_path / convert.CRAFT_FILE))
    torch.jit.trace(parseq, torch.zeros(1, 3, 32, 128), check_trace=False).save(str(tmp_path / convert.PARSEQ_FILE))
    assert convert.main(["--weights-dir", str(tmp_path)]) == 0
    eng2 = tb.Engine(str(tmp_path), devices=[0])
    try:
        crops = _crops(24, seed=11)
        l1, i1 = engine.parseq_forward(crops)
        l2, i2 = eng2.parseq_forward(crops)
        assert np.array_equal(i1, i2) and np.array_equal(l1, l2)
        craft_in, _ = tb.preprocess(synth.synth_page(5)[:320, :384])
        assert np.array_equal(engine.craft_forward(craft_in), eng2.craft_forward(craft_in))
    finally:
        eng2.close()


def test_decode_matches_reference_tokenizer(native_lib):
    rng = np.random.default_rng(0)
    tok = R.Tokenizer()
    logits = torch.from_numpy(rng.standard_normal((200, 26, 95)).astype(np.float32))
    ref = [R.truncate_at_eos(t) for t in tok.decode(torch.softmax(logits, -1))]
    ids = logits.argmax(-1).numpy().astype(np.int32)
    assert tb.decode_ids(ids) == ref


def test_parseq_full_batch_is_batch_invariant(engine):
    """configs[2] size (1024 synthetic 32x128 crops): a size-independent property instead of the oracle, which needs
    minutes for this batch -- every crop's logits / ids must not depend on what else is in the batch (tile schedule,
    CTA pairs, weight-resident vs streaming, TMA vs register epilogues all change with M; the arithmetic must not)."""
    crops = np.random.default_rng(0).integers(0, 256, (1024, 32, 128, 3), dtype=np.uint8)
    crops[:512] = _crops(342, seed=5)[:512]  # half page-like crops, half noise
    logits, ids = engine.parseq_forward(crops)
    assert np.isfinite(logits).all()
    for s in range(0, 1024, 192):  # ragged sub-batches: 192, ..., 64
        l2, i2 = engine.parseq_forward(crops[s:s + 192])
        assert np.array_equal(i2, ids[s:s + 192])
        assert float(np.abs(l2 - logits[s:s + 192]).max()) <= 1e-4
