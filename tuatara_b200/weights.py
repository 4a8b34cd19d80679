"""Exporter for the engine's flat weight files (``craft.ttw`` / ``parseq.ttw``).

The reference loads two TorchScript blobs on every call (tuatara.cpp:333-336, :423-428);
TorchScript cannot be parsed without LibTorch, so the engine reads its own flat format,
produced here from a ``state_dict`` with the *upstream* parameter names
(clovaai/CRAFT-pytorch, baudm/parseq) -- a real checkpoint's state_dict exports the same way
as the seeded random-init models the tests use.

Layouts are the ones the CUDA kernels consume directly:
  * conv weights  bf16 [Cout][tap][Cin]   (K-major rows for the implicit GEMM), BatchNorm
    folded in fp32 before the cast, bias fp32;
  * conv1_1       bf16 [64][32]: k = tap*3 + c (27 used), scaled by 1/255 because the kernel
    feeds raw 0..255 pixels (the /255 of tuatara.cpp:370 lives in the weights);
  * linear weights bf16 [N][K]; patch-embed [384][96] scaled by 1/255 (tuatara.cpp:446);
    head padded to 96 rows; embedding pre-multiplied by sqrt(embed_dim).

File format (little endian): b"TTW1", u32 n, then n x {char name[64]; u32 dtype (0 f32, 1 bf16,
2 i32); u32 ndim; u64 dims[4]; u64 offset; u64 nbytes}, then payloads at 256-byte aligned offsets.
"""
from __future__ import annotations

import math
import struct
from pathlib import Path

import numpy as np
import torch

_DT = {torch.float32: 0, torch.bfloat16: 1, torch.int32: 2}


def write_ttw(path: str | Path, tensors: dict[str, torch.Tensor]) -> None:
    names = list(tensors)
    header = 8 + len(names) * (64 + 4 + 4 + 32 + 8 + 8)
    off = (header + 255) // 256 * 256
    entries, blobs = [], []
    for n in names:
        t = tensors[n].detach().contiguous().cpu()
        raw = t.view(torch.uint8).numpy().tobytes() if t.dtype != torch.bfloat16 else t.view(torch.int16).numpy().tobytes()
        dims = list(t.shape) + [1] * (4 - t.dim())
        entries.append(struct.pack("<64sII4QQQ", n.encode(), _DT[t.dtype], t.dim(), *dims, off, len(raw)))
        blobs.append((off, raw))
        off = (off + len(raw) + 255) // 256 * 256
    with open(path, "wb") as f:
        f.write(b"TTW1" + struct.pack("<I", len(names)))
        for e in entries:
            f.write(e)
        for o, raw in blobs:
            f.seek(o)
            f.write(raw)
        f.truncate(off)


def read_ttw(path: str | Path) -> dict[str, torch.Tensor]:
    buf = Path(path).read_bytes()
    assert buf[:4] == b"TTW1"
    (n,) = struct.unpack_from("<I", buf, 4)
    out = {}
    for i in range(n):
        name, dt, nd, d0, d1, d2, d3, off, nb = struct.unpack_from("<64sII4QQQ", buf, 8 + i * 120)
        shape = [d0, d1, d2, d3][:nd]
        raw = np.frombuffer(buf, np.uint8, nb, off).copy()
        if dt == 0:
            t = torch.from_numpy(raw.view(np.float32))
        elif dt == 1:
            t = torch.from_numpy(raw.view(np.int16)).view(torch.bfloat16)
        else:
            t = torch.from_numpy(raw.view(np.int32))
        out[name.rstrip(b"\0").decode()] = t.reshape(shape)
    return out


# ----------------------------------------------------------------------------------- CRAFT
def _fold_bn(w, b, sd, bn_prefix):
    """conv (w, b) followed by eval-mode BatchNorm(bn_prefix) -> equivalent (w', b') in fp32."""
    g, beta = sd[bn_prefix + ".weight"].double(), sd[bn_prefix + ".bias"].double()
    mean, var = sd[bn_prefix + ".running_mean"].double(), sd[bn_prefix + ".running_var"].double()
    s = g / torch.sqrt(var + 1e-5)
    return (w.double() * s[:, None, None, None]).float(), ((b.double() - mean) * s + beta).float()


def _khwc(w):  # [Cout][Cin][kh][kw] -> [Cout][kh*kw][Cin]
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci).contiguous()


# (engine layer name, conv key, bn key or None) in execution order
CRAFT_LAYERS = [
    ("c1_1", "basenet.slice1.0", "basenet.slice1.1"), ("c1_2", "basenet.slice1.3", "basenet.slice1.4"),
    ("c2_1", "basenet.slice1.7", "basenet.slice1.8"), ("c2_2", "basenet.slice1.10", "basenet.slice1.11"),
    ("c3_1", "basenet.slice2.14", "basenet.slice2.15"), ("c3_2", "basenet.slice2.17", "basenet.slice2.18"),
    ("c3_3", "basenet.slice3.20", "basenet.slice3.21"), ("c4_1", "basenet.slice3.24", "basenet.slice3.25"),
    ("c4_2", "basenet.slice3.27", "basenet.slice3.28"), ("c4_3", "basenet.slice4.30", "basenet.slice4.31"),
    ("c5_1", "basenet.slice4.34", "basenet.slice4.35"), ("c5_2", "basenet.slice4.37", "basenet.slice4.38"),
    ("fc6", "basenet.slice5.1", None), ("fc7", "basenet.slice5.2", None),
    ("up1a", "upconv1.conv.0", "upconv1.conv.1"), ("up1b", "upconv1.conv.3", "upconv1.conv.4"),
    ("up2a", "upconv2.conv.0", "upconv2.conv.1"), ("up2b", "upconv2.conv.3", "upconv2.conv.4"),
    ("up3a", "upconv3.conv.0", "upconv3.conv.1"), ("up3b", "upconv3.conv.3", "upconv3.conv.4"),
    ("up4a", "upconv4.conv.0", "upconv4.conv.1"), ("up4b", "upconv4.conv.3", "upconv4.conv.4"),
    ("cls1", "conv_cls.0", None), ("cls2", "conv_cls.2", None), ("cls3", "conv_cls.4", None),
]


def export_craft(state_dict: dict, path: str | Path) -> None:
    sd = {k: v.detach().float() for k, v in state_dict.items()}
    out = {}
    for name, ck, bk in CRAFT_LAYERS:
        w, b = sd[ck + ".weight"], sd[ck + ".bias"]
        if bk is not None:
            w, b = _fold_bn(w, b, sd, bk)
        w = _khwc(w)
        if name == "c1_1":  # raw 0..255 pixels in, K padded 27 -> 32
            w = (w.double() / 255.0).float().reshape(w.shape[0], 27)
            w = torch.cat([w, torch.zeros(w.shape[0], 5)], 1)
        out[name + ".w"] = w.reshape(w.shape[0], -1).to(torch.bfloat16)
        out[name + ".b"] = b.float()
    # tail of conv_cls: 1x1 16->16 (+ReLU) and 1x1 16->2 run in fp32 inside cls3's epilogue
    w4, b4 = sd["conv_cls.6.weight"].reshape(16, 16), sd["conv_cls.6.bias"]
    w5, b5 = sd["conv_cls.8.weight"].reshape(2, 16), sd["conv_cls.8.bias"]
    out["cls.tail"] = torch.cat([w4.reshape(-1), b4, w5.reshape(-1), b5]).float()
    write_ttw(path, out)


# ---------------------------------------------------------------------------------- PARSeq
def export_parseq(state_dict: dict, path: str | Path) -> None:
    sd = {k: v.detach().float() for k, v in state_dict.items()}
    d = sd["pos_queries"].shape[-1]
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.blocks."))
    n_cls = sd["head.weight"].shape[0]
    out = {}
    bf = lambda t: t.contiguous().to(torch.bfloat16)  # noqa: E731
    pe = sd["encoder.patch_embed.proj.weight"]  # [D][3][4][8] -> [D][96], k = c*32 + ky*8 + kx
    out["pe.w"] = bf((pe.double() / 255.0).float().reshape(d, -1))
    out["pe.b"] = sd["encoder.patch_embed.proj.bias"]
    out["pos"] = sd["encoder.pos_embed"].reshape(-1, d)
    for i in range(depth):
        p = f"encoder.blocks.{i}."
        out[f"b{i}.ln1.g"], out[f"b{i}.ln1.b"] = sd[p + "norm1.weight"], sd[p + "norm1.bias"]
        out[f"b{i}.ln2.g"], out[f"b{i}.ln2.b"] = sd[p + "norm2.weight"], sd[p + "norm2.bias"]
        out[f"b{i}.qkv.w"], out[f"b{i}.qkv.b"] = bf(sd[p + "attn.qkv.weight"]), sd[p + "attn.qkv.bias"]
        out[f"b{i}.proj.w"], out[f"b{i}.proj.b"] = bf(sd[p + "attn.proj.weight"]), sd[p + "attn.proj.bias"]
        out[f"b{i}.fc1.w"], out[f"b{i}.fc1.b"] = bf(sd[p + "mlp.fc1.weight"]), sd[p + "mlp.fc1.bias"]
        out[f"b{i}.fc2.w"], out[f"b{i}.fc2.b"] = bf(sd[p + "mlp.fc2.weight"]), sd[p + "mlp.fc2.bias"]
    out["enc.ln.g"], out["enc.ln.b"] = sd["encoder.norm.weight"], sd["encoder.norm.bias"]

    # LayerNorm folded into the Linear that follows it (gemm_tc.cuh, Epilogue::ln_*):
    #   Linear(LN(x))[n] = rstd * (x . W'[n]) - rstd * mean * c1[n] + c0[n],   W' = W * gamma (bf16),
    #   c1[n] = sum_k W'[n][k] (of the ROUNDED values, so the mean term cancels exactly), c0[n] = b[n] + beta . W[n]
    def fold_ln(name, w, b, g, beta):
        wf_ = (w * g[None, :]).to(torch.bfloat16)
        out[name + ".wf"] = wf_.contiguous()
        out[name + ".c1"] = wf_.double().sum(1).float()
        out[name + ".c0"] = (b.double() + w.double() @ beta.double()).float()

    for i in range(depth):
        p = f"encoder.blocks.{i}."
        fold_ln(f"b{i}.qkv", sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"], sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        fold_ln(f"b{i}.fc1", sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"], sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    ca_w, ca_b = sd["decoder.layers.0.cross_attn.in_proj_weight"], sd["decoder.layers.0.cross_attn.in_proj_bias"]
    fold_ln("dec.ca.kvf", ca_w[d:], ca_b[d:], sd["encoder.norm.weight"], sd["encoder.norm.bias"])
    L = "decoder.layers.0."
    for short, full in (("nq", "norm_q"), ("nc", "norm_c"), ("n1", "norm1"), ("n2", "norm2")):
        out[f"dec.{short}.g"], out[f"dec.{short}.b"] = sd[L + full + ".weight"], sd[L + full + ".bias"]
    out["dec.norm.g"], out["dec.norm.b"] = sd["decoder.norm.weight"], sd["decoder.norm.bias"]
    for short, full in (("sa", "self_attn"), ("ca", "cross_attn")):
        out[f"dec.{short}.in.w"], out[f"dec.{short}.in.b"] = bf(sd[L + full + ".in_proj_weight"]), sd[L + full + ".in_proj_bias"]
        out[f"dec.{short}.out.w"], out[f"dec.{short}.out.b"] = bf(sd[L + full + ".out_proj.weight"]), sd[L + full + ".out_proj.bias"]
    out["dec.l1.w"], out["dec.l1.b"] = bf(sd[L + "linear1.weight"]), sd[L + "linear1.bias"]
    out["dec.l2.w"], out["dec.l2.b"] = bf(sd[L + "linear2.weight"]), sd[L + "linear2.bias"]
    pad = (-n_cls) % 16
    out["head.w"] = bf(torch.cat([sd["head.weight"], torch.zeros(pad, d)], 0))
    out["head.b"] = torch.cat([sd["head.bias"], torch.zeros(pad)], 0)
    out["embed"] = (sd["text_embed.embedding.weight"] * math.sqrt(d)).contiguous()
    out["posq"] = sd["pos_queries"].reshape(-1, d).contiguous()
    enc_heads = d // 64
    dec_heads = d // 32
    out["meta"] = torch.tensor([d, depth, enc_heads, dec_heads, sd[L + "linear1.weight"].shape[0], n_cls,
                                sd["pos_queries"].shape[1], sd["text_embed.embedding.weight"].shape[0]],
                               dtype=torch.int32)
    write_ttw(path, out)


# ------------------------------------------------------------- seeded random-init checkpoints
def random_craft_state_dict(seed: int = 0) -> dict:
    """A CRAFT ``state_dict`` (upstream names/shapes) with He-initialised convs and randomised
    BatchNorm statistics -- stands in for the HuggingFace checkpoint (setup.sh:6) offline."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(key, cout, cin, k):
        sd[key + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))
        sd[key + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def bn(key, c):
        sd[key + ".weight"] = 1.0 + 0.2 * (torch.rand(c, generator=g) - 0.5)
        sd[key + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[key + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[key + ".running_var"] = 0.75 + 0.5 * torch.rand(c, generator=g)

    chans = {"c1_1": (3, 64), "c1_2": (64, 64), "c2_1": (64, 128), "c2_2": (128, 128), "c3_1": (128, 256),
             "c3_2": (256, 256), "c3_3": (256, 256), "c4_1": (256, 512), "c4_2": (512, 512), "c4_3": (512, 512),
             "c5_1": (512, 512), "c5_2": (512, 512), "fc6": (512, 1024), "fc7": (1024, 1024),
             "up1a": (1536, 512), "up1b": (512, 256), "up2a": (768, 256), "up2b": (256, 128),
             "up3a": (384, 128), "up3b": (128, 64), "up4a": (192, 64), "up4b": (64, 32),
             "cls1": (32, 32), "cls2": (32, 32), "cls3": (32, 16)}
    one_by_one = {"fc7", "up1a", "up2a", "up3a", "up4a"}
    for name, ck, bk in CRAFT_LAYERS:
        cin, cout = chans[name]
        conv(ck, cout, cin, 1 if name in one_by_one else 3)
        if bk is not None:
            bn(bk, cout)
    conv("conv_cls.6", 16, 16, 1)
    conv("conv_cls.8", 2, 16, 1)
    return sd


def random_parseq_state_dict(seed: int = 0, d: int = 384, depth: int = 12, mlp: int = 1536, n_tok: int = 97,
                             max_len: int = 25) -> dict:
    """A PARSeq ``state_dict`` (upstream names/shapes), scaled so attention is not uniform."""
    g = torch.Generator().manual_seed(seed + 1000)
    rn = lambda shape, std: torch.randn(shape, generator=g) * std  # noqa: E731
    sd = {}

    def linear(key, out, inp, std=0.05):
        sd[key + ".weight"], sd[key + ".bias"] = rn((out, inp), std), rn((out,), 0.02)

    def norm(key):
        sd[key + ".weight"], sd[key + ".bias"] = 1.0 + 0.2 * (torch.rand(d, generator=g) - 0.5), rn((d,), 0.05)

    sd["encoder.patch_embed.proj.weight"], sd["encoder.patch_embed.proj.bias"] = rn((d, 3, 4, 8), 0.1), rn((d,), 0.02)
    sd["encoder.pos_embed"] = rn((1, 128, d), 0.2)
    for i in range(depth):
        p = f"encoder.blocks.{i}."
        norm(p + "norm1"); norm(p + "norm2")
        linear(p + "attn.qkv", 3 * d, d); linear(p + "attn.proj", d, d)
        linear(p + "mlp.fc1", mlp, d); linear(p + "mlp.fc2", d, mlp)
    norm("encoder.norm")
    L = "decoder.layers.0."
    for a in ("self_attn", "cross_attn"):
        sd[L + a + ".in_proj_weight"], sd[L + a + ".in_proj_bias"] = rn((3 * d, d), 0.05), rn((3 * d,), 0.02)
        linear(L + a + ".out_proj", d, d)
    linear(L + "linear1", mlp, d); linear(L + "linear2", d, mlp)
    for k in ("norm1", "norm2", "norm_q", "norm_c"):
        norm(L + k)
    norm("decoder.norm")
    linear("head", n_tok - 2, d, 0.2)
    sd["text_embed.embedding.weight"] = rn((n_tok, d), 0.05)
    sd["pos_queries"] = rn((1, max_len + 1, d), 0.5)
    return sd


def export_random(out_dir: str | Path, seed: int = 0) -> str:
    """craft.ttw + parseq.ttw with seeded random-init weights of the named architectures."""
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    if not (out_dir / "craft.ttw").exists():
        export_craft(random_craft_state_dict(seed), out_dir / "craft.ttw")
    if not (out_dir / "parseq.ttw").exists():
        export_parseq(random_parseq_state_dict(seed), out_dir / "parseq.ttw")
    return str(out_dir)
