"""fp32 torch-CPU restatements of the two networks the reference loads as opaque TorchScript.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference never defines these graphs: it does ``torch::jit::load`` of
``craft_traced_torchscript_model.pt`` (tuatara.cpp:333-336, forward at :376) and
``parseq_torchscript.bin`` (tuatara.cpp:423-428, forward at :307).  The files come from
HuggingFace (setup.sh:6) and are unavailable offline, so the architectures are restated
from the public definitions the reference names (clovaai/CRAFT-pytorch ``craft.py`` +
``basenet/vgg16_bn.py``; baudm/parseq ``strhub/models/parseq`` + timm ViT), with parameter
names kept identical to upstream so ``tuatara_b200.weights`` can export either these
random-init models or a real checkpoint's ``state_dict``.

Evidence inside the reference that these are the graphs: CRAFT returns a tuple whose
element 0 is (1, H/2, W/2, 2) channels-last (tuatara.cpp:377-394), input padded to x32
(:225-226); PARSeq takes (N,3,32,128) (:440) and yields (N, L, C) consumed by softmax(-1)
(:486) and a 94-symbol tokenizer (:32-39).

NOTE on CRAFT skip tensors: torchvision's vgg16_bn uses ``nn.ReLU(inplace=True)`` and
upstream does ``h_relu2_2 = h; h = self.slice2(h)`` where slice2 *starts* with that
in-place ReLU, so the skip tensors of slices 1-3 are ReLU'd through aliasing (a traced
TorchScript keeps the ``relu_``).  Slice 5 starts with a MaxPool, so ``relu5_3`` really is
the pre-ReLU BatchNorm output.  The module structure below reproduces this naturally.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# CRAFT  (clovaai/CRAFT-pytorch craft.py / basenet/vgg16_bn.py)
# --------------------------------------------------------------------------------------

_VGG16_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]


def _vgg16_bn_features() -> nn.Sequential:
    """torchvision.models.vgg16_bn().features, restated (indices 0..43)."""
    layers, c_in = [], 3
    for v in _VGG16_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(c_in, v, kernel_size=3, padding=1), nn.BatchNorm2d(v), nn.ReLU(inplace=True)]
            c_in = v
    return nn.Sequential(*layers)


class Vgg16BN(nn.Module):
    def __init__(self):
        super().__init__()
        feats = _vgg16_bn_features()
        self.slice1 = nn.Sequential()
        self.slice2 = nn.Sequential()
        self.slice3 = nn.Sequential()
        self.slice4 = nn.Sequential()
        self.slice5 = nn.Sequential()
        for x in range(12):  # conv2_2 + BN
            self.slice1.add_module(str(x), feats[x])
        for x in range(12, 19):  # conv3_2 + BN
            self.slice2.add_module(str(x), feats[x])
        for x in range(19, 29):  # conv4_2 + BN  (upstream calls it relu4_3)
            self.slice3.add_module(str(x), feats[x])
        for x in range(29, 39):  # conv5_2 + BN  (upstream calls it relu5_3)
            self.slice4.add_module(str(x), feats[x])
        self.slice5 = nn.Sequential(
            nn.MaxPool2d(kernel_size=3, stride=1, padding=1),
            nn.Conv2d(512, 1024, kernel_size=3, padding=6, dilation=6),
            nn.Conv2d(1024, 1024, kernel_size=1),
        )

    def forward(self, x):
        h = self.slice1(x)
        h_relu2_2 = h
        h = self.slice2(h)  # first op is an in-place ReLU: h_relu2_2 is ReLU'd too
        h_relu3_2 = h
        h = self.slice3(h)
        h_relu4_3 = h
        h = self.slice4(h)
        h_relu5_3 = h
        h = self.slice5(h)  # starts with MaxPool (out of place): h_relu5_3 stays pre-ReLU
        h_fc7 = h
        return h_fc7, h_relu5_3, h_relu4_3, h_relu3_2, h_relu2_2


class DoubleConv(nn.Module):
    def __init__(self, in_ch, mid_ch, out_ch):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(in_ch + mid_ch, mid_ch, kernel_size=1),
            nn.BatchNorm2d(mid_ch),
            nn.ReLU(inplace=True),
            nn.Conv2d(mid_ch, out_ch, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_ch),
            nn.ReLU(inplace=True),
        )

    def forward(self, x):
        return self.conv(x)


class CRAFT(nn.Module):
    """VGG16-BN backbone + U-net head -> ((N,H/2,W/2,2) region/affinity, feature)."""

    def __init__(self):
        super().__init__()
        self.basenet = Vgg16BN()
        self.upconv1 = DoubleConv(1024, 512, 256)
        self.upconv2 = DoubleConv(512, 256, 128)
        self.upconv3 = DoubleConv(256, 128, 64)
        self.upconv4 = DoubleConv(128, 64, 32)
        self.conv_cls = nn.Sequential(
            nn.Conv2d(32, 32, kernel_size=3, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(32, 32, kernel_size=3, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(32, 16, kernel_size=3, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(16, 16, kernel_size=1), nn.ReLU(inplace=True),
            nn.Conv2d(16, 2, kernel_size=1),
        )

    def forward(self, x, taps: dict | None = None):
        sources = self.basenet(x)
        if taps is not None:
            for name, t in zip(("fc7", "relu5_3", "relu4_3", "relu3_2", "relu2_2"), sources):
                taps[name] = t
        y = torch.cat([sources[0], sources[1]], dim=1)
        y = self.upconv1(y)
        if taps is not None:
            taps["up1"] = y
        y = F.interpolate(y, size=sources[2].size()[2:], mode="bilinear", align_corners=False)
        y = torch.cat([y, sources[2]], dim=1)
        y = self.upconv2(y)
        if taps is not None:
            taps["up2"] = y
        y = F.interpolate(y, size=sources[3].size()[2:], mode="bilinear", align_corners=False)
        y = torch.cat([y, sources[3]], dim=1)
        y = self.upconv3(y)
        if taps is not None:
            taps["up3"] = y
        y = F.interpolate(y, size=sources[4].size()[2:], mode="bilinear", align_corners=False)
        y = torch.cat([y, sources[4]], dim=1)
        feature = self.upconv4(y)
        if taps is not None:
            taps["up4"] = feature
        y = self.conv_cls(feature)
        return y.permute(0, 2, 3, 1), feature


def make_craft(seed: int = 0) -> CRAFT:
    """Seeded random init that keeps activations O(1) through 27 convs (He init, BN stats
    randomised so folding is exercised).  Not upstream's init: checkpoints are unavailable
    offline and upstream's xavier init + fresh BN gives near-constant maps."""
    g = torch.Generator().manual_seed(seed)
    m = CRAFT()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, nn.Conv2d):
                fan_in = mod.in_channels * mod.kernel_size[0] * mod.kernel_size[1]
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * math.sqrt(2.0 / fan_in))
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.05)
            elif isinstance(mod, nn.BatchNorm2d):
                mod.weight.copy_(1.0 + 0.2 * (torch.rand(mod.weight.shape, generator=g) - 0.5))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
                mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=g))
                mod.running_var.copy_(0.75 + 0.5 * torch.rand(mod.running_var.shape, generator=g))
    return m.eval()


# --------------------------------------------------------------------------------------
# PARSeq  (baudm/parseq strhub/models/parseq/{system,modules}.py + timm VisionTransformer)
# --------------------------------------------------------------------------------------


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)  # (N, 128, D), row-major over (py, px)


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = (q * self.scale) @ k.transpose(-2, -1)
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class Encoder(nn.Module):
    """timm VisionTransformer(num_classes=0, global_pool='', class_token=False)."""

    def __init__(self, img_size, patch_size, embed_dim, depth, num_heads, mlp_ratio):
        super().__init__()
        self.patch_embed = PatchEmbed(img_size, patch_size, 3, embed_dim)
        n_tok = (img_size[0] // patch_size[0]) * (img_size[1] // patch_size[1])
        self.pos_embed = nn.Parameter(torch.zeros(1, n_tok, embed_dim))
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)

    def forward(self, x, taps=None):
        x = self.patch_embed(x) + self.pos_embed
        if taps is not None:
            taps["embed"] = x
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if taps is not None:
                taps[f"block{i}"] = x
        return self.norm(x)


class DecoderLayer(nn.Module):
    """Two-stream layer; with depth 1 only the query stream is evaluated (update_content=False)."""

    def __init__(self, d_model, nhead, dim_feedforward, layer_norm_eps=1e-5):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=0.0, batch_first=True)
        self.cross_attn = nn.MultiheadAttention(d_model, nhead, dropout=0.0, batch_first=True)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm2 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm_q = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm_c = nn.LayerNorm(d_model, eps=layer_norm_eps)

    def forward_stream(self, tgt, tgt_norm, tgt_kv, memory, tgt_mask, tgt_key_padding_mask):
        tgt2, _ = self.self_attn(tgt_norm, tgt_kv, tgt_kv, attn_mask=tgt_mask,
                                 key_padding_mask=tgt_key_padding_mask, need_weights=False)
        tgt = tgt + tgt2
        tgt2, _ = self.cross_attn(self.norm1(tgt), memory, memory, need_weights=False)
        tgt = tgt + tgt2
        tgt2 = self.linear2(F.gelu(self.linear1(self.norm2(tgt))))
        return tgt + tgt2

    def forward(self, query, content, memory, query_mask=None, content_key_padding_mask=None):
        query_norm = self.norm_q(query)
        content_norm = self.norm_c(content)
        return self.forward_stream(query, query_norm, content_norm, memory, query_mask, content_key_padding_mask)


class Decoder(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward):
        super().__init__()
        self.layers = nn.ModuleList([DecoderLayer(d_model, nhead, dim_feedforward)])
        self.norm = nn.LayerNorm(d_model)

    def forward(self, query, content, memory, query_mask=None, content_key_padding_mask=None):
        query = self.layers[0](query, content, memory, query_mask, content_key_padding_mask)
        return self.norm(query)


class TokenEmbedding(nn.Module):
    def __init__(self, charset_size, embed_dim):
        super().__init__()
        self.embedding = nn.Embedding(charset_size, embed_dim)
        self.embed_dim = embed_dim

    def forward(self, tokens):
        return math.sqrt(self.embed_dim) * self.embedding(tokens)


PARSEQ_DIMS = {
    "base": dict(embed_dim=384, enc_num_heads=6, dec_num_heads=12),
    "tiny": dict(embed_dim=192, enc_num_heads=3, dec_num_heads=6),
}


class PARSeq(nn.Module):
    """Inference graph of upstream PARSeq with decode_ar=True, refine_iters=1.

    Upstream ids: eos 0, chars 1..94, bos 95, pad 96 (the *model's* ids; the reference's C++
    tokenizer re-interprets the 95 output classes differently, see oracle/tokenizer.py)."""

    def __init__(self, embed_dim=384, enc_num_heads=6, dec_num_heads=12, enc_depth=12,
                 mlp_ratio=4, max_label_length=25, num_tokens=97):
        super().__init__()
        self.max_label_length = max_label_length
        self.eos_id, self.bos_id, self.pad_id = 0, num_tokens - 2, num_tokens - 1
        self.encoder = Encoder((32, 128), (4, 8), embed_dim, enc_depth, enc_num_heads, mlp_ratio)
        self.decoder = Decoder(embed_dim, dec_num_heads, embed_dim * mlp_ratio)
        self.head = nn.Linear(embed_dim, num_tokens - 2)
        self.text_embed = TokenEmbedding(num_tokens, embed_dim)
        self.pos_queries = nn.Parameter(torch.zeros(1, max_label_length + 1, embed_dim))

    def encode(self, img, taps=None):
        return self.encoder(img, taps)

    def decode(self, tgt, memory, tgt_padding_mask=None, tgt_query=None, tgt_query_mask=None):
        N, L = tgt.shape
        null_ctx = self.text_embed(tgt[:, :1])
        tgt_emb = self.pos_queries[:, : L - 1] + self.text_embed(tgt[:, 1:])
        tgt_emb = torch.cat([null_ctx, tgt_emb], dim=1)
        if tgt_query is None:
            tgt_query = self.pos_queries[:, :L].expand(N, -1, -1)
        return self.decoder(tgt_query, tgt_emb, memory, tgt_query_mask, tgt_padding_mask)

    @torch.no_grad()
    def forward(self, images, forced_tokens: torch.Tensor | None = None, taps: dict | None = None):
        """Fixed 26-step schedule by default (upstream's early exit only shortens the tensor;
        positions up to the first EOS are unaffected -- SURVEY App. B).  With ``self.early_exit``
        set, upstream's ``if testing and (tgt_in == self.eos_id).any(dim=-1).all(): break`` is
        applied as upstream does, per forward() batch -- the reference calls forward() on 4 crops
        at a time (tuatara.cpp:452-475); bench.py's CPU baseline / reference arm time it that way.
        The refinement then sees a shorter key sequence; every dropped key is behind an EOS and
        would be masked, so the returned (N, 26, C) logits are the same.
        ``forced_tokens`` (N,25) int64, when given, replaces the argmax feedback so a bf16
        implementation can be compared position-by-position without chaotic divergence (teacher
        forcing; test-only; no early exit)."""
        bs = images.shape[0]
        num_steps = self.max_label_length + 1
        memory = self.encode(images, taps)
        if taps is not None:
            taps["memory"] = memory
        pos_queries = self.pos_queries[:, :num_steps].expand(bs, -1, -1)
        query_mask = torch.triu(torch.full((num_steps, num_steps), float("-inf")), 1)
        tgt_in = torch.full((bs, num_steps), self.pad_id, dtype=torch.long)
        tgt_in[:, 0] = self.bos_id
        logits = []
        for i in range(num_steps):
            j = i + 1
            tgt_out = self.decode(tgt_in[:, :j], memory, tgt_query=pos_queries[:, i:j],
                                  tgt_query_mask=query_mask[i:j, :j])
            p_i = self.head(tgt_out)
            logits.append(p_i)
            if j < num_steps:
                tgt_in[:, j] = p_i[:, 0].argmax(-1) if forced_tokens is None else forced_tokens[:, i]
                if getattr(self, "early_exit", False) and forced_tokens is None and taps is None \
                        and bool((tgt_in == self.eos_id).any(dim=-1).all()):
                    break
        logits = torch.cat(logits, dim=1)
        if taps is not None:
            taps["ar_logits"] = logits
            taps["ar_tokens"] = tgt_in.clone()
        # one refinement iteration with the cloze mask
        query_mask[torch.triu(torch.ones(num_steps, num_steps, dtype=torch.bool), 2)] = 0
        bos = torch.full((bs, 1), self.bos_id, dtype=torch.long)
        tgt_in = torch.cat([bos, logits[:, :-1].argmax(-1)], dim=1)
        if forced_tokens is not None:
            tgt_in = torch.cat([bos, forced_tokens], dim=1)
        tgt_padding_mask = (tgt_in == self.eos_id).int().cumsum(-1) > 0
        tgt_out = self.decode(tgt_in, memory, tgt_padding_mask, tgt_query=pos_queries,
                              tgt_query_mask=query_mask[:, : tgt_in.shape[1]])
        return self.head(tgt_out)


def make_parseq(variant: str = "base", seed: int = 0) -> PARSeq:
    """Seeded random init scaled so attention is not uniform and logits have usable margins."""
    g = torch.Generator().manual_seed(seed + 1000)
    m = PARSeq(**PARSEQ_DIMS[variant])

    def rn(shape, std):
        return torch.randn(shape, generator=g) * std

    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith("pos_embed"):
                p.copy_(rn(p.shape, 0.2))
            elif name == "pos_queries":
                p.copy_(rn(p.shape, 0.5))
            elif "text_embed" in name:
                p.copy_(rn(p.shape, 0.05))
            elif name.startswith("head.weight"):
                p.copy_(rn(p.shape, 0.2))
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.2 * (torch.rand(p.shape, generator=g) - 0.5))
            elif "norm" in name and name.endswith("bias"):
                p.copy_(rn(p.shape, 0.05))
            elif name.endswith("patch_embed.proj.weight"):
                p.copy_(rn(p.shape, 0.1))
            elif name.endswith("bias"):
                p.copy_(rn(p.shape, 0.02))
            else:  # every Linear / in_proj weight
                p.copy_(rn(p.shape, 0.05))
    return m.eval()
