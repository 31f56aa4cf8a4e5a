"""Parameter inventory of the in-scope FeMaSRNet graph and seeded random weights for it.

The names/shapes are the reference's ``state_dict`` contract (built by
femasr_arch.py:216-309, fema_utils.py:65-99, network_swinir.py:65-145,164-214,419-482) so that a
checkpoint written by the reference loads here and vice versa.  ``tests/test_boundary.py`` checks
this list against the reference's own ``state_dict()`` when /root/reference is present.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, List, Tuple

import torch

CHANNELS = {8: 256, 16: 256, 32: 256, 64: 256, 128: 128, 256: 64, 512: 32}   # femasr_arch.py:244-252
GT_RES = 256
CB_SCALE = 32
WINDOW = 8
HEADS = 8
SWIN_DIM = 256
SWIN_DEPTH = 6
N_RSTB = 4
MLP_RATIO = 4
SWIN_INIT_RES = 32        # SwinLayers default input_resolution=(32,32), femasr_arch.py:115


def encode_depth(scale: int) -> int:
    """femasr_arch.py:256."""
    return int(math.log2(GT_RES // scale // CB_SCALE))


def _res_block(p: str, c: int) -> List[Tuple[str, tuple, str, int]]:
    return [
        (f"{p}.conv.0.norm.weight", (c,), "norm_w", 0), (f"{p}.conv.0.norm.bias", (c,), "norm_b", 0),
        (f"{p}.conv.2.weight", (c, c, 3, 3), "w", c * 9), (f"{p}.conv.2.bias", (c,), "b", c * 9),
        (f"{p}.conv.3.norm.weight", (c,), "norm_w", 0), (f"{p}.conv.3.norm.bias", (c,), "norm_b", 0),
        (f"{p}.conv.5.weight", (c, c, 3, 3), "w", c * 9), (f"{p}.conv.5.bias", (c,), "b", c * 9),
    ]


def _conv(p: str, ci: int, co: int, k: int) -> List[Tuple[str, tuple, str, int]]:
    return [(f"{p}.weight", (co, ci, k, k), "w", ci * k * k), (f"{p}.bias", (co,), "b", ci * k * k)]


def _linear(p: str, ci: int, co: int) -> List[Tuple[str, tuple, str, int]]:
    return [(f"{p}.weight", (co, ci), "w", ci), (f"{p}.bias", (co,), "b", ci)]


def normalize_codebooks(codebooks, n_e: int = 1024, e_dim: int = 256):
    """[(scale, n_e, e_dim), ...] as ints (the reference's ``codebook_params`` rows, femasr_arch.py:231-235).
    The first codebook must sit at scale 32 (it fixes the encoder/decoder depth, :255-256); further ones at
    strictly increasing decoder resolutions 64 / 128 (:329-331)."""
    if codebooks is None:
        return [(CB_SCALE, int(n_e), int(e_dim))]
    cbs = [(int(s), int(n), int(e)) for s, n, e in codebooks]
    if not cbs or cbs[0][0] != CB_SCALE:
        raise NotImplementedError("the first codebook must be at scale 32")
    scales = [s for s, _, _ in cbs]
    if any(s not in (32, 64, 128) for s in scales) or scales != sorted(set(scales)):
        raise NotImplementedError(f"codebook scales must be an increasing subset of 32, 64, 128, got {scales}")
    return cbs


def param_spec(scale: int, e_dim: int, n_e: int = 1024, in_channel: int = 3, codebooks=None):
    """Ordered [(name, shape, kind, fan_in)].  scale 4 | 2: LQ_stage=True;
    scale 1: the HQ autoencoder (LQ_stage=False, femasr_arch.py:241: scale_factor forced to 1; no Swin, no up branches).
    ``codebooks`` = [(scale, n_e, e_dim), ...] for the multi-scale variant (femasr_arch.py:280-299); default: one
    codebook (32, n_e, e_dim).

    kind: w | b (kaiming-uniform bound 1/sqrt(fan_in)), norm_w | norm_b, rpb (trunc-normal .02),
    rpi | mask (buffers), codebook (U(+-1/n_e)).
    """
    d = encode_depth(scale)
    res = GT_RES // scale
    spec: List[Tuple[str, tuple, str, int]] = []
    enc = "multiscale_encoder"
    spec += _conv(f"{enc}.in_conv", in_channel, CHANNELS[res], 4)
    for i in range(d):
        ci, co = CHANNELS[res], CHANNELS[res // 2]
        spec += _conv(f"{enc}.blocks.{i}.0", ci, co, 3)
        spec += _res_block(f"{enc}.blocks.{i}.1", co) + _res_block(f"{enc}.blocks.{i}.2", co)
        res //= 2
    C = SWIN_DIM
    hq = scale == 1
    for r in range(0 if hq else N_RSTB):
        for b in range(SWIN_DEPTH):
            p = f"{enc}.blocks.{d}.swin_blks.{r}.residual_group.blocks.{b}"
            if b % 2 == 1:
                nw = (SWIN_INIT_RES // WINDOW) ** 2
                spec.append((f"{p}.attn_mask", (nw, WINDOW ** 2, WINDOW ** 2), "mask", 0))
            spec += [(f"{p}.norm1.weight", (C,), "norm_w", 0), (f"{p}.norm1.bias", (C,), "norm_b", 0),
                     (f"{p}.attn.relative_position_bias_table", ((2 * WINDOW - 1) ** 2, HEADS), "rpb", 0),
                     (f"{p}.attn.relative_position_index", (WINDOW ** 2, WINDOW ** 2), "rpi", 0)]
            spec += _linear(f"{p}.attn.qkv", C, 3 * C) + _linear(f"{p}.attn.proj", C, C)
            spec += [(f"{p}.norm2.weight", (C,), "norm_w", 0), (f"{p}.norm2.bias", (C,), "norm_b", 0)]
            spec += _linear(f"{p}.mlp.fc1", C, MLP_RATIO * C) + _linear(f"{p}.mlp.fc2", MLP_RATIO * C, C)
        spec += _conv(f"{enc}.blocks.{d}.swin_blks.{r}.conv", C, C, 3)
    for j in (() if hq else (d + 1, d + 2)):
        ci, co = CHANNELS[res], CHANNELS[res * 2]
        spec += _conv(f"{enc}.blocks.{j}.1", ci, co, 3)
        spec += _res_block(f"{enc}.blocks.{j}.2", co) + _res_block(f"{enc}.blocks.{j}.3", co)
        res *= 2
    for i in range(3):
        r = GT_RES // 8 * 2 ** i
        ci, co = CHANNELS[r], CHANNELS[r * 2]
        spec += _conv(f"decoder_group.{i}.block.1", ci, co, 3)
        spec += _res_block(f"decoder_group.{i}.block.2", co) + _res_block(f"decoder_group.{i}.block.3", co)
    spec += _conv("out_conv", CHANNELS[GT_RES], 3, 3)
    cbs = normalize_codebooks(codebooks, n_e, e_dim)
    for k, (cs, ne, ed) in enumerate(cbs):                 # femasr_arch.py:280-299
        ch = CHANNELS[cs]
        spec.append((f"quantize_group.{k}.embedding.weight", (ne, ed), "codebook", ne))
        spec += _conv(f"before_quant_group.{k}", ch if k == 0 else 2 * ch, ed, 1)
        spec += _conv(f"after_quant_group.{k}.conv", ed if k == 0 else cbs[k - 1][2] + ed, ch, 3)
    return spec


def relative_position_index(ws: int = WINDOW) -> torch.Tensor:
    """The fixed [ws^2, ws^2] int64 buffer of network_swinir.py:91-101."""
    ar = torch.arange(ws)
    cy, cx = torch.meshgrid(ar, ar, indexing="ij")
    cy, cx = cy.flatten(), cx.flatten()
    dy = cy[:, None] - cy[None, :] + ws - 1
    dx = cx[:, None] - cx[None, :] + ws - 1
    return dy * (2 * ws - 1) + dx


def shift_attn_mask(H: int, W: int, ws: int = WINDOW, shift: int = WINDOW // 2) -> torch.Tensor:
    """0 / -100 shifted-window mask [nW, ws^2, ws^2] (what network_swinir.py:216-237 builds),
    computed from region ids: rows/cols in [0,H-ws) -> 0, [H-ws,H-shift) -> 1, [H-shift,H) -> 2."""
    def region(n):
        r = torch.zeros(n, dtype=torch.int64)
        r[n - ws:n - shift] = 1
        r[n - shift:] = 2
        return r
    ids = region(H)[:, None] * 3 + region(W)[None, :]
    ids = ids.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = ids[:, None, :] != ids[:, :, None]
    return torch.where(diff, torch.tensor(-100.0), torch.tensor(0.0))


def _gen(seed: int, name: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator()
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def random_state_dict(scale: int, e_dim: int, seed: int = 0, init: str = "default",
                      n_e: int = 1024, codebooks=None) -> Dict[str, torch.Tensor]:
    """Seeded random weights with the reference's default-init distributions.

    Each tensor is drawn from its own generator keyed by (seed, name), so the dict is reproducible
    anywhere without the reference.  ``init='default'``: exactly the reference's distributions
    (conv/linear U(+-1/sqrt(fan_in)); GN/LN weight 1 bias 0; rel-pos table trunc-normal(.02),
    network_swinir.py:111; codebook U(+-1/n_e), femasr_arch.py:33).  ``init='perturbed'``: norm affine
    parameters and the codebook get non-trivial values so tests exercise them.
    """
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind, fan_in in param_spec(scale, e_dim, n_e, codebooks=codebooks):
        g = _gen(seed, name)
        if kind in ("w", "b"):
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "norm_w":
            t = torch.ones(shape)
            if init == "perturbed":
                t = t + 0.2 * torch.randn(shape, generator=g)
        elif kind == "norm_b":
            t = torch.zeros(shape)
            if init == "perturbed":
                t = 0.2 * torch.randn(shape, generator=g)
        elif kind == "rpb":
            std = 0.02 if init == "default" else 0.5
            t = (torch.randn(shape, generator=g) * std).clamp_(-2 * std, 2 * std)
        elif kind == "rpi":
            t = relative_position_index()
        elif kind == "mask":
            t = shift_attn_mask(SWIN_INIT_RES, SWIN_INIT_RES)
        elif kind == "codebook":
            if init == "default":
                t = (torch.rand(shape, generator=g) * 2 - 1) / fan_in      # fan_in carries this codebook's n_e
            else:
                t = torch.randn(shape, generator=g) * 0.5
        else:
            raise ValueError(kind)
        sd[name] = t.contiguous()
    return sd
