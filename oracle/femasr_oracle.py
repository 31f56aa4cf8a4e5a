"""CPU oracle for the FeMaSR inference hot path (TEST INFRASTRUCTURE ONLY).

This file is a functional restatement, in plain torch-CPU ops driven by a
``state_dict``, of what chaofengc/FeMaSR computes on the path
``FeMaSRNet.test_tile -> test -> encode_and_decode`` for the in-scope
configurations (norm 'gn', act 'silu'; one codebook at scale 32 or the multi-scale
variant with further codebooks at 64 / 128; LQ_stage=True with scale 2|4, and the HQ
autoencoder LQ_stage=False passed as scale 1; optional gt_indices loss branch).
It is NOT the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
only as the checker / the CPU arm.  The product path (``femasr_b200``) never
imports this module and fails loudly without its CUDA library.

Parity pin: the reference has no tests or golden vectors of its own (SURVEY.md
section 4), so this restatement is pinned against the reference itself, imported
unmodified from /root/reference in the build container by
``tests/golden/make_golden.py``; the resulting small input/output fixtures are
committed under ``tests/golden/`` and re-checked by ``tests/test_oracle.py``.

All arithmetic of the reference on this path is PyTorch ATen (un-vendored,
requirements.txt:12 ``torch>=1.7``; container torch 2.11.0), so ATen CPU ops
are the faithful arithmetic to restate with; every function cites the reference
file:line it follows.  ``dtype=torch.float64`` runs the same graph in double for
error-budget studies (then the result is "truth", not the parity target).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

GN_GROUPS = 32        # fema_utils.py:21-22
GN_EPS = 1e-6         # fema_utils.py:22
LN_EPS = 1e-5         # nn.LayerNorm default, network_swinir.py:199,207
WINDOW = 8            # femasr_arch.py:118 window_size=8
HEADS = 8             # femasr_arch.py:117 num_heads=8
SWIN_DEPTH = 6        # femasr_arch.py:116 blk_depth=6
N_RSTB = 4            # femasr_arch.py:122
VQ_BETA = 0.25        # femasr_arch.py:26


# --------------------------------------------------------------------------- helpers
def _conv(sd: SD, p: str, x: Tensor, stride: int = 1, pad: int = 1) -> Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=pad)


def res_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """fema_utils.py:65-84: x + conv3(SiLU(GN(conv3(SiLU(GN(x))))))."""
    t = F.silu(F.group_norm(x, GN_GROUPS, sd[p + ".conv.0.norm.weight"], sd[p + ".conv.0.norm.bias"], GN_EPS))
    t = _conv(sd, p + ".conv.2", t)
    t = F.silu(F.group_norm(t, GN_GROUPS, sd[p + ".conv.3.norm.weight"], sd[p + ".conv.3.norm.bias"], GN_EPS))
    t = _conv(sd, p + ".conv.5", t)
    return t + x


def upsample2(x: Tensor) -> Tensor:
    """nn.Upsample(scale_factor=2) (nearest), femasr_arch.py:172,202."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


# --------------------------------------------------------------------------- swin
def window_partition(x: Tensor, ws: int) -> Tensor:
    """network_swinir.py:33-45.  x [B,H,W,C] -> [B*nW, ws*ws, C]."""
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def window_reverse(w: Tensor, ws: int, H: int, W: int) -> Tensor:
    """network_swinir.py:48-62."""
    B = w.shape[0] // ((H // ws) * (W // ws))
    x = w.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def shift_mask(H: int, W: int, ws: int, shift: int, dtype) -> Tensor:
    """network_swinir.py:216-237 (calculate_mask): 0 / -100 mask [nW, ws*ws, ws*ws]."""
    img = torch.zeros((1, H, W, 1), dtype=dtype)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def rel_pos_index(ws: int) -> Tensor:
    """network_swinir.py:91-101."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def window_attention(sd: SD, p: str, xw: Tensor, mask) -> Tensor:
    """network_swinir.py:114-145.  xw [B_, 64, C]."""
    B_, N, C = xw.shape
    hd = C // HEADS
    qkv = F.linear(xw, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    qkv = qkv.reshape(B_, N, 3, HEADS, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (hd ** -0.5)                                   # :124
    attn = q @ k.transpose(-2, -1)                         # :125
    idx = rel_pos_index(WINDOW).view(-1)
    bias = sd[p + ".relative_position_bias_table"][idx].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)                        # :130
    if mask is not None:                                   # :132-136
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, HEADS, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, HEADS, N, N)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B_, N, C)       # :142
    return F.linear(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def swin_block(sd: SD, p: str, x: Tensor, hw: Tuple[int, int], shift: int) -> Tensor:
    """network_swinir.py:239-279.  x [B, HW, C] tokens."""
    H, W = hw
    B, L, C = x.shape
    shortcut = x
    t = F.layer_norm(x, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], LN_EPS).view(B, H, W, C)
    if shift > 0:
        t = torch.roll(t, shifts=(-shift, -shift), dims=(1, 2))
        mask = shift_mask(H, W, WINDOW, shift, x.dtype)
    else:
        mask = None
    tw = window_partition(t, WINDOW)
    aw = window_attention(sd, p + ".attn", tw, mask)
    t = window_reverse(aw, WINDOW, H, W)
    if shift > 0:
        t = torch.roll(t, shifts=(shift, shift), dims=(1, 2))
    x = shortcut + t.reshape(B, L, C)
    h = F.layer_norm(x, (C,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], LN_EPS)
    h = F.linear(h, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])
    h = F.gelu(h)                                          # exact erf, network_swinir.py:15,20
    h = F.linear(h, sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
    return x + h


def rstb(sd: SD, p: str, x: Tensor, hw: Tuple[int, int]) -> Tensor:
    """network_swinir.py:481-482: 6 blocks -> NCHW -> conv3x3 -> tokens, + x."""
    H, W = hw
    B, L, C = x.shape
    t = x
    for i in range(SWIN_DEPTH):
        shift = 0 if i % 2 == 0 else WINDOW // 2           # network_swinir.py:380
        t = swin_block(sd, f"{p}.residual_group.blocks.{i}", t, hw, shift)
    t = t.transpose(1, 2).reshape(B, C, H, W)
    t = _conv(sd, p + ".conv", t)
    return t.flatten(2).transpose(1, 2) + x


def swin_layers(sd: SD, p: str, x: Tensor) -> Tensor:
    """femasr_arch.py:126-132."""
    B, C, H, W = x.shape
    if H % WINDOW or W % WINDOW:
        raise RuntimeError(f"Swin stage {H}x{W} not a multiple of window {WINDOW}")  # view() error in the reference
    t = x.reshape(B, C, H * W).transpose(1, 2)
    for i in range(N_RSTB):
        t = rstb(sd, f"{p}.swin_blks.{i}", t, (H, W))
    return t.transpose(1, 2).reshape(B, C, H, W)


# --------------------------------------------------------------------------- encoder / vq / decoder
def encode_depth(scale: int, gt_res: int = 256, cb_scale: int = 32) -> int:
    """femasr_arch.py:256."""
    return int(math.log2(gt_res // scale // cb_scale))


def multiscale_encoder(sd: SD, x: Tensor, scale: int) -> List[Tensor]:
    """femasr_arch.py:184-192.  Returns the per-block outputs.  scale 4|2: LQ stage (down, Swin, two up blocks);
    scale 1: HQ stage (LQ_stage=False: three down blocks only, :166-182)."""
    p = "multiscale_encoder"
    d = encode_depth(scale)
    outs = []
    x = _conv(sd, p + ".in_conv", x, 1, 1)                 # 4x4 p1, :150
    for i in range(d):                                     # :156-164
        x = _conv(sd, f"{p}.blocks.{i}.0", x, 2, 1)
        x = res_block(sd, f"{p}.blocks.{i}.1", x)
        x = res_block(sd, f"{p}.blocks.{i}.2", x)
        outs.append(x)
    if scale == 1:
        return outs
    x = swin_layers(sd, f"{p}.blocks.{d}", x)              # :166-167
    outs.append(x)
    for j in (d + 1, d + 2):                               # :168-180
        x = _conv(sd, f"{p}.blocks.{j}.1", upsample2(x))
        x = res_block(sd, f"{p}.blocks.{j}.2", x)
        x = res_block(sd, f"{p}.blocks.{j}.3", x)
        outs.append(x)
    return outs


def vq_dist(z: Tensor, cb: Tensor) -> Tensor:
    """femasr_arch.py:35-38: (sum z^2 + sum e^2) - 2 z e^T, in that operator order."""
    return torch.sum(z ** 2, dim=1, keepdim=True) + torch.sum(cb ** 2, dim=1) - 2 * torch.matmul(z, cb.t())


def gram_loss(x: Tensor, y: Tensor) -> Tensor:
    """femasr_arch.py:40-48: mean squared difference of the per-image [c,c] Gram matrices (x, y: [b,h,w,c])."""
    b, h, w, c = x.shape
    x = x.reshape(b, h * w, c)
    y = y.reshape(b, h * w, c)
    gmx = x.transpose(1, 2) @ x / (h * w)
    gmy = y.transpose(1, 2) @ y / (h * w)
    return (gmx - gmy).square().mean()


def vector_quantize(cb: Tensor, z_nchw: Tensor, gt_indices: Tensor | None = None, lq: bool = True):
    """femasr_arch.py:50-100.  Returns (z_q NCHW, loss, idx [B,1,h,w]).  ``gt_indices`` (the HQ stage's codes,
    any shape with B*h*w entries) only changes the loss, and only in the LQ stage (:84-92)."""
    z = z_nchw.permute(0, 2, 3, 1).contiguous()
    zf = z.view(-1, cb.shape[1])
    d = vq_dist(zf, cb)
    idx = torch.argmin(d, dim=1)                           # lowest index wins ties, :66
    zq = cb[idx].view(z.shape)                             # == onehot @ codebook, :67-82
    e_lat = torch.mean((zq - z) ** 2)
    q_lat = torch.mean((zq - z) ** 2)
    if lq and gt_indices is not None:                      # :70-78, :87-90
        zq_gt = cb[gt_indices.reshape(-1)].view(z.shape)
        loss = VQ_BETA * ((zq_gt - z) ** 2).mean()
        loss = loss + gram_loss(z, zq_gt)
    else:
        loss = q_lat + e_lat * VQ_BETA                     # :92
    zq = z + (zq - z)                                      # straight-through, :95 (not bit-equal to zq)
    zq = zq.permute(0, 3, 1, 2).contiguous()
    return zq, loss, idx.reshape(zq.shape[0], 1, zq.shape[2], zq.shape[3])


def decoder_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """femasr_arch.py:195-211."""
    x = _conv(sd, p + ".block.1", upsample2(x))
    x = res_block(sd, p + ".block.2", x)
    return res_block(sd, p + ".block.3", x)


def encode_and_decode(sd: SD, x: Tensor, scale: int, taps: dict | None = None, cb_scales=(32,),
                      gt_indices=None, use_residual: bool = True, use_quantize: bool = True):
    """femasr_arch.py:311-374.  ``cb_scales``: decoder resolutions that carry a codebook (first one 32, :231);
    ``gt_indices``: list of index maps, one per codebook (loss only, :339-342).

    Returns (out_img, codebook_loss, semantic_loss, [indices per codebook]).  ``taps`` (optional dict)
    receives the stage-boundary tensors used by the stage-level parity tests.
    """
    lq = scale != 1
    outs = multiscale_encoder(sd, x, scale)
    feats = outs[-3:] if lq else outs[::-1]                # :313-316
    if taps is not None:
        taps.update(enc0=feats[0], enc1=feats[1], enc2=feats[2])
    losses, indices = [], []
    k = 0
    prev_dec = prev_quant = None
    t = feats[0]
    for i in range(3):                                     # max_depth = 3, :255
        if 32 * 2 ** i in cb_scales:                       # :330-331
            bq = feats[i] if prev_dec is None else torch.cat((feats[i], prev_dec), dim=1)     # :332-335
            z = _conv(sd, f"before_quant_group.{k}", bq, 1, 0)                                 # 1x1, :337
            zq, loss, idx = vector_quantize(sd[f"quantize_group.{k}.embedding.weight"], z,
                                            None if gt_indices is None else gt_indices[k], lq)  # :339-342
            if not use_quantize:
                zq = z                                     # :349-350
            aq = zq                                        # CombineQuantBlock, fema_utils.py:92-99
            if prev_quant is not None:
                aq = torch.cat((zq, F.interpolate(prev_quant, zq.shape[2:])), dim=1)
            t = _conv(sd, f"after_quant_group.{k}.conv", aq)
            if taps is not None and k == 0:
                taps.update(z=z, zq=zq, after_quant=t)
            losses.append(loss)
            indices.append(idx)
            k += 1
            prev_quant = zq
        elif lq and use_residual:
            t = t + feats[i]                               # :361-362
        t = decoder_block(sd, f"decoder_group.{i}", t)
        prev_dec = t
        if taps is not None:
            taps[f"dec{i}"] = t
    out = _conv(sd, "out_conv", t)                         # :369
    loss = sum(losses)                                     # :371
    return out, loss, loss * 0, indices


def decode_indices(sd: SD, indices: Tensor) -> Tensor:
    """femasr_arch.py:376-385 + get_codebook_entry :102-112."""
    assert indices.dim() == 4, f"shape of indices must be (b, 1, h, w), but got {indices.shape}"
    b, _, h, w = indices.shape
    cb = sd["quantize_group.0.embedding.weight"]
    zq = cb[indices.flatten()].view(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
    t = _conv(sd, "after_quant_group.0.conv", zq)
    for i in range(3):
        t = decoder_block(sd, f"decoder_group.{i}", t)
    return _conv(sd, "out_conv", t)


def flip_pad(x: Tensor, scale: int) -> Tensor:
    """femasr_arch.py:455-461: flip-pad to (h//wsz+1)*wsz (always pads, even when h is a multiple of wsz)."""
    wsz = 8 // scale * 8
    _, _, h, w = x.shape
    hp = (h // wsz + 1) * wsz - h
    wp = (w // wsz + 1) * wsz - w
    x = torch.cat([x, torch.flip(x, [2])], 2)[:, :, : h + hp, :]
    return torch.cat([x, torch.flip(x, [3])], 3)[:, :, :, : w + wp]


def test(sd: SD, x: Tensor, scale: int) -> Tensor:
    """femasr_arch.py:449-468: flip-pad, run, crop."""
    _, _, h, w = x.shape
    out = encode_and_decode(sd, flip_pad(x, scale), scale)[0]
    return out[..., : h * scale, : w * scale]


def tile_plan(height: int, width: int, tile_size: int, tile_pad: int):
    """femasr_arch.py:401-426, 431-441: the (input window, output window, crop) of every tile."""
    plan = []
    for y in range(math.ceil(height / tile_size)):
        for x in range(math.ceil(width / tile_size)):
            sx, sy = x * tile_size, y * tile_size
            ex, ey = min(sx + tile_size, width), min(sy + tile_size, height)
            sxp, exp_ = max(sx - tile_pad, 0), min(ex + tile_pad, width)
            syp, eyp = max(sy - tile_pad, 0), min(ey + tile_pad, height)
            plan.append(dict(in_win=(syp, eyp, sxp, exp_), out_win=(sy, ey, sx, ex),
                             crop=(sy - syp, sx - sxp, ey - sy, ex - sx)))
    return plan


def test_tile(sd: SD, x: Tensor, scale: int, tile_size: int = 240, tile_pad: int = 16) -> Tensor:
    """femasr_arch.py:387-447."""
    b, c, h, w = x.shape
    out = x.new_zeros((b, c, h * scale, w * scale))
    for t in tile_plan(h, w, tile_size, tile_pad):
        y0, y1, x0, x1 = t["in_win"]
        o = test(sd, x[:, :, y0:y1, x0:x1], scale)
        cy, cx, th, tw = t["crop"]
        oy0, oy1, ox0, ox1 = t["out_win"]
        out[:, :, oy0 * scale:oy1 * scale, ox0 * scale:ox1 * scale] = \
            o[:, :, cy * scale:(cy + th) * scale, cx * scale:(cx + tw) * scale]
    return out


# --------------------------------------------------------------------------- random-init weights
def flops_per_image(scale: int, h: int, w: int, e_dim: int) -> float:
    assert scale in (2, 4), "LQ-stage model"
    """Algorithmic FLOPs (2*MAC over conv + linear + QK^T/PV + VQ distance) of encode_and_decode on
    one h x w LR image.  Matches SURVEY.md section 8a [probe]: x4 128x128 e256 = 754.53 GF."""
    d = encode_depth(scale)
    c_in = {4: 256, 2: 128}[scale]
    f = 2.0 * 3 * 16 * c_in * (h - 1) * (w - 1)                      # in_conv
    ch, hh, ww = c_in, h, w
    for i in range(d):
        co = 256
        hh, ww = hh // 2, ww // 2
        f += 2.0 * 9 * ch * co * hh * ww + 4 * 2.0 * 9 * co * co * hh * ww
        ch = co
    px = hh * ww
    lin = 2.0 * 256 * (768 + 256 + 1024 + 1024) * px
    att = 2 * 2.0 * 64 * 256 * px
    f += N_RSTB * (SWIN_DEPTH * (lin + att) + 2.0 * 9 * 256 * 256 * px)
    f += 2.0 * 256 * e_dim * px + 2.0 * 1024 * e_dim * px + 2.0 * 9 * e_dim * 256 * px
    for (ci, co, m) in ((256, 256, 2), (256, 128, 4)):               # encoder up branches
        f += (2.0 * 9 * ci * co + 4 * 2.0 * 9 * co * co) * px * m * m
    for (ci, co, m) in ((256, 256, 2), (256, 128, 4), (128, 64, 8)):  # decoder
        f += (2.0 * 9 * ci * co + 4 * 2.0 * 9 * co * co) * px * m * m
    f += 2.0 * 9 * 64 * 3 * px * 64
    return f
