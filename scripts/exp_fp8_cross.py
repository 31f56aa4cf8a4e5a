#!/usr/bin/env python
"""CPU study for DESIGN.md section 8 item 1: the layers BEHIND the VQ with the two cross products of the split-fp16 scheme
evaluated on fp8 (e4m3) operands:   D = a_hi*w_hi  +  e4m3(a_lo)*e4m3(w_hi)  +  e4m3(a_hi)*e4m3(w_lo)
(hi = fp16 round-to-nearest, lo = the fp32 remainder; fp8 operands scaled per tensor by a power of two into e4m3's range).
On the tensor cores the fp8 products run at twice the fp16 rate: executed work 2.0 instead of 3.0 per algorithmic MMA.
Reports output max-abs against the fp32 oracle (bar 1e-3; indices cannot change).  Offline study, not product code.

    python scripts/exp_fp8_cross.py [--images 2]
"""
import argparse
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from femasr_b200.spec import random_state_dict  # noqa: E402
from oracle import femasr_oracle as O  # noqa: E402


def e4m3(t):
    """Round to e4m3 after a per-tensor power-of-two scale that puts max|t| in [128, 256); returns the dequantised tensor."""
    m = float(t.abs().max())
    if m == 0.0:
        return t
    s = 2.0 ** (7 - math.floor(math.log2(m)) - 1 + 1)          # max * s in [128, 256)
    return (t * s).to(torch.float8_e4m3fn).float() / s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=2)
    args = ap.parse_args()
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    sd = random_state_dict(4, 256, seed=0, init="default")
    x = torch.rand(args.images, 3, 128, 128, generator=torch.Generator().manual_seed(1))
    d = O.encode_depth(4)
    post = ("decoder_group", "after_quant_group", f"multiscale_encoder.blocks.{d + 1}.", f"multiscale_encoder.blocks.{d + 2}.")
    orig = O._conv
    mode = {"m": None}

    def patched(sd_, p, xx, stride=1, pad=1):
        if mode["m"] and p.startswith(post):
            w = sd_[p + ".weight"]
            a_hi, w_hi = xx.half().float(), w.half().float()
            a_lo, w_lo = xx - a_hi, w - w_hi
            y = F.conv2d(a_hi, w_hi, sd_[p + ".bias"], stride=stride, padding=pad)
            if mode["m"] == "fp8":
                y = y + F.conv2d(e4m3(a_lo), e4m3(w_hi), None, stride=stride, padding=pad)
                y = y + F.conv2d(e4m3(a_hi), e4m3(w_lo), None, stride=stride, padding=pad)
            elif mode["m"] in ("fixed_e4m3", "fixed_e5m2"):
                # the recipe a kernel can use without any per-tensor statistics of the activations:
                #   A8 = [e4m3(a_lo * 2^10) | e4m3(a_hi * 2^-2)],  B8 = [f8(w_hi * 2^(s-10)) | e4m3(w_lo * 2^(s+2))],  s: max|w| * 2^s in [512, 1024)
                # (scales: scripts/exp_fp8_scales.py)
                mx = float(w.abs().max())
                sc = 2.0 ** (10 - math.frexp(mx)[1])
                ws = w * sc
                wh = ws.half().float()
                wl = ws - wh
                f8hi = torch.float8_e4m3fn if mode["m"] == "fixed_e4m3" else torch.float8_e5m2
                q = lambda t, dt=torch.float8_e4m3fn: t.clamp(-448, 448).to(dt).float()
                y = F.conv2d(a_hi, wh, None, stride=stride, padding=pad)
                y = y + F.conv2d(q(a_lo * 1024.0), q(wh / 1024.0, f8hi), None, stride=stride, padding=pad)
                y = y + F.conv2d(q(a_hi * 0.25), q(wl * 4.0), None, stride=stride, padding=pad)
                y = y / sc + sd_[p + ".bias"].view(1, -1, 1, 1)
            elif mode["m"] == "exact3":
                y = y + F.conv2d(a_lo, w_hi, None, stride=stride, padding=pad) + F.conv2d(a_hi, w_lo, None, stride=stride, padding=pad)
            return y
        return orig(sd_, p, xx, stride, pad)

    O._conv = patched
    res = {}
    with torch.no_grad():
        want = O.encode_and_decode(sd, x, 4)[0]
        for m in ("exact3", "fp8", "fixed_e4m3", "fixed_e5m2", "hi_only"):
            mode["m"] = m
            got = O.encode_and_decode(sd, x, 4)[0]
            e = (got - want).abs()
            res[m] = {"max_abs": e.max().item(), "mean_abs": e.mean().item()}
            print(m, res[m], flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
