#!/usr/bin/env python
"""CPU experiment: can the layers BEHIND the VQ drop from three split-fp16 products to two?

Runs the oracle (tests-only infrastructure; this script is an offline study, not product code) with the operands of
selected conv layers rounded to fp16 (round-to-nearest, like the engine's hi plane):
    "w"  : weights rounded      == dropping a_hi*w_lo   (activations keep hi+lo)
    "a"  : activations rounded  == dropping a_lo*w_hi   (weights keep hi+lo; the A operand shrinks to ONE plane)
    "aw" : both                 == a single product
and reports output max-abs against the unmodified fp32 oracle at the benchmark geometry.  Indices cannot change (the
VQ input is untouched), so the only bar is 1e-3 on the output.

    python scripts/exp_two_product.py [--images 4] [--seed 0]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from femasr_b200.spec import random_state_dict  # noqa: E402
from oracle import femasr_oracle as O  # noqa: E402


def r16(t):
    return t.half().float()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=4)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--init", default="default")
    args = ap.parse_args()
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    sd = random_state_dict(4, 256, seed=args.seed, init=args.init)
    x = torch.rand(args.images, 3, 128, 128, generator=torch.Generator().manual_seed(1))
    d = O.encode_depth(4)
    groups = {
        "dec2 (64ch@512)": ("decoder_group.2.block.2", "decoder_group.2.block.3"),
        "dec2.up": ("decoder_group.2.block.1",),
        "dec1 (128ch@256)": ("decoder_group.1.block.2", "decoder_group.1.block.3"),
        "dec1.up": ("decoder_group.1.block.1",),
        "dec0 (256ch@128)": ("decoder_group.0.block.2", "decoder_group.0.block.3"),
        "dec0.up+after_quant": ("decoder_group.0.block.1", "after_quant_group"),
        "encup1 (256ch@128)": (f"multiscale_encoder.blocks.{d + 1}.",),
        "encup2 (128ch@256)": (f"multiscale_encoder.blocks.{d + 2}.",),
    }
    all_post = tuple(p for g in groups.values() for p in g)
    orig = O._conv
    state = {"prefixes": (), "mode": ""}

    def patched(sd_, p, xx, stride=1, pad=1):
        if any(p.startswith(pre) for pre in state["prefixes"]):
            w = sd_[p + ".weight"]
            if "w" in state["mode"]:
                w = r16(w)
            if "a" in state["mode"]:
                xx = r16(xx)
            return F.conv2d(xx, w, sd_[p + ".bias"], stride=stride, padding=pad)
        return orig(sd_, p, xx, stride, pad)

    O._conv = patched
    res = {}
    with torch.no_grad():
        want = O.encode_and_decode(sd, x, 4)[0]
        for mode in ("a", "w", "aw"):
            state["mode"] = mode
            state["prefixes"] = all_post
            got = O.encode_and_decode(sd, x, 4)[0]
            err = (got - want).abs()
            res[f"all post-VQ, {mode}"] = [err.max().item(), err.mean().item()]
            print(f"all post-VQ, mode {mode:2s}: max-abs {err.max():.3e} mean {err.mean():.3e}", flush=True)
        for mode in ("a", "w"):
            for gname, pre in groups.items():
                state["mode"] = mode
                state["prefixes"] = pre
                got = O.encode_and_decode(sd, x, 4)[0]
                err = (got - want).abs()
                res[f"{gname}, {mode}"] = [err.max().item(), err.mean().item()]
                print(f"{gname:24s} mode {mode}: max-abs {err.max():.3e} mean {err.mean():.3e}", flush=True)
    O._conv = orig
    print(json.dumps(res))


if __name__ == "__main__":
    main()
