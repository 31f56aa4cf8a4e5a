#!/usr/bin/env python
"""Digest of the public-surface outputs for a few geometries with the library selected by FEMASR_LIB: two library builds
that claim to be arithmetic-identical (an epilogue / scheduling refactor) must print the same JSON.
    FEMASR_LIB=... python scripts/ab_digest.py > a.json"""
import hashlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr.archs.femasr_arch import FeMaSRNet  # noqa: E402
from femasr_b200.spec import random_state_dict  # noqa: E402


def dig(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


def main():
    dev = torch.device("cuda", 0)
    out = {}
    for scale, cb, shapes in ((4, [[32, 1024, 256]], [(4, 128, 128), (1, 96, 160)]), (2, [[32, 1024, 512]], [(2, 64, 96)])):
        net = FeMaSRNet(codebook_params=cb, LQ_stage=True, scale_factor=scale).to(dev).eval()
        net.load_state_dict(random_state_dict(scale, cb[0][2], seed=3, init="default"), strict=False)
        for (b, h, w) in shapes:
            x = torch.rand(b, 3, h, w, generator=torch.Generator().manual_seed(5)).to(dev)
            y, loss, _, idx = net(x)
            out[f"x{scale}_fwd_{b}x{h}x{w}"] = [dig(y), dig(idx[0]), float(loss)]
        x = torch.rand(1, 3, 75, 52, generator=torch.Generator().manual_seed(6)).to(dev)      # ragged: edge tiles everywhere
        out[f"x{scale}_test_75x52"] = [dig(net.test(x))]
        if scale == 4:
            x = torch.rand(1, 3, 200, 136, generator=torch.Generator().manual_seed(7)).to(dev)
            out["x4_tile_200x136"] = [dig(net.test_tile(x, tile_size=96, tile_pad=8))]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
