#!/usr/bin/env python
"""CPU study: how reproducible are the ORACLE's own codebook indices at the benchmark geometry (x4, 128x128, default
init) when only the fp32 summation order changes - different thread counts and batch chunking make ATen/oneDNN pick
different blockings.  Any implementation whose summation order differs from one particular ATen-CPU run is exposed to
the same effect; this measures its size (index differences per 131072 rows = batch 32).

    python scripts/oracle_selfcheck.py [--images 32] > profiles/oracle_selfcheck_r2.json
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femasr_b200.spec import random_state_dict  # noqa: E402
from oracle import femasr_oracle as O  # noqa: E402


def run(sd, x, threads, chunk, dtype=torch.float32):
    torch.set_num_threads(threads)
    sdd = {k: v.to(dtype) for k, v in sd.items()} if dtype != torch.float32 else sd
    idx = []
    with torch.no_grad():
        for i in range(0, x.shape[0], chunk):
            idx.append(O.encode_and_decode(sdd, x[i:i + chunk].to(dtype), 4)[3][0])
    return torch.cat(idx)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=32)
    a = ap.parse_args()
    sd = random_state_dict(4, 256, seed=0, init="default")
    x = torch.rand(32, 3, 128, 128, generator=torch.Generator().manual_seed(1))[:a.images]
    ncpu = os.cpu_count() or 1
    base = run(sd, x, min(16, ncpu), 8)
    out = {"rows": int(base.numel()), "baseline": f"threads {min(16, ncpu)}, chunks of 8 images"}
    for name, (t, c, dt) in {"threads 1, chunks of 8": (1, 8, torch.float32),
                             f"threads {max(2, ncpu // 2)}, chunks of 1": (max(2, ncpu // 2), 1, torch.float32),
                             "float64 (truth), chunks of 4": (min(16, ncpu), 4, torch.float64)}.items():
        other = run(sd, x, t, c, dt)
        out[name] = int((other != base).sum())
        print(name, out[name], file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
