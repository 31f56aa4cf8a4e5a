"""Timing of the shifted-window attention kernel and the LayerNorm operand staging at benchmark geometry
(B=32, 64x64 tokens, C=256) through the C ABI.  Usage (GPU box): python scripts/microbench_attn.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femasr_b200 import lib as L  # noqa: E402
from tests import gpu_util as G  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = L.load()


def timeit(fn, reps=5):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


B, H, W, C = 32, 64, 64, 256
qkv = torch.randn(B, H, W, 3 * C, device=dev)
table = torch.randn(225, 8, device=dev) * 0.02
bias = torch.empty(8, 64, 64, device=dev)
L.check(lib.femasr_expand_rel_bias_mma(G.p(table), G.p(bias), 8, G.S()))
oh = torch.empty(B, H, W, C, dtype=torch.float16, device=dev)
ol = torch.empty_like(oh)
for shift in (0, 4):
    ms = timeit(lambda: L.check(lib.femasr_window_attention_mma(G.p(qkv), G.p(bias), None, G.p(oh), G.p(ol), B, H, W, C, 8, shift, G.S())))
    gb = (qkv.numel() * 4 + oh.numel() * 4) / 1e9
    print(f"window_attention_mma shift {shift}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s (qkv read + split planes written)")
x = torch.randn(B, H, W, C, device=dev)
g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
ms = timeit(lambda: L.check(lib.femasr_tc_prepare(G.p(x), G.p(oh), G.p(ol), L.PRO_LN, None, None, G.p(g), G.p(b), B, H, W, C, 0, 1e-5, G.S())))
print(f"tc_prepare LN: {ms:.3f} ms  {x.numel() * 8 / ms / 1e6:.0f} GB/s")

xo = torch.randn(32, 512, 512, 64, device=dev)
wo = torch.randn(9 * 64 * 3, device=dev) * 0.05
bo = torch.randn(3, device=dev)
yo = torch.empty(32, 3, 512, 512, device=dev)
for name, fn in (("out_conv SIMT", lib.femasr_out_conv3x3), ("out_conv mma", lib.femasr_out_conv3x3_mma)):
    ms = timeit(lambda: L.check(fn(G.p(xo), G.p(wo), G.p(bo), G.p(yo), 32, 512, 512, 64, G.S())), reps=3)
    print(f"{name} 32x512x512x64 -> 3: {ms:.3f} ms  {(xo.numel() + yo.numel()) * 4 / ms / 1e6:.0f} GB/s")
