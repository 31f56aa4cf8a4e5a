"""Timing of the other BASELINE.json configurations (1 GPU, CUDA events, median of 3 after warm-up):
  config 2t: x4 128x128 B=32 through test()  (flip-pad to 144, crop)
  config 3 : x2 256x256 B=16 forward
  config 5 : x4 1024x1024 LR, test_tile(256, 32)  (16 tiles, 3 shape classes, same-shape tiles batched)
Prints one JSON object.  Usage (GPU box): python scripts/bench_configs.py > gpurun_out/configs.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr.archs import build_network  # noqa: E402
from femasr_b200.spec import random_state_dict  # noqa: E402


def flops_per_image(scale, h, w, e_dim):
    """Algorithmic FLOPs of one forward from the engine's own model (femasr_net_flops; no oracle import here)."""
    import ctypes as C
    from femasr_b200 import lib as L
    lib = L.load()
    cfg = L.NetConfig(scale, 1024, e_dim, 3, 1, 1, 1, 0, (C.c_int * 3)(), (C.c_int * 3)(), (C.c_int * 3)())
    hnd = C.c_void_p()
    L.check(lib.femasr_net_create(C.byref(cfg), C.byref(hnd)))
    try:
        return float(lib.femasr_net_flops(hnd, 1, h, w))
    finally:
        lib.femasr_net_destroy(hnd)

dev = torch.device("cuda", 0)


def net_for(scale, e_dim=256):
    net = build_network(dict(type="FeMaSRNet", codebook_params=[[32, 1024, e_dim]], LQ_stage=True, scale_factor=scale))
    net.load_state_dict(random_state_dict(scale, e_dim, seed=0, init="default"), strict=True)
    return net.to(dev).eval()


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


out = {}
g = torch.Generator().manual_seed(1)
n4 = net_for(4)
x = torch.rand(32, 3, 128, 128, generator=g).to(dev)
ms = timed(lambda: n4.test(x))
out["config2_test_x4_128_b32"] = {"ms": round(ms, 2), "images_per_s": round(32 / ms * 1e3, 1),
                                  "tflops_algorithmic": round(32 * flops_per_image(4, 144, 144, 256) / ms / 1e9, 1)}
ms = timed(lambda: n4(x))
out["config2_forward_x4_128_b32"] = {"ms": round(ms, 2), "images_per_s": round(32 / ms * 1e3, 1),
                                     "tflops_algorithmic": round(32 * flops_per_image(4, 128, 128, 256) / ms / 1e9, 1)}
xb = torch.rand(1, 3, 1024, 1024, generator=g).to(dev)
ms = timed(lambda: n4.test_tile(xb, 256, 32), reps=3, warm=1)
out["config5_tile_x4_1024"] = {"ms": round(ms, 2), "images_per_s": round(1e3 / ms, 3),
                               "tflops_algorithmic": round(46.05e6 * 1638400 / ms / 1e9, 1)}
del n4
torch.cuda.empty_cache()
n2 = net_for(2)
x2 = torch.rand(16, 3, 256, 256, generator=g).to(dev)
ms = timed(lambda: n2(x2))
out["config3_forward_x2_256_b16"] = {"ms": round(ms, 2), "images_per_s": round(16 / ms * 1e3, 1),
                                     "tflops_algorithmic": round(16 * flops_per_image(2, 256, 256, 256) / ms / 1e9, 1)}
print(json.dumps(out, indent=1))
