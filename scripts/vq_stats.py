"""GPU: timing and refinement statistics of the fused VQ on the REAL features of the benchmark network (x4, 128x128,
batch 32: 131072 rows).  Usage: python scripts/vq_stats.py > gpurun_out/vq_stats.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femasr_b200 import lib as L  # noqa: E402
from femasr_b200.net import NativeNet  # noqa: E402
from femasr_b200.spec import random_state_dict  # noqa: E402
from tests import gpu_util as G  # noqa: E402

dev = torch.device("cuda", 0)
sd = random_state_dict(4, 256, seed=0, init="default")
net = NativeNet(4, 1024, 256, gemm_path=1)
net.load_state_dict(sd, dev)
x = torch.rand(32, 3, 128, 128, generator=torch.Generator().manual_seed(1)).to(dev)
_, _, idx_net, taps = net.forward(x, taps=["z"])
z = taps["z"].reshape(-1, 256).contiguous()
cb = sd["quantize_group.0.embedding.weight"].to(dev).contiguous()
lib = L.load()
N, e_dim, n_e = z.shape[0], 256, 1024
a = torch.empty(N, device=dev); esq = torch.empty(n_e, device=dev)
L.check(lib.femasr_row_sumsq(z.data_ptr(), a.data_ptr(), N, e_dim, G.S()))
L.check(lib.femasr_row_sumsq(cb.data_ptr(), esq.data_ptr(), n_e, e_dim, G.S()))
hi, lo = G.tc_prepare(z.view(1, 1, N, e_dim))
blob = G.tc_pack(cb.view(n_e, e_dim, 1, 1))
cand = torch.empty(N, 4, 2, dtype=torch.int32, device=dev)
idx = torch.empty(N, dtype=torch.int64, device=dev)
zq = torch.empty(N, e_dim, device=dev); lrows = torch.empty(N, device=dev)
st = torch.zeros(3, dtype=torch.int32, device=dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


t_match = timed(lambda: L.check(lib.femasr_vq_match_tc(hi.data_ptr(), lo.data_ptr(), blob.data_ptr(), a.data_ptr(), esq.data_ptr(), cand.data_ptr(), N, n_e, e_dim, G.S())))
st.zero_()
L.check(lib.femasr_vq_finish(z.data_ptr(), a.data_ptr(), cand.data_ptr(), cb.data_ptr(), esq.data_ptr(), idx.data_ptr(), zq.data_ptr(), lrows.data_ptr(), st.data_ptr(), N, n_e, e_dim, G.S()))
torch.cuda.synchronize()
stats = st.cpu().tolist()
t_fin = timed(lambda: L.check(lib.femasr_vq_finish(z.data_ptr(), a.data_ptr(), cand.data_ptr(), cb.data_ptr(), esq.data_ptr(), idx.data_ptr(), zq.data_ptr(), lrows.data_ptr(), None, N, n_e, e_dim, G.S())))
d = cand[:, :, 0].contiguous().view(torch.float32)
print(json.dumps({"rows": N, "ms_vq_match_tc": round(t_match, 4), "ms_vq_finish": round(t_fin, 4),
                  "refined_rows": stats[0], "rescanned_rows": stats[1], "changed_by_refinement": stats[2],
                  "a_min_mean_max": [a.min().item(), a.mean().item(), a.max().item()],
                  "gap12_quantiles": torch.quantile((d[:, 1] - d[:, 0]).float().cpu(), torch.tensor([0.01, 0.05, 0.5])).tolist(),
                  "equal_to_engine_indices": bool(torch.equal(idx, idx_net.reshape(-1)))}))
