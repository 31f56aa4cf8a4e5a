"""GPU: every exported operator kernel against the same ATen op on CPU fp32 (the arithmetic the
reference runs), through the C ABI.  Tolerances are fp32-summation-order bounds, written per test."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from femasr_b200 import lib as L
from femasr_b200.spec import relative_position_index, shift_attn_mask
from oracle import femasr_oracle as O
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def close(got, want, atol, what=""):
    err = (got.cpu() - want).abs().max().item()
    assert err <= atol, f"{what}: max-abs {err:.3e} > {atol:.1e}"


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,up", [
    (2, 16, 24, 64, 64, 1, 0), (1, 9, 11, 128, 256, 1, 0), (2, 15, 13, 256, 256, 2, 0),
    (1, 8, 8, 256, 128, 1, 1), (1, 6, 10, 128, 64, 1, 1), (3, 8, 8, 512, 256, 1, 0)])
def test_conv3x3(cuda, B, H, W, Cin, Cout, stride, up):
    x, w, b = rnd(B, Cin, H, W, seed=1), rnd(Cout, Cin, 3, 3, seed=2, scale=0.05), rnd(Cout, seed=3)
    xin = O.upsample2(x) if up else x
    want = F.conv2d(xin, w, b, stride=stride, padding=1)
    y = G.igemm(G.nhwc(x).to(cuda), G.pack_weight(w.to(cuda)), b.to(cuda), B, H, W, Cin, Cout, 3, stride, up)
    close(G.nchw(y), want, 2e-5 * (Cin * 9) ** 0.5, "conv3x3")


def test_conv3x3_gn_silu_residuals(cuda):
    B, H, W, Cc = 2, 12, 20, 128
    x, w, b = rnd(B, Cc, H, W, seed=4, scale=2.0) + 0.5, rnd(Cc, Cc, 3, 3, seed=5, scale=0.03), rnd(Cc, seed=6)
    gamma, beta = 1 + 0.2 * rnd(Cc, seed=7), 0.2 * rnd(Cc, seed=8)
    r1, r2 = rnd(B, Cc, H, W, seed=9), rnd(B, Cc, H, W, seed=10)
    t = F.silu(F.group_norm(x, 32, gamma, beta, 1e-6))
    want = (F.conv2d(t, w, b, padding=1) + r1) + r2
    xg = G.nhwc(x).to(cuda)
    sc, sh = G.gn_tables(xg, gamma.to(cuda), beta.to(cuda))
    y = G.igemm(xg, G.pack_weight(w.to(cuda)), b.to(cuda), B, H, W, Cc, Cc, 3, 1, 0, L.PRO_GN_SILU, sc, sh,
                res1=G.nhwc(r1).to(cuda), res2=G.nhwc(r2).to(cuda))
    close(G.nchw(y), want, 1e-4, "gn+silu conv")


@pytest.mark.parametrize("C_", [64, 128, 256])
def test_gn_tables(cuda, C_):
    B, H, W = 3, 37, 29    # HW = 1073: partial chunks
    x = rnd(B, C_, H, W, seed=11, scale=3.0) + 1.5
    gamma, beta = 1 + 0.2 * rnd(C_, seed=12), 0.2 * rnd(C_, seed=13)
    want = F.group_norm(x, 32, gamma, beta, 1e-6)
    sc, sh = G.gn_tables(G.nhwc(x).to(cuda), gamma.to(cuda), beta.to(cuda))
    got = x * sc.cpu()[:, :, None, None] + sh.cpu()[:, :, None, None]
    close(got, want, 5e-6, "group norm")


def test_linear_ln_gelu_residual(cuda):
    M, K, N = 333, 256, 1024
    x, w, b = rnd(M, K, seed=14, scale=2.0) + 0.3, rnd(N, K, seed=15, scale=0.05), rnd(N, seed=16)
    gamma, beta = 1 + 0.2 * rnd(K, seed=17), 0.2 * rnd(K, seed=18)
    want = F.gelu(F.linear(F.layer_norm(x, (K,), gamma, beta, 1e-5), w, b))
    xg = x.to(cuda)
    mu, rs = G.ln_stats(xg)
    y = G.igemm(xg, G.pack_weight(w.view(N, K, 1, 1).to(cuda)), b.to(cuda), 1, 1, M, K, N, 1, 1, 0, L.PRO_LN, mu, rs,
                gamma.to(cuda), beta.to(cuda), act=L.ACT_GELU)
    close(y.view(M, N), want, 5e-5, "ln+linear+gelu")
    # fc2-like: K=1024 -> 256 with in-place residual
    w2, b2 = rnd(K, N, seed=19, scale=0.03), rnd(K, seed=20)
    res = rnd(M, K, seed=21).to(cuda)
    want2 = F.linear(want, w2, b2) + res.cpu()
    h = want.to(cuda).contiguous()
    w2p, b2g = G.pack_weight(w2.view(K, N, 1, 1).to(cuda)), b2.to(cuda)
    a = L.IgemmArgs(h.data_ptr(), w2p.data_ptr(), b2g.data_ptr(),
                    res.data_ptr(), None, res.data_ptr(), None, None, None, None, 1, 1, M, N, K, 1, 1, 0, 0, 0)
    L.check(L.load().femasr_igemm_simt(C.byref(a), G.S()))
    close(res.view(M, K), want2, 1e-4, "linear + in-place residual")


@pytest.mark.parametrize("H,W,shift", [(16, 24, 0), (16, 24, 4), (8, 8, 4), (72, 8, 4)])
def test_window_attention(cuda, H, W, shift):
    B, Cc = 2, 256
    qkv = rnd(B, H * W, 3 * Cc, seed=22)
    table = rnd(225, 8, seed=23, scale=0.5)
    # reference semantics via the oracle helpers (network_swinir.py:114-145,239-279)
    t = qkv.view(B, H, W, 3 * Cc)
    if shift:
        t = torch.roll(t, (-shift, -shift), (1, 2))
    tw = O.window_partition(t, 8)
    q, k, v = tw.reshape(-1, 64, 3, 8, 32).permute(2, 0, 3, 1, 4)
    attn = (q * 32 ** -0.5) @ k.transpose(-2, -1)
    bias = table[relative_position_index().view(-1)].view(64, 64, 8).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if shift:
        mask = O.shift_mask(H, W, 8, shift, torch.float32)
        assert torch.equal(mask, shift_attn_mask(H, W, 8, shift))
        nW = mask.shape[0]
        attn = (attn.view(-1, nW, 8, 64, 64) + mask[None, :, None]).view(-1, 8, 64, 64)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, 64, Cc)
    o = O.window_reverse(o, 8, H, W)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    want = o.reshape(B, H * W, Cc)
    lib = L.load()
    tb = table.to(cuda)
    full = torch.empty(8, 64, 64, device=cuda)
    L.check(lib.femasr_expand_rel_bias(tb.data_ptr(), full.data_ptr(), 8, G.S()))
    close(full, bias, 0.0, "rel-pos bias expansion")
    out = torch.empty(B, H * W, Cc, device=cuda)
    qkvg = qkv.to(cuda)
    L.check(lib.femasr_window_attention(qkvg.data_ptr(), full.data_ptr(), out.data_ptr(), B, H, W, Cc, 8, shift, G.S()))
    close(out, want, 2e-5, "window attention")
    out2 = torch.zeros(B, H * W, Cc, device=cuda)
    frag = torch.empty(8 * 64 * 64, device=cuda)         # the same bias in the mma kernel's accumulator-fragment order
    L.check(lib.femasr_expand_rel_bias_mma(tb.data_ptr(), frag.data_ptr(), 8, G.S()))
    fr = frag.view(8, 4, 8, 8, 4, 2, 2).cpu()            # [h][warp][nt][g][c][row half][col]
    rows = (16 * torch.arange(4)[:, None, None] + torch.arange(8)[None, :, None] + 8 * torch.arange(2)[None, None, :])   # [warp][g][rh]
    cols = (8 * torch.arange(8)[:, None, None] + 2 * torch.arange(4)[None, :, None] + torch.arange(2)[None, None, :])    # [nt][c][e]
    want_fr = bias[:, rows[:, None, :, None, :, None], cols[None, :, None, :, None, :]]
    assert torch.equal(fr, want_fr), "fragment-order bias expansion"
    L.check(lib.femasr_window_attention_mma(qkvg.data_ptr(), frag.data_ptr(), out2.data_ptr(), None, None, B, H, W, Cc, 8, shift, G.S()))
    close(out2, want, 2e-5, "window attention (mma.sync split-fp16)")
    oh = torch.zeros(B, H * W, Cc, dtype=torch.float16, device=cuda)
    ol = torch.zeros_like(oh)
    L.check(lib.femasr_window_attention_mma(qkvg.data_ptr(), frag.data_ptr(), None, oh.data_ptr(), ol.data_ptr(), B, H, W, Cc, 8, shift, G.S()))
    close(oh.float() + ol.float(), want, 2e-5, "window attention (split fp16 planes out)")


@pytest.mark.parametrize("e_dim,init", [(256, "tiny"), (512, "tiny"), (256, "wide")])
def test_vq_select_bit_exact(cuda, e_dim, init):
    """Indices must be bit-exact given the same z and the same z.e^T (here: ATen's own matmul result, so
    the test isolates the distance formula + tie rule; ties are frequent with the U(+-1/1024) codebook)."""
    N, n_e = 4096, 1024
    g = torch.Generator().manual_seed(24)
    z = torch.randn(N, e_dim, generator=g) * 1.1
    cb = (torch.rand(n_e, e_dim, generator=g) * 2 - 1) / n_e if init == "tiny" else torch.randn(n_e, e_dim, generator=g)
    zc = z @ cb.t()
    d = O.vq_dist(z, cb)
    assert torch.equal(d, torch.sum(z ** 2, 1, keepdim=True) + torch.sum(cb ** 2, 1) - 2 * zc)
    want_idx = torch.argmin(d, 1)
    lib = L.load()
    zg, cbg, zcg = z.to(cuda), cb.to(cuda), zc.to(cuda)
    esq = torch.sum(cb ** 2, 1).to(cuda)     # ATen's B_j: isolates the select kernel
    idx = torch.empty(N, dtype=torch.int64, device=cuda)
    zq = torch.empty(N, e_dim, device=cuda)
    lrows = torch.empty(N, device=cuda)
    # A = sum z^2 is computed in-kernel with a different summation order than ATen; the argmin is invariant
    # to 1-ulp shifts of A (all d_j move together), see DESIGN.md "VQ rounding".
    L.check(lib.femasr_vq_select(zg.data_ptr(), zcg.data_ptr(), cbg.data_ptr(), esq.data_ptr(), idx.data_ptr(),
                                 zq.data_ptr(), lrows.data_ptr(), N, n_e, e_dim, 0, G.S()))
    mism = (idx.cpu() != want_idx).sum().item()
    assert mism == 0, f"{mism}/{N} index mismatches"
    e = cb[want_idx]
    want_zq = z + (e - z)
    assert torch.equal(zq.cpu(), want_zq), "straight-through z + (e - z) must be bit-exact"
    loss = torch.empty((), device=cuda)
    L.check(lib.femasr_sum_scaled(lrows.data_ptr(), loss.data_ptr(), N, 1.25 / (N * e_dim), G.S()))
    want_loss = torch.mean((e - z) ** 2) * 1.25
    assert abs(loss.item() - want_loss.item()) <= 2e-6 * abs(want_loss.item())
    # own esq kernel: fp32-accurate
    esq2 = torch.empty(n_e, device=cuda)
    L.check(lib.femasr_row_sumsq(cbg.data_ptr(), esq2.data_ptr(), n_e, e_dim, G.S()))
    assert (esq2.cpu() - esq.cpu()).abs().max().item() <= 4e-7 * esq.abs().max().item()


def test_vq_ties_pick_lowest_index(cuda):
    N, n_e, e_dim = 64, 1024, 256
    z = rnd(N, e_dim, seed=25)
    cb = rnd(n_e, e_dim, seed=26)
    cb[700] = cb[3]
    cb[512] = cb[3]          # exact duplicates of code 3
    z[:] = cb[3] + 1e-3 * rnd(N, e_dim, seed=27)
    zc = z @ cb.t()
    lib = L.load()
    idx = torch.empty(N, dtype=torch.int64, device=cuda)
    esq = torch.sum(cb ** 2, 1).to(cuda)
    zg, zcg, cbg = z.to(cuda), zc.to(cuda), cb.to(cuda)
    L.check(lib.femasr_vq_select(zg.data_ptr(), zcg.data_ptr(), cbg.data_ptr(),
                                 esq.data_ptr(), idx.data_ptr(), None, None, N, n_e, e_dim, 0, G.S()))
    assert (idx.cpu() == 3).all()


def test_in_conv_out_conv(cuda):
    lib = L.load()
    B, H, W = 2, 18, 22
    for cout in (256, 128):
        x, w, b = torch.rand(B, 3, H, W), rnd(cout, 3, 4, 4, seed=28, scale=0.15), rnd(cout, seed=29)
        want = F.conv2d(x, w, b, padding=1)
        y = torch.empty(B, H - 1, W - 1, cout, device=cuda)
        xg, wp, bg = x.to(cuda), G.pack_weight(w.to(cuda)), b.to(cuda)
        L.check(lib.femasr_in_conv4x4(xg.data_ptr(), wp.data_ptr(), bg.data_ptr(), y.data_ptr(), B, 3, H, W, cout, G.S()))
        close(G.nchw(y), want, 2e-6 * 48 ** 0.5 * 4, "in_conv")
    x, w, b = rnd(B, 64, 21, 130, seed=30), rnd(3, 64, 3, 3, seed=31, scale=0.05), rnd(3, seed=32)
    want = F.conv2d(x, w, b, padding=1)
    y = torch.empty(B, 3, 21, 130, device=cuda)
    xg, wp, bg = G.nhwc(x).to(cuda), G.pack_weight(w.to(cuda)), b.to(cuda)
    L.check(lib.femasr_out_conv3x3(xg.data_ptr(), wp.data_ptr(), bg.data_ptr(), y.data_ptr(), B, 21, 130, 64, G.S()))
    close(y, want, 2e-5, "out_conv")
    # tensor-core variant (mma.sync, split fp16, horizontal taps folded into N): ragged sizes around the 8x30 tile
    for (Hh, Ww) in ((21, 130), (8, 30), (9, 31), (5, 7), (64, 64)):
        x2 = rnd(B, 64, Hh, Ww, seed=33, scale=1.5)
        want2 = F.conv2d(x2.double(), w.double(), b.double(), padding=1)
        y2 = torch.full((B, 3, Hh, Ww), float("nan"), device=cuda)
        x2g = G.nhwc(x2).to(cuda)
        L.check(lib.femasr_out_conv3x3_mma(x2g.data_ptr(), wp.data_ptr(), bg.data_ptr(), y2.data_ptr(), B, Hh, Ww, 64, G.S()))
        err = (y2.cpu().double() - want2).abs().max().item()
        assert err <= 5e-6 * want2.abs().max().item(), f"out_conv_mma {Hh}x{Ww}: max-abs {err:.3e}"


def test_flip_pad_copy_window_layouts(cuda):
    lib = L.load()
    x = torch.rand(2, 3, 40, 24)
    hp, wp = 48, 32
    want = torch.cat([x, torch.flip(x, [2])], 2)[:, :, :hp, :]
    want = torch.cat([want, torch.flip(want, [3])], 3)[:, :, :, :wp]
    y = torch.empty(2, 3, hp, wp, device=cuda)
    xg = x.to(cuda)
    L.check(lib.femasr_flip_pad(xg.data_ptr(), y.data_ptr(), 2, 3, 40, 24, hp, wp, G.S()))
    assert torch.equal(y.cpu(), want)
    dst = torch.zeros(2, 3, 30, 30, device=cuda)
    L.check(lib.femasr_copy_window(y.data_ptr(), dst.data_ptr(), 2, 3, hp, wp, 30, 30, 5, 7, 2, 3, 20, 11, G.S()))
    ref = torch.zeros(2, 3, 30, 30)
    ref[:, :, 2:22, 3:14] = want[:, :, 5:25, 7:18]
    assert torch.equal(dst.cpu(), ref)
    a = torch.rand(2, 5, 7, 9)
    b = torch.empty(2, 7, 9, 5, device=cuda)
    ag = a.to(cuda)
    L.check(lib.femasr_nchw_to_nhwc(ag.data_ptr(), b.data_ptr(), 2, 5, 7, 9, G.S()))
    assert torch.equal(b.cpu(), a.permute(0, 2, 3, 1))
    c = torch.empty(2, 5, 7, 9, device=cuda)
    L.check(lib.femasr_nhwc_to_nchw(b.data_ptr(), c.data_ptr(), 2, 5, 7, 9, G.S()))
    assert torch.equal(c.cpu(), a)


def test_bad_arguments_raise(cuda):
    lib = L.load()
    out = torch.empty(4, device=cuda)
    assert lib.femasr_window_attention(out.data_ptr(), out.data_ptr(), out.data_ptr(), 1, 12, 8, 256, 8, 0, G.S()) == -1
    assert lib.femasr_ln_stats(out.data_ptr(), out.data_ptr(), out.data_ptr(), 4, 128, 1e-5, G.S()) == -1
    with pytest.raises(L.FemasrError):
        L.check(lib.femasr_flip_pad(out.data_ptr(), out.data_ptr(), 1, 1, 4, 4, 9, 4, G.S()))


@pytest.mark.parametrize("B,H,W,Ca,Hb,Wb,Cb", [(2, 8, 12, 256, 8, 12, 256), (1, 16, 8, 128, 8, 4, 256),
                                              (2, 16, 16, 64, 4, 4, 128), (1, 6, 9, 64, 3, 3, 64)])
def test_concat_channels_nearest(cuda, B, H, W, Ca, Hb, Wb, Cb):
    """torch.cat((a, F.interpolate(b, a.shape[2:])), 1): femasr_arch.py:332-335 (same size) and fema_utils.py:92-99."""
    a, b = rnd(B, Ca, H, W, seed=60), rnd(B, Cb, Hb, Wb, seed=61)
    want = torch.cat((a, F.interpolate(b, (H, W))), dim=1)
    out = torch.empty(B, H, W, Ca + Cb, device=cuda)
    ag, bg = G.nhwc(a).to(cuda), G.nhwc(b).to(cuda)
    L.check(L.load().femasr_concat_channels(G.p(ag), Ca, G.p(bg), Hb, Wb, Cb, G.p(out), B, H, W, G.S()))
    assert torch.equal(G.nchw(out).cpu(), want)


@pytest.mark.parametrize("B,h,w,e,n_e", [(2, 16, 16, 256, 1024), (1, 8, 12, 128, 512), (3, 5, 7, 64, 256)])
def test_gt_indices_loss_terms(cuda, B, h, w, e, n_e):
    """The supervised VQ loss (femasr_arch.py:70-78, 84-90) assembled from vq_gt_rows + gram_diff + sum_scaled(_add)."""
    lib = L.load()
    z = rnd(B, e, h, w, seed=62)
    cb = rnd(n_e, e, seed=63, scale=0.7)
    gt = torch.randint(0, n_e, (B, 1, h, w), generator=torch.Generator().manual_seed(64))
    _, want, _ = O.vector_quantize(cb, z, gt, lq=True)
    N = B * h * w
    zg, cbg, gtg = G.nhwc(z).to(cuda), cb.to(cuda), gt.to(cuda)
    zq_gt, rows = torch.empty(N, e, device=cuda), torch.empty(N, device=cuda)
    L.check(lib.femasr_vq_gt_rows(G.p(zg), G.p(cbg), G.p(gtg), G.p(zq_gt), G.p(rows), N, n_e, e, G.S()))
    assert torch.equal(zq_gt.cpu(), cb[gt.reshape(-1)])
    tiles = lib.femasr_gram_diff_tiles(e)
    assert tiles == (e // 32) ** 2
    part = torch.empty(B * tiles, device=cuda)
    L.check(lib.femasr_gram_diff(G.p(zg), G.p(zq_gt), G.p(part), B, h * w, e, G.S()))
    zf = z.permute(0, 2, 3, 1).reshape(B, h * w, e).double()
    yf = cb[gt.reshape(-1)].reshape(B, h * w, e).double()
    gram = ((zf.transpose(1, 2) @ zf - yf.transpose(1, 2) @ yf) / (h * w)).square().sum()
    assert abs(part.double().sum().item() - gram.item()) <= 2e-5 * gram.item()
    loss = torch.full((1,), 123.0, device=cuda)
    L.check(lib.femasr_sum_scaled(G.p(rows), G.p(loss), N, 0.25 / (N * e), G.S()))
    L.check(lib.femasr_sum_scaled_add(G.p(part), G.p(loss), B * tiles, 1.0 / (B * e * e), G.S()))
    assert abs(loss.item() - want.item()) <= 2e-5 * abs(want.item())


@pytest.mark.parametrize("scale,cbs,shape,path", [
    (4, [[32, 1024, 256], [64, 512, 128]], (2, 3, 48, 32), 0),
    (4, [[32, 1024, 256], [64, 512, 128]], (2, 3, 48, 32), 1),
    (2, [[32, 512, 256], [64, 512, 256], [128, 256, 128]], (1, 3, 64, 96), 1),
    (1, [[32, 1024, 256], [128, 256, 64]], (1, 3, 64, 128), 1),
    (4, [[32, 1024, 256], [128, 512, 64]], (1, 3, 32, 48), 1)])
def test_multiscale_codebooks_against_oracle(cuda, scale, cbs, shape, path):
    """Multi-scale codebooks (femasr_arch.py:280-299, 329-359) on fresh seeded inputs, both GEMM paths: every codebook's
    indices bit-exact, features in front of the later quantisers and the output within the fp32 bars."""
    from basicsr.archs.femasr_arch import FeMaSRNet
    from femasr_b200.spec import random_state_dict
    sd = random_state_dict(scale, cbs[0][2], seed=70, init="perturbed", codebooks=cbs)
    net = FeMaSRNet(codebook_params=cbs, LQ_stage=scale != 1, scale_factor=scale, gemm_path=path)
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda).eval()
    x = torch.rand(shape, generator=torch.Generator().manual_seed(71))
    gt = None
    if scale != 1:
        div = {4: 2, 2: 4}[scale]
        gt = [torch.randint(0, n, (shape[0], 1, shape[2] // div * s // 32, shape[3] // div * s // 32),
                            generator=torch.Generator().manual_seed(72 + k)) for k, (s, n, _e) in enumerate(cbs)]
    with torch.no_grad():
        want, wloss, _, widx = O.encode_and_decode(sd, x, scale, cb_scales=[c[0] for c in cbs])
        out, loss, sem, idx = net(x.to(cuda))
        eng = net._native(torch.device(cuda))
        names = [f"z{k}" for k in range(1, len(cbs))]
        taps = eng.forward(x.to(cuda), taps=names)[3]
    assert len(idx) == len(cbs)
    for k, (a, b) in enumerate(zip(idx, widx)):
        assert tuple(a.shape) == tuple(b.shape)
        mism = int((a.cpu() != b).sum())
        assert mism == 0, f"codebook {k}: {mism}/{b.numel()} index mismatches"
    assert abs(loss.item() - wloss.item()) <= (1e-4 if path else 2e-5) * abs(wloss.item())
    err = (out.cpu() - want).abs().max().item()
    print(f"multi-scale x{scale} {cbs} path {path}: output max-abs {err:.2e}")
    assert err <= 1e-3
    assert all(t.isfinite().all() for t in taps.values())
    if gt is not None:
        with torch.no_grad():
            _, wl2, _, _ = O.encode_and_decode(sd, x, scale, cb_scales=[c[0] for c in cbs], gt_indices=gt)
            out2, l2, _, idx2 = net(x.to(cuda), gt)
        assert torch.equal(out2, out) and all(torch.equal(a, b) for a, b in zip(idx, idx2)), "gt_indices only change the loss"
        assert abs(l2.item() - wl2.item()) <= 1e-4 * abs(wl2.item())
