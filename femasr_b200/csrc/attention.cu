// Shifted-window multi-head self-attention (8x8 windows, head_dim 32), fp32 SIMT.
// One CTA per (window, head); one thread per query token.  The cyclic shift (torch.roll), the window
// partition/reverse permutes and the 0/-100 shift mask of network_swinir.py:216-279 are pure index
// arithmetic here; nothing is materialised.  0.9% of the path's FLOPs.
#include "common.cuh"

namespace femasr {

constexpr int WS = 8, WT = 64, HD = 32;

__device__ __forceinline__ int shift_region(int p, int n, int shift) {
  // img_mask regions of calculate_mask(): [0,n-8) -> 0, [n-8,n-shift) -> 1, [n-shift,n) -> 2
  return p < n - WS ? 0 : (p < n - shift ? 1 : 2);
}

__global__ void __launch_bounds__(64) window_attention_kernel(const float* __restrict__ qkv,
                                                              const float* __restrict__ bias_full,
                                                              float* __restrict__ out, int H, int W, int C,
                                                              int heads, int shift) {
  __shared__ __align__(16) float ks[WT][HD];
  __shared__ __align__(16) float vs[WT][HD];
  __shared__ int region[WT];
  const int head = blockIdx.x % heads;
  const int win = blockIdx.x / heads;
  const int nwx = W / WS, nwy = H / WS;
  const int b = win / (nwx * nwy);
  const int wrem = win - b * nwx * nwy;
  const int wy = wrem / nwx, wx = wrem - wy * nwx;
  const int t = threadIdx.x;               // token within window
  const int ys = wy * WS + t / WS, xs = wx * WS + (t % WS);          // shifted-frame coordinates
  const int yo = (ys + shift) % H, xo = (xs + shift) % W;            // original coordinates (roll by -shift)
  const long tok = ((long)b * H + yo) * W + xo;
  const float* row = qkv + tok * (3 * C);
  const float scale = 0.17677669529663687f;   // head_dim ** -0.5 for head_dim 32 (network_swinir.py:84)

  float q[HD];
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row + head * HD) + i);
    q[4 * i] = v.x * scale; q[4 * i + 1] = v.y * scale; q[4 * i + 2] = v.z * scale; q[4 * i + 3] = v.w * scale;
  }
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) {
    reinterpret_cast<float4*>(ks[t])[i] = __ldg(reinterpret_cast<const float4*>(row + C + head * HD) + i);
    reinterpret_cast<float4*>(vs[t])[i] = __ldg(reinterpret_cast<const float4*>(row + 2 * C + head * HD) + i);
  }
  region[t] = shift > 0 ? shift_region(ys, H, shift) * 3 + shift_region(xs, W, shift) : 0;
  __syncthreads();

  const float* brow = bias_full + ((long)head * WT + t) * WT;
  const int myreg = region[t];
  float s[WT];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < WT; ++j) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < HD; ++k) a = fmaf(q[k], ks[j][k], a);
    a += __ldg(brow + j);
    if (shift > 0 && region[j] != myreg) a += -100.0f;
    s[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < WT; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
  const float inv = 1.0f / sum;
  float o[HD];
#pragma unroll
  for (int k = 0; k < HD; ++k) o[k] = 0.f;
#pragma unroll
  for (int j = 0; j < WT; ++j) {
    const float pj = s[j] * inv;
#pragma unroll
    for (int k = 0; k < HD; ++k) o[k] = fmaf(pj, vs[j][k], o[k]);
  }
  float* orow = out + tok * C + head * HD;
#pragma unroll
  for (int i = 0; i < HD / 4; ++i)
    reinterpret_cast<float4*>(orow)[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
}

// bias_full[h][i][j] = table[rel_index(i,j)][h],  rel_index = (yi-yj+7)*15 + (xi-xj+7)   (network_swinir.py:91-101,127-129)
__global__ void expand_rel_bias_kernel(const float* __restrict__ table, float* __restrict__ bias_full, int heads) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= heads * WT * WT) return;
  const int j = idx % WT, i = (idx / WT) % WT, h = idx / (WT * WT);
  const int dy = i / WS - j / WS + WS - 1, dx = i % WS - j % WS + WS - 1;
  bias_full[idx] = table[(dy * (2 * WS - 1) + dx) * heads + h];
}

}  // namespace femasr

using namespace femasr;

extern "C" int femasr_window_attention(const float* qkv, const float* bias_full, float* out, int B, int H, int W,
                                       int C, int heads, int shift, void* stream) {
  FEMASR_CHECK_ARG(qkv && bias_full && out, "window_attention: null pointer");
  FEMASR_CHECK_ARG(B > 0 && H > 0 && W > 0, "window_attention: empty input");
  FEMASR_CHECK_ARG(H % WS == 0 && W % WS == 0, "window_attention: H and W must be multiples of the 8x8 window");
  FEMASR_CHECK_ARG(heads > 0 && C == heads * HD, "window_attention: C must equal heads*32");
  FEMASR_CHECK_ARG(shift == 0 || shift == WS / 2, "window_attention: shift must be 0 or 4");
  // note: the reference fixes shift_size from the constructor's input_resolution (32,32), not from the
  // runtime map size (network_swinir.py:190-193), so a one-window-high map is still shifted and masked.
  const long blocks = (long)B * (H / WS) * (W / WS) * heads;
  window_attention_kernel<<<(unsigned)blocks, 64, 0, as_stream(stream)>>>(qkv, bias_full, out, H, W, C, heads, shift);
  return launch_status("window_attention_kernel");
}

extern "C" int femasr_expand_rel_bias(const float* table, float* bias_full, int heads, void* stream) {
  FEMASR_CHECK_ARG(table && bias_full && heads > 0, "expand_rel_bias: bad argument");
  const int n = heads * WT * WT;
  expand_rel_bias_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(table, bias_full, heads);
  return launch_status("expand_rel_bias_kernel");
}
