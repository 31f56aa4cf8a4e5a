// The skinny ends of the path and the layout / padding helpers (all HBM-bound, SIMT):
//   in_conv  4x4 p1, Cin=3   NCHW -> NHWC   (femasr_arch.py:150)
//   out_conv 3x3 p1, Cout=3  NHWC -> NCHW   (femasr_arch.py:273)
//   weight repack, NCHW<->NHWC, flip-pad (test(), :459-460), window copy (crop / tile paste)
#include "common.cuh"

namespace femasr {

template <int KS>
__global__ void __launch_bounds__(256) in_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y,
                                                      int B, int Cin, int H, int W, int Cout) {
  const int Ho = H + 2 - KS + 1, Wo = W + 2 - KS + 1;
  const int quads = Cout / 4;
  const int ppb = 256 / quads;                       // pixels per block
  const int q = threadIdx.x % quads;
  const long pix = (long)blockIdx.x * ppb + threadIdx.x / quads;
  const long npix = (long)B * Ho * Wo;
  if (pix >= npix) return;
  const int b = (int)(pix / (Ho * Wo));
  const int r = (int)(pix - (long)b * Ho * Wo);
  const int oy = r / Wo, ox = r - oy * Wo;
  float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int kh = 0; kh < KS; ++kh) {
    const int iy = oy + kh - 1;
    if (iy < 0 || iy >= H) continue;
    for (int kw = 0; kw < KS; ++kw) {
      const int ix = ox + kw - 1;
      if (ix < 0 || ix >= W) continue;
      for (int ci = 0; ci < Cin; ++ci) {
        const float xv = __ldg(x + (((long)b * Cin + ci) * H + iy) * W + ix);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + ((long)(kh * KS + kw) * Cin + ci) * Cout) + q);
        acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y);
        acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
      }
    }
  }
  reinterpret_cast<float4*>(y + pix * Cout)[q] = acc;
}

// one thread per output pixel; weights [9*Cin][3] in shared memory (broadcast reads)
template <int CIN>
__global__ void __launch_bounds__(128) out_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ y,
                                                       int B, int H, int W) {
  __shared__ __align__(16) float ws[9 * CIN * 3 + 4];
  for (int i = threadIdx.x; i < 9 * CIN * 3; i += 128) ws[i] = w[i];
  __syncthreads();
  const int ox = blockIdx.x * 128 + threadIdx.x;
  const int oy = blockIdx.y, b = blockIdx.z;
  if (ox >= W) return;
  float a0 = bias[0], a1 = bias[1], a2 = bias[2];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int iy = oy + kh - 1;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ix = ox + kw - 1;
      if (ix < 0 || ix >= W) continue;
      const float4* px = reinterpret_cast<const float4*>(x + (((long)b * H + iy) * W + ix) * CIN);
      const float* wt = ws + (kh * 3 + kw) * CIN * 3;
#pragma unroll 4
      for (int c4 = 0; c4 < CIN / 4; ++c4) {
        const float4 v = __ldg(px + c4);
        const float* wk = wt + c4 * 12;
        a0 = fmaf(v.x, wk[0], a0); a1 = fmaf(v.x, wk[1], a1); a2 = fmaf(v.x, wk[2], a2);
        a0 = fmaf(v.y, wk[3], a0); a1 = fmaf(v.y, wk[4], a1); a2 = fmaf(v.y, wk[5], a2);
        a0 = fmaf(v.z, wk[6], a0); a1 = fmaf(v.z, wk[7], a1); a2 = fmaf(v.z, wk[8], a2);
        a0 = fmaf(v.w, wk[9], a0); a1 = fmaf(v.w, wk[10], a1); a2 = fmaf(v.w, wk[11], a2);
      }
    }
  }
  const long plane = (long)H * W;
  float* o = y + (long)b * 3 * plane + (long)oy * W + ox;
  o[0] = a0; o[plane] = a1; o[2 * plane] = a2;
}

__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int KH, int KW) {
  const long n = (long)Cout * Cin * KH * KW;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // i indexes the packed layout [(kh*KW+kw)*Cin + ci][co]
  const int co = (int)(i % Cout);
  long k = i / Cout;
  const int ci = (int)(k % Cin);
  const int tap = (int)(k / Cin);
  const int kh = tap / KW, kw = tap - kh * KW;
  out[i] = w[(((long)co * Cin + ci) * KH + kh) * KW + kw];
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, long HW, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // NHWC index
  if (i >= n) return;
  const int c = (int)(i % C);
  const long p = (i / C) % HW, b = i / C / HW;
  y[i] = x[(b * C + c) * HW + p];
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int C, long HW, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // NCHW index
  if (i >= n) return;
  const long p = i % HW;
  const int c = (int)((i / HW) % C);
  const long b = i / HW / C;
  y[i] = x[(b * HW + p) * C + c];
}

__global__ void flip_pad_kernel(const float* __restrict__ x, float* __restrict__ y, int h, int w, int hp, int wp, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = (int)(i % wp);
  const int r = (int)((i / wp) % hp);
  const long bc = i / wp / hp;
  const int sr = r < h ? r : 2 * h - 1 - r;          // cat([x, flip(x)])[:h+pad]
  const int sj = j < w ? j : 2 * w - 1 - j;
  y[i] = x[(bc * h + sr) * w + sj];
}

__global__ void copy_window_kernel(const float* __restrict__ src, float* __restrict__ dst, int sh, int sw, int dh, int dw,
                                   int sy, int sx, int dy, int dx, int ch, int cw, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = (int)(i % cw);
  const int r = (int)((i / cw) % ch);
  const long bc = i / cw / ch;
  dst[(bc * dh + dy + r) * dw + dx + j] = src[(bc * sh + sy + r) * sw + sx + j];
}

}  // namespace femasr

using namespace femasr;

extern "C" int femasr_in_conv4x4(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int H,
                                 int W, int Cout, void* stream) {
  FEMASR_CHECK_ARG(x && w && y, "in_conv: null pointer");
  FEMASR_CHECK_ARG(B > 0 && H >= 3 && W >= 3 && Cin > 0, "in_conv: input too small");
  FEMASR_CHECK_ARG(Cout % 4 == 0 && 256 % (Cout / 4) == 0 && Cout <= 1024, "in_conv: unsupported Cout");
  const long npix = (long)B * (H - 1) * (W - 1);
  const int ppb = 256 / (Cout / 4);
  in_conv_kernel<4><<<(unsigned)cdiv(npix, ppb), 256, 0, as_stream(stream)>>>(x, w, bias, y, B, Cin, H, W, Cout);
  return launch_status("in_conv_kernel");
}

extern "C" int femasr_out_conv3x3(const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                                  int Cin, void* stream) {
  FEMASR_CHECK_ARG(x && w && bias && y, "out_conv: null pointer");
  FEMASR_CHECK_ARG(B > 0 && H > 0 && W > 0, "out_conv: empty input");
  FEMASR_CHECK_ARG(Cin == 64, "out_conv: Cin must be 64 (channel_query_dict[256])");
  FEMASR_CHECK_ARG(H <= 65535 && B <= 65535, "out_conv: grid too large");
  dim3 grid((unsigned)cdiv(W, 128), H, B);
  out_conv_kernel<64><<<grid, 128, 0, as_stream(stream)>>>(x, w, bias, y, B, H, W);
  return launch_status("out_conv_kernel");
}

extern "C" int femasr_pack_weight(const float* w, float* out, int Cout, int Cin, int kh, int kw, void* stream) {
  FEMASR_CHECK_ARG(w && out && Cout > 0 && Cin > 0 && kh > 0 && kw > 0, "pack_weight: bad argument");
  const long n = (long)Cout * Cin * kh * kw;
  pack_weight_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(w, out, Cout, Cin, kh, kw);
  return launch_status("pack_weight_kernel");
}

extern "C" int femasr_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, void* stream) {
  FEMASR_CHECK_ARG(x && y && B > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad argument");
  const long n = (long)B * C * H * W;
  nchw_to_nhwc_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(x, y, C, (long)H * W, n);
  return launch_status("nchw_to_nhwc_kernel");
}
extern "C" int femasr_nhwc_to_nchw(const float* x, float* y, int B, int C, int H, int W, void* stream) {
  FEMASR_CHECK_ARG(x && y && B > 0 && C > 0 && H > 0 && W > 0, "nhwc_to_nchw: bad argument");
  const long n = (long)B * C * H * W;
  nhwc_to_nchw_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(x, y, C, (long)H * W, n);
  return launch_status("nhwc_to_nchw_kernel");
}

extern "C" int femasr_flip_pad(const float* x, float* y, int B, int C, int h, int w, int hp, int wp, void* stream) {
  FEMASR_CHECK_ARG(x && y && B > 0 && C > 0 && h > 0 && w > 0, "flip_pad: bad argument");
  FEMASR_CHECK_ARG(hp >= h && wp >= w && hp <= 2 * h && wp <= 2 * w, "flip_pad: pad must be within one reflection");
  const long n = (long)B * C * hp * wp;
  flip_pad_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(x, y, h, w, hp, wp, n);
  return launch_status("flip_pad_kernel");
}

extern "C" int femasr_copy_window(const float* src, float* dst, int B, int C, int sh, int sw, int dh, int dw, int sy,
                                  int sx, int dy, int dx, int ch, int cw, void* stream) {
  FEMASR_CHECK_ARG(src && dst && B > 0 && C > 0, "copy_window: bad argument");
  FEMASR_CHECK_ARG(ch > 0 && cw > 0 && sy >= 0 && sx >= 0 && dy >= 0 && dx >= 0 && sy + ch <= sh && sx + cw <= sw &&
                       dy + ch <= dh && dx + cw <= dw, "copy_window: window out of bounds");
  const long n = (long)B * C * ch * cw;
  copy_window_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(src, dst, sh, sw, dh, dw, sy, sx, dy, dx, ch, cw, n);
  return launch_status("copy_window_kernel");
}
