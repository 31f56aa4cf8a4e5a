// The skinny ends of the path and the layout / padding helpers (all HBM-bound, SIMT):
//   in_conv  4x4 p1, Cin=3   NCHW -> NHWC   (femasr_arch.py:150)
//   out_conv 3x3 p1, Cout=3  NHWC -> NCHW   (femasr_arch.py:273)
//   weight repack, NCHW<->NHWC, flip-pad (test(), :459-460), window copy (crop / tile paste)
#include <cuda_fp16.h>

#include "common.cuh"

namespace femasr {

// in_conv: 4x4, pad 1, Cin = 3.  Thread = (4 output channels) x (4 consecutive output pixels of one row): the
// 4x7x3 input patch lives in registers (loads are shared by all channel-quad threads of the pixel group through
// L1), every weight float4 is used for 16 FMAs, and the 4 stores per thread are contiguous across the warp's
// channel quads.  SPLIT: write the fp16 hi/lo operand planes of the following tensor-core conv instead of fp32.
template <bool SPLIT>
__global__ void __launch_bounds__(256) in_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y,
                                                      __half* __restrict__ yhi, __half* __restrict__ ylo,
                                                      int B, int H, int W, int Cout) {
  constexpr int KS = 4, CIN = 3, PX = 4;
  const int Ho = H - 1, Wo = W - 1;
  const int quads = Cout / 4;
  const int gpb = 256 / quads;                         // pixel groups per block
  const int q = threadIdx.x % quads;
  const int groups_x = (Wo + PX - 1) / PX;
  const long grp = (long)blockIdx.x * gpb + threadIdx.x / quads;
  const long ngrp = (long)B * Ho * groups_x;
  if (grp >= ngrp) return;
  const int gx = (int)(grp % groups_x);
  const long t = grp / groups_x;
  const int oy = (int)(t % Ho), b = (int)(t / Ho);
  const int ox0 = gx * PX;
  float patch[CIN][KS][PX + KS - 1];
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
    for (int r = 0; r < KS; ++r) {
      const int iy = oy + r - 1;
#pragma unroll
      for (int cidx = 0; cidx < PX + KS - 1; ++cidx) {
        const int ix = ox0 + cidx - 1;
        patch[ci][r][cidx] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(x + (((long)b * CIN + ci) * H + iy) * W + ix) : 0.f;
      }
    }
  float4 acc[PX];
  const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int p = 0; p < PX; ++p) acc[p] = bv;
#pragma unroll
  for (int kh = 0; kh < KS; ++kh)
#pragma unroll
    for (int kw = 0; kw < KS; ++kw)
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + ((long)(kh * KS + kw) * CIN + ci) * Cout) + q);
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const float xv = patch[ci][kh][p + kw];
          acc[p].x = fmaf(xv, wv.x, acc[p].x); acc[p].y = fmaf(xv, wv.y, acc[p].y);
          acc[p].z = fmaf(xv, wv.z, acc[p].z); acc[p].w = fmaf(xv, wv.w, acc[p].w);
        }
      }
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int ox = ox0 + p;
    if (ox >= Wo) break;
    const long e = ((((long)b * Ho + oy) * Wo + ox) * Cout) + q * 4;
    if (SPLIT) {
      const float v[4] = {acc[p].x, acc[p].y, acc[p].z, acc[p].w};
      __align__(8) __half h[4], l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float cl = fminf(fmaxf(v[k], -65504.f), 65504.f);
        h[k] = __float2half_rn(cl);
        l[k] = __float2half_rn(cl - __half2float(h[k]));
      }
      *reinterpret_cast<uint2*>(yhi + e) = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(ylo + e) = *reinterpret_cast<const uint2*>(l);
    } else {
      *reinterpret_cast<float4*>(y + e) = acc[p];
    }
  }
}

// out_conv: 3x3, Cin=64 -> 3.  The input tile (with halo) is staged in shared memory by coalesced 16-byte
// loads (NHWC rows are contiguous), pixel stride padded to 68 floats so the per-thread float4 reads are
// bank-conflict free; the 1728 weights sit in __constant__ memory (uniform across the warp -> FFMA with a
// constant operand, no load instruction).  One thread per output pixel, tile = 4 rows x 32 cols.
constexpr int OC_CIN = 64, OC_TH = 4, OC_TW = 32, OC_PS = 68;   // pixel stride in floats
__constant__ float c_outconv_w[9 * OC_CIN * 3];
__constant__ float c_outconv_b[4];

__global__ void __launch_bounds__(OC_TH * OC_TW) out_conv_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                 int B, int H, int W) {
  extern __shared__ __align__(16) float tile[];      // [(OC_TH+2)][(OC_TW+2)][OC_PS]
  const int x0 = blockIdx.x * OC_TW, y0 = blockIdx.y * OC_TH, b = blockIdx.z;
  constexpr int TWH = OC_TW + 2, THH = OC_TH + 2;
  // cooperative halo load: (THH*TWH) pixels x 16 float4
  for (int i = threadIdx.x; i < THH * TWH * (OC_CIN / 4); i += OC_TH * OC_TW) {
    const int c4 = i % (OC_CIN / 4);
    const int pp = i / (OC_CIN / 4);
    const int px = pp % TWH, py = pp / TWH;
    const int gy = y0 + py - 1, gx = x0 + px - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W)
      v = __ldg(reinterpret_cast<const float4*>(x + (((long)b * H + gy) * W + gx) * OC_CIN) + c4);
    *reinterpret_cast<float4*>(&tile[(py * TWH + px) * OC_PS + c4 * 4]) = v;
  }
  __syncthreads();
  const int lx = threadIdx.x % OC_TW, ly = threadIdx.x / OC_TW;
  const int ox = x0 + lx, oy = y0 + ly;
  float a0 = c_outconv_b[0], a1 = c_outconv_b[1], a2 = c_outconv_b[2];
#pragma unroll 1
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const float* px = &tile[((ly + kh) * TWH + lx + kw) * OC_PS];
      const float* wt = c_outconv_w + (kh * 3 + kw) * OC_CIN * 3;
#pragma unroll 4
      for (int c4 = 0; c4 < OC_CIN / 4; ++c4) {
        const float4 v = *reinterpret_cast<const float4*>(px + c4 * 4);
        const float* wk = wt + c4 * 12;
        a0 = fmaf(v.x, wk[0], a0); a1 = fmaf(v.x, wk[1], a1); a2 = fmaf(v.x, wk[2], a2);
        a0 = fmaf(v.y, wk[3], a0); a1 = fmaf(v.y, wk[4], a1); a2 = fmaf(v.y, wk[5], a2);
        a0 = fmaf(v.z, wk[6], a0); a1 = fmaf(v.z, wk[7], a1); a2 = fmaf(v.z, wk[8], a2);
        a0 = fmaf(v.w, wk[9], a0); a1 = fmaf(v.w, wk[10], a1); a2 = fmaf(v.w, wk[11], a2);
      }
    }
  if (ox < W && oy < H) {
    const long plane = (long)H * W;
    float* o = y + (long)b * 3 * plane + (long)oy * W + ox;
    o[0] = a0; o[plane] = a1; o[2 * plane] = a2;
  }
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

// uint8 image boundary (inference_femasr.py:54-56,64 + basicsr/utils/img_util.py:9-35,38-94):
//   in : uint8 HWC BGR [B,h,w,3]  -> fp32 NCHW RGB /255, flip-padded to [B,3,hp,wp]  (img2tensor, /255., test() padding)
//   out: fp32 NCHW RGB [B,3,SH,SW] -> clamp to [0,1], *255, round-half-even, uint8 HWC BGR, cropped to [B,ch,cw,3]
__global__ void u8_to_input_kernel(const uint8_t* __restrict__ img, float* __restrict__ x, int h, int w, int hp, int wp, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;     // index into [B,3,hp,wp]
  if (i >= n) return;
  const int j = (int)(i % wp);
  const int r = (int)((i / wp) % hp);
  const int c = (int)((i / wp / hp) % 3);
  const long b = i / wp / hp / 3;
  const int sr = r < h ? r : 2 * h - 1 - r, sj = j < w ? j : 2 * w - 1 - j;
  x[i] = (float)img[((b * h + sr) * w + sj) * 3 + (2 - c)] / 255.0f;   // torch: float32(img) / 255.
}
__global__ void output_to_u8_kernel(const float* __restrict__ y, uint8_t* __restrict__ img, int SH, int SW, int ch, int cw, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;     // index into [B,ch,cw,3]
  if (i >= n) return;
  const int c = (int)(i % 3);
  const int j = (int)((i / 3) % cw);
  const int r = (int)((i / 3 / cw) % ch);
  const long b = i / 3 / cw / ch;
  float v = y[((b * 3 + (2 - c)) * SH + r) * SW + j];
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  img[i] = (uint8_t)rintf(v * 255.0f);                              // numpy round(): half to even
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

static int in_conv_launch(const float* x, const float* w, const float* bias, float* y, void* yhi, void* ylo, int B,
                          int Cin, int H, int W, int Cout, void* stream) {
  FEMASR_CHECK_ARG(x && w && (y || (yhi && ylo)), "in_conv: null pointer");
  FEMASR_CHECK_ARG(B > 0 && H >= 3 && W >= 3 && Cin == 3, "in_conv: needs Cin == 3 and H, W >= 3");
  FEMASR_CHECK_ARG(Cout % 4 == 0 && 256 % (Cout / 4) == 0 && Cout <= 1024, "in_conv: unsupported Cout");
  const long ngrp = (long)B * (H - 1) * cdiv(W - 1, 4);
  const int gpb = 256 / (Cout / 4);
  const unsigned grid = (unsigned)cdiv(ngrp, gpb);
  if (y)
    in_conv_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(x, w, bias, y, nullptr, nullptr, B, H, W, Cout);
  else
    in_conv_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(x, w, bias, nullptr, reinterpret_cast<__half*>(yhi),
                                                                reinterpret_cast<__half*>(ylo), B, H, W, Cout);
  return launch_status("in_conv_kernel");
}

extern "C" int femasr_in_conv4x4(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int H,
                                 int W, int Cout, void* stream) {
  FEMASR_CHECK_ARG(y, "in_conv: null output");
  return in_conv_launch(x, w, bias, y, nullptr, nullptr, B, Cin, H, W, Cout, stream);
}

extern "C" int femasr_in_conv4x4_split(const float* x, const float* w, const float* bias, void* y_hi, void* y_lo, int B,
                                       int Cin, int H, int W, int Cout, void* stream) {
  FEMASR_CHECK_ARG(y_hi && y_lo, "in_conv_split: null output");
  return in_conv_launch(x, w, bias, nullptr, y_hi, y_lo, B, Cin, H, W, Cout, stream);
}

extern "C" int femasr_out_conv3x3(const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                                  int Cin, void* stream) {
  FEMASR_CHECK_ARG(x && w && bias && y, "out_conv: null pointer");
  FEMASR_CHECK_ARG(B > 0 && H > 0 && W > 0, "out_conv: empty input");
  FEMASR_CHECK_ARG(Cin == OC_CIN, "out_conv: Cin must be 64 (channel_query_dict[256])");
  FEMASR_CHECK_ARG(cdiv(H, OC_TH) <= 65535 && B <= 65535, "out_conv: grid too large");
  cudaStream_t st = as_stream(stream);
  // weights travel through __constant__ memory; refreshed per call (stream-ordered) so several engines can coexist
  FEMASR_CUDA(cudaMemcpyToSymbolAsync(c_outconv_w, w, sizeof(float) * 9 * OC_CIN * 3, 0, cudaMemcpyDeviceToDevice, st));
  FEMASR_CUDA(cudaMemcpyToSymbolAsync(c_outconv_b, bias, sizeof(float) * 3, 0, cudaMemcpyDeviceToDevice, st));
  constexpr int smem = (OC_TH + 2) * (OC_TW + 2) * OC_PS * (int)sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    FEMASR_CUDA(cudaFuncSetAttribute(out_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((unsigned)cdiv(W, OC_TW), (unsigned)cdiv(H, OC_TH), B);
  out_conv_kernel<<<grid, OC_TH * OC_TW, smem, st>>>(x, y, B, H, W);
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

extern "C" int femasr_u8_to_input(const uint8_t* bgr_hwc, float* x_nchw, int B, int h, int w, int hp, int wp, void* stream) {
  FEMASR_CHECK_ARG(bgr_hwc && x_nchw && B > 0 && h > 0 && w > 0, "u8_to_input: bad argument");
  FEMASR_CHECK_ARG(hp >= h && wp >= w && hp <= 2 * h && wp <= 2 * w, "u8_to_input: pad must be within one reflection");
  const long n = (long)B * 3 * hp * wp;
  u8_to_input_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(bgr_hwc, x_nchw, h, w, hp, wp, n);
  return launch_status("u8_to_input_kernel");
}

extern "C" int femasr_output_to_u8(const float* y_nchw, uint8_t* bgr_hwc, int B, int SH, int SW, int ch, int cw, void* stream) {
  FEMASR_CHECK_ARG(y_nchw && bgr_hwc && B > 0 && ch > 0 && cw > 0 && ch <= SH && cw <= SW, "output_to_u8: bad argument");
  const long n = (long)B * ch * cw * 3;
  output_to_u8_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(y_nchw, bgr_hwc, SH, SW, ch, cw, n);
  return launch_status("output_to_u8_kernel");
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
