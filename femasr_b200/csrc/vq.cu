// VectorQuantizer feature matching (femasr_arch.py:50-100): fp32-exact distance formula + argmin +
// codebook gather + straight-through residual.  The z.e^T products come from the GEMM path; this file
// reproduces the reference's rounding sequence d_j = fl(fl(A + B_j) - 2 C_j) and its tie rule
// (lowest index wins), which is what makes the indices bit-exact.
#include "common.cuh"

namespace femasr {

// out[r] = sum_k x[r][k]^2   (torch.sum(y**2, dim=1), femasr_arch.py:37)
__global__ void row_sumsq_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int cols) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int k = lane; k < cols; k += 32) { const float v = x[(long)r * cols + k]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  if (lane == 0) out[r] = s;
}

// one warp per feature row
__global__ void __launch_bounds__(256) vq_select_kernel(const float* __restrict__ z, const float* __restrict__ zc,
                                                        const float* __restrict__ codebook,
                                                        const float* __restrict__ esq, int64_t* __restrict__ idx,
                                                        float* __restrict__ zq, float* __restrict__ loss_rows,
                                                        int N, int n_e, int e_dim) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= N) return;
  const int lane = threadIdx.x & 31;
  const float* zr = z + (long)r * e_dim;
  float a = 0.f;
  for (int k = lane; k < e_dim; k += 32) { const float v = zr[k]; a = fmaf(v, v, a); }
  a = warp_sum(a);                                     // A = sum z^2 (all lanes hold the same value)
  const float* cr = zc + (long)r * n_e;
  float best = INFINITY;
  int bj = 0x7fffffff;
  for (int j = lane; j < n_e; j += 32) {
    const float ab = __fadd_rn(a, __ldg(esq + j));     // fl(A + B_j)
    const float d = __fsub_rn(ab, __fmul_rn(2.0f, cr[j]));   // fl(. - 2 C_j); 2*C_j is exact
    if (d < best) { best = d; bj = j; }                // strict <: the lowest j of this lane wins
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (od < best || (od == best && oj < bj)) { best = od; bj = oj; }
  }
  if (bj == 0x7fffffff) bj = 0;                        // all-NaN row: torch.argmin would also return a valid index
  if (lane == 0 && idx) idx[r] = (int64_t)bj;
  const float* er = codebook + (long)bj * e_dim;
  float l = 0.f;
  for (int k = lane; k < e_dim; k += 32) {
    const float zv = zr[k];
    const float diff = __fsub_rn(__ldg(er + k), zv);   // (z_q - z)
    l = fmaf(diff, diff, l);
    if (zq) zq[(long)r * e_dim + k] = __fadd_rn(zv, diff);   // z + (z_q - z).detach(), femasr_arch.py:95
  }
  l = warp_sum(l);
  if (lane == 0 && loss_rows) loss_rows[r] = l;
}

__global__ void __launch_bounds__(1024) sum_scaled_kernel(const float* __restrict__ x, float* __restrict__ out, size_t n,
                                                          double scale) {
  __shared__ double red[32];
  double s = 0.0;
  for (size_t i = threadIdx.x; i < n; i += 1024) s += (double)x[i];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = warp_sum_d(red[threadIdx.x]);
    if (threadIdx.x == 0) out[0] = (float)(s * scale);
  }
}

__global__ void codebook_gather_kernel(const int64_t* __restrict__ idx, const float* __restrict__ codebook,
                                       float* __restrict__ zq, int N, int n_e, int e_dim) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= N) return;
  long j = idx[r];
  j = j < 0 ? 0 : (j >= n_e ? n_e - 1 : j);
  for (int k = threadIdx.x & 31; k < e_dim; k += 32) zq[(long)r * e_dim + k] = __ldg(codebook + j * e_dim + k);
}

}  // namespace femasr

using namespace femasr;

extern "C" int femasr_row_sumsq(const float* x, float* out, int rows, int cols, void* stream) {
  FEMASR_CHECK_ARG(x && out && rows > 0 && cols > 0, "row_sumsq: bad argument");
  row_sumsq_kernel<<<(rows + 7) / 8, 256, 0, as_stream(stream)>>>(x, out, rows, cols);
  return launch_status("row_sumsq_kernel");
}

extern "C" int femasr_vq_select(const float* z, const float* zc, const float* codebook, const float* esq,
                                int64_t* idx, float* zq, float* loss_rows, int N, int n_e, int e_dim,
                                int write_zq_passthrough, void* stream) {
  (void)write_zq_passthrough;
  FEMASR_CHECK_ARG(z && zc && codebook && esq, "vq_select: null pointer");
  FEMASR_CHECK_ARG(N > 0 && n_e > 0 && e_dim > 0, "vq_select: empty input");
  vq_select_kernel<<<(N + 7) / 8, 256, 0, as_stream(stream)>>>(z, zc, codebook, esq, idx, zq, loss_rows, N, n_e, e_dim);
  return launch_status("vq_select_kernel");
}

extern "C" int femasr_sum_scaled(const float* x, float* out, size_t n, double scale, void* stream) {
  FEMASR_CHECK_ARG(x && out && n > 0, "sum_scaled: bad argument");
  sum_scaled_kernel<<<1, 1024, 0, as_stream(stream)>>>(x, out, n, scale);
  return launch_status("sum_scaled_kernel");
}

extern "C" int femasr_codebook_gather(const int64_t* idx, const float* codebook, float* zq, int N, int n_e, int e_dim,
                                      void* stream) {
  FEMASR_CHECK_ARG(idx && codebook && zq && N > 0, "codebook_gather: bad argument");
  codebook_gather_kernel<<<(N + 7) / 8, 256, 0, as_stream(stream)>>>(idx, codebook, zq, N, n_e, e_dim);
  return launch_status("codebook_gather_kernel");
}
