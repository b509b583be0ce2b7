// ☆ f4: low-rank multimodal fusion block (LMF, code/model_fusion.py:214-310; att_type='lmf_only').
//   h_m = Linear_m(x_m)                                   (mmdfn_gemm, by the caller: the three 300 -> 300 sub-networks)
//   f[m][r] = h_m . factor_m[r, 1:, :] + factor_m[r, 0, :]   (the reference appends a constant 1 to h: row 0 is a bias)
//   out = sum_r w_r f[0][r] * f[1][r] * f[2][r] + bias
// The rank products are plain GEMMs against contiguous (H, O) slices of the (rank, H + 1, O) factors (mmdfn_gemm with the
// slice's row 0 as bias); this file holds the orchestration and the elementwise triple product with its backward.
#include "internal.cuh"
#include "../../include/mmdfn_b200.h"

namespace mmdfn {

constexpr int LMF_MAXR = 8;

// fz: (3, R, n) with n = N * O flattened; out (n), bias indexed by (idx % O)
__global__ void lmf_combine_fwd_kernel(i64 n, int O, int R, const float* __restrict__ fz, const float* __restrict__ w,
                                       const float* __restrict__ bias, float* __restrict__ out) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = bias[i % O];
  for (int r = 0; r < R; r++) acc = fmaf(w[r], fz[(i64)r * n + i] * fz[(i64)(R + r) * n + i] * fz[(i64)(2 * R + r) * n + i], acc);
  out[i] = acc;
}

// dfz[m][r] = w_r dout prod_{m' != m} f[m'][r];  dw_r += sum dout f0 f1 f2 (block reduction + one atomic per block and rank)
__global__ void lmf_combine_bwd_kernel(i64 n, int R, const float* __restrict__ fz, const float* __restrict__ w,
                                       const float* __restrict__ dout, float* __restrict__ dfz, float* __restrict__ dw) {
  __shared__ float red[LMF_MAXR][8];
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  float part[LMF_MAXR];
#pragma unroll
  for (int r = 0; r < LMF_MAXR; r++) part[r] = 0.f;
  if (i < n) {
    const float g = dout[i];
#pragma unroll
    for (int r = 0; r < LMF_MAXR; r++) {
      if (r < R) {
        const float a = fz[(i64)r * n + i], v = fz[(i64)(R + r) * n + i], t = fz[(i64)(2 * R + r) * n + i];
        const float gw = g * w[r];
        dfz[(i64)r * n + i] = gw * v * t;
        dfz[(i64)(R + r) * n + i] = gw * a * t;
        dfz[(i64)(2 * R + r) * n + i] = gw * a * v;
        part[r] = g * a * v * t;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < LMF_MAXR; r++) {
    if (r < R) {
      const float s = warp_sum(part[r]);
      if ((threadIdx.x & 31) == 0) red[r][threadIdx.x >> 5] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < R) {
    float s = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); k++) s += red[threadIdx.x][k];
    atomicAdd(dw + threadIdx.x, s);
  }
}

}  // namespace mmdfn

using namespace mmdfn;

extern "C" long long mmdfn_lmf_ws_floats(int N, int R, int O) { return (long long)3 * R * N * O; }

/* h: three (N, H) hidden activations {audio, video, text}; factor: three (R, H + 1, O) tensors; w (R); bias (O).
   fz: workspace / saved for backward, 3 R N O floats; out (N, O). */
extern "C" int mmdfn_lmf_fuse_fwd(int N, int H, int O, int R, const float* const* h, const float* const* factor, const float* w,
                                  const float* bias, float* fz, float* out, void* stream) {
  if (!h || !factor || !w || !bias || !fz || !out) return MMDFN_ENULL;
  if (N < 0 || H <= 0 || O <= 0 || R <= 0 || R > LMF_MAXR) return MMDFN_EINVAL;
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const i64 n = (i64)N * O;
  for (int m = 0; m < 3; m++)
    for (int r = 0; r < R; r++) {
      const float* Fr = factor[m] + (i64)r * (H + 1) * O;
      MMDFN_TRY(gemm(false, false, N, O, H, 1.f, h[m], H, Fr + O, O, 0.f, fz + (i64)(m * R + r) * n, O, Fr, 0, st));
    }
  lmf_combine_fwd_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(n, O, R, fz, w, bias, out);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

/* dfz: workspace, 3 R N O floats.  Overwrites dh[m] (N, H), dfactor[m] (R, H + 1, O), dbias (O); dw (R) is accumulated into
   (zero it first). */
extern "C" int mmdfn_lmf_fuse_bwd(int N, int H, int O, int R, const float* const* h, const float* const* factor, const float* w,
                                  const float* fz, const float* dout, float* dfz, float* const* dh, float* const* dfactor,
                                  float* dw, float* dbias, void* stream) {
  if (!h || !factor || !w || !fz || !dout || !dfz || !dh || !dfactor || !dw || !dbias) return MMDFN_ENULL;
  if (N < 0 || H <= 0 || O <= 0 || R <= 0 || R > LMF_MAXR) return MMDFN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    for (int m = 0; m < 3; m++) MMDFN_TRY(fill_zero(dfactor[m], (size_t)R * (H + 1) * O * sizeof(float), st));
    return fill_zero(dbias, (size_t)O * sizeof(float), st);
  }
  const i64 n = (i64)N * O;
  lmf_combine_bwd_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(n, R, fz, w, dout, dfz, dw);
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(colsum(N, O, dout, O, 0.f, dbias, st));
  for (int m = 0; m < 3; m++)
    for (int r = 0; r < R; r++) {
      const float* Fr = factor[m] + (i64)r * (H + 1) * O;
      float* dFr = dfactor[m] + (i64)r * (H + 1) * O;
      const float* g = dfz + (i64)(m * R + r) * n;
      MMDFN_TRY(colsum(N, O, g, O, 0.f, dFr, st));                                                      // row 0: the constant-1 input
      MMDFN_TRY(gemm(true, false, H, O, N, 1.f, h[m], H, g, O, 0.f, dFr + O, O, nullptr, 0, st));          // rows 1..H
      MMDFN_TRY(gemm(false, true, N, H, O, 1.f, g, O, Fr + O, O, r ? 1.f : 0.f, dh[m], H, nullptr, 0, st));   // dh += g F^T
    }
  return 0;
}
