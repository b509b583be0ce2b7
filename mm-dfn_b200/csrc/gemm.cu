// Dense fp32 GEMM entry point (k1 projections and every other dense contraction of the
// path: GRU/LSTM gate GEMMs, fcs[0], [hi|h0]W, classifier, and all weight gradients).
// Replaces the cuBLAS sgemm calls behind nn.Linear / torch.mm at
// code/model.py:1065,1094,1129,1337 ; code/model_GCN.py:186,454,466.
#include "gemm_tile.cuh"
#include "internal.cuh"
#include "../../include/mmdfn_b200.h"

namespace mmdfn {

struct GemmArgs {
  const float* A; i64 lda;
  const float* B; i64 ldb;
  float* C; i64 ldc;
  const float* bias;
  int M, N, K;
  float alpha, beta;
  int act, splits;
};

template <int BM, int BN, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_kernel(GemmArgs p) {
  __shared__ __align__(16) float smem[GemmSmem<BM, BN>::FLOATS];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  int kb = 0, ke = p.K;
  if (p.splits > 1) {
    const int chunk = ((p.K + p.splits - 1) / p.splits + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
    kb = blockIdx.z * chunk;
    ke = min(p.K, kb + chunk);
  }
  float acc[TM][TN];
  zero_acc(acc);
  gemm_tile_accum<BM, BN, TM, TN, TA, TB>(p.A, p.lda, p.B, p.ldb, p.M, p.N, m0, n0, kb, ke, acc, smem);
  const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);
  const bool vec_ok = (p.splits <= 1) && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll
  for (int i = 0; i < TM; i++) {
    const int m = m0 + tile_row<BM, TM>(ty, i);
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      const int n = n0 + tile_col<BN, TN>(tx, j);
      if (n >= p.N) continue;
      float* c = p.C + (i64)m * p.ldc + n;
      if (p.splits > 1) {
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (n + q < p.N) atomicAdd(c + q, p.alpha * acc[i][j + q]);
        continue;
      }
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        float t = p.alpha * acc[i][j + q];
        if (n + q < p.N) {
          if (p.beta != 0.f) t = fmaf(p.beta, c[q], t);
          if (p.bias) t += p.bias[n + q];
        }
        if (p.act == 1) t = fmaxf(t, 0.f);
        v[q] = t;
      }
      if (vec_ok && n + 3 < p.N) {
        *reinterpret_cast<float4*>(c) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (n + q < p.N) c[q] = v[q];
      }
    }
  }
}

__global__ void scale2d_kernel(float* C, i64 ldc, int M, int N, float beta) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (i64)M * N) return;
  const int m = (int)(idx / N), n = (int)(idx % N);
  float* c = C + (i64)m * ldc + n;
  *c = (beta == 0.f) ? 0.f : beta * *c;
}

template <int BM, int BN, int TM, int TN>
static int launch_gemm(bool ta, bool tb, const GemmArgs& p, cudaStream_t st) {
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, BN), p.splits > 1 ? p.splits : 1);
  if (!ta && tb) gemm_kernel<BM, BN, TM, TN, false, true><<<grid, GEMM_THREADS, 0, st>>>(p);
  else if (!ta && !tb) gemm_kernel<BM, BN, TM, TN, false, false><<<grid, GEMM_THREADS, 0, st>>>(p);
  else if (ta && !tb) gemm_kernel<BM, BN, TM, TN, true, false><<<grid, GEMM_THREADS, 0, st>>>(p);
  else return MMDFN_EINVAL;
  MMDFN_LAUNCH_CHECK();
  return 0;
}

int gemm(bool ta, bool tb, int M, int N, int K, float alpha, const float* A, i64 lda, const float* B, i64 ldb,
         float beta, float* C, i64 ldc, const float* bias, int act, cudaStream_t st) {
  if (M < 0 || N < 0 || K < 0) return MMDFN_EINVAL;
  if (M == 0 || N == 0) return 0;
  if (!A || !B || !C) return MMDFN_ENULL;
  if (ta && tb) return MMDFN_EINVAL;
  // large contractions run on the tensor cores (tcgen05 kind::tf32, 3-term split: fp32-level accuracy);
  // small ones stay on the FFMA tile kernel, whose prologue is cheaper than a TMEM allocation
  // (long TN contractions -- weight gradients -- earlier: the split-K FFMA kernel takes 33 us for the text projection's
  // 200 x 100 x 3200 weight gradient, the tensor-core kernel ~12 us for 300 x 100 x 3168)
  const double tc_min = (ta && K >= 2048) ? 1.0e8 : 1.8e8;
  if (2.0 * (double)M * (double)N * (double)K >= tc_min)
    return umma_gemm(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act, st);
  GemmArgs p{A, lda, B, ldb, C, ldc, bias, M, N, K, alpha, beta, act, 1};
  // 128 x 64 tiles (8 x 4 per thread) only when even they give every SM several CTAs; otherwise 64 x 64 tiles: twice the
  // CTAs and warps per SM (the 128 x 64 kernel ran one 8-warp CTA per SM at 12 % occupancy and 40 % issue utilisation
  // on the 9600 x 100 x 100 products of the graph layers)
  const bool big = (i64)ceil_div(M, 128) * ceil_div(N, 64) >= 3 * 148;
  const i64 tiles = big ? (i64)ceil_div(M, 128) * ceil_div(N, 64) : (i64)ceil_div(M, 64) * ceil_div(N, 64);
  // split the contraction when the output is too small to fill 148 SMs (weight gradients)
  if (tiles < 296 && K >= 1024 && bias == nullptr && act == 0) {
    i64 s = ceil_div64(592, tiles);
    const i64 smax = ceil_div(K, 128);
    p.splits = (int)(s < smax ? s : smax);
    if (p.splits < 1) p.splits = 1;
  }
  if (p.splits > 1 && beta != 1.f) {
    const i64 tot = (i64)M * N;
    scale2d_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(C, ldc, M, N, beta);
    MMDFN_LAUNCH_CHECK();
  }
  return big ? launch_gemm<128, 64, 8, 4>(ta, tb, p, st) : launch_gemm<64, 64, 4, 4>(ta, tb, p, st);
}

// ---- column sums (bias gradients) ----------------------------------------------------
__global__ void colsum_kernel(const float* __restrict__ A, i64 lda, int M, int N, int rows_per_block,
                              float* __restrict__ out) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int mb = blockIdx.y * rows_per_block;
  const int me = min(M, mb + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int m = mb + threadIdx.y; m < me; m += 8) s += A[(i64)m * lda + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
#pragma unroll
    for (int i = 1; i < 8; i++) s += red[i][threadIdx.x];
    atomicAdd(out + n, s);
  }
}

int colsum(int M, int N, const float* A, i64 lda, float beta, float* out, cudaStream_t st) {
  if (N <= 0) return 0;
  if (beta != 1.f) {
    scale2d_kernel<<<ceil_div(N, 256), 256, 0, st>>>(out, N, 1, N, beta);
    MMDFN_LAUNCH_CHECK();
  }
  if (M <= 0) return 0;
  const int rpb = 256;
  dim3 grid(ceil_div(N, 32), ceil_div(M, rpb));
  colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(A, lda, M, N, rpb, out);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

int fill_zero(void* p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return 0;
  MMDFN_CUDA(cudaMemsetAsync(p, 0, bytes, st));
  return 0;
}

}  // namespace mmdfn

extern "C" int mmdfn_gemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, long long lda,
                          const float* B, long long ldb, float beta, float* C, long long ldc, const float* bias,
                          int act, void* stream) {
  return mmdfn::gemm(transA != 0, transB != 0, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act,
                     (cudaStream_t)stream);
}

/* zero-fill through cudaMemsetAsync (a memset node: no fill kernel on the stream) */
extern "C" int mmdfn_memset_zero(void* p, long long bytes, void* stream) {
  if (bytes < 0) return MMDFN_EINVAL;
  if (bytes == 0) return 0;
  if (!p) return MMDFN_ENULL;
  return mmdfn::fill_zero(p, (size_t)bytes, (cudaStream_t)stream);
}

extern "C" int mmdfn_colsum(int M, int N, const float* A, long long lda, float beta, float* out, void* stream) {
  if (!A || !out) return MMDFN_ENULL;
  return mmdfn::colsum(M, N, A, lda, beta, out, (cudaStream_t)stream);
}
