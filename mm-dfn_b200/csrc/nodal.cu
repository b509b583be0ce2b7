// k13 (☆ SURVEY 8f rank 3): nodal attention of the `relation` graph type's classifier head --
// attentive_node_features + MatchingAttention('general2') (code/model.py:614-645, 66-76) on the RAGGED node rows.
// Per dialogue b with rows E_b (L, D) and projected candidates Q_b = E_b W^T + b (a plain GEMM, done by the caller):
//     S = tanh(Q_b E_b^T)   (L x L)      P = softmax_rows(S)      O_b = P E_b
// (the reference pads every dialogue to T, masks, soft-maxes over all T positions, re-masks and re-normalises: the padded
// positions' exp(0) cancels, so this is a softmax over the valid positions; it also evaluates padded candidates, which
// classify_node_features drops again, :663).  Backward in closed form:
//     dP = dO E_b^T ; dA = P (dP - rowsum(P dP)) (1 - S^2) ; dQ = dA E_b ; dE_b = P^T dO + dA^T Q_b   (+ dQ W by the caller)
// Three small kernels over (32-row tile, dialogue) grids: block products X_b Y_b^T, the row-wise softmax / its
// backward, and block-times-rows products (plain and transposed).  This head is an ablation path (multi_modal=False): the
// kernels are written for clarity, in fp32 FFMA, and are HBM/L2-bound at the dialogue sizes of the data sets.
#include "internal.cuh"
#include "../../include/mmdfn_b200.h"

namespace mmdfn {

constexpr int NA_T = 32;            // tile edge
constexpr int NA_KC = 64;           // contraction chunk of the block products

// R_b[t][s] = X_b[t] . Y_b[s]  for the rows t of this tile and all s < L
__global__ void __launch_bounds__(256) nodal_outer_kernel(int D, const int* __restrict__ dia_off, const i64* __restrict__ sq_off,
                                                           const float* __restrict__ X, const float* __restrict__ Y,
                                                           float* __restrict__ R) {
  __shared__ float xs[NA_T][NA_KC + 1];
  __shared__ float ys[NA_T][NA_KC + 1];
  const int b = blockIdx.y, off = dia_off[b], L = dia_off[b + 1] - off;
  const int t0 = blockIdx.x * NA_T;
  if (t0 >= L) return;
  float* Rb = R + sq_off[b];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // output (rows 4 ty .. 4 ty + 3, column tx) of each 32 x 32 tile
  for (int s0 = 0; s0 < L; s0 += NA_T) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < D; k0 += NA_KC) {
      for (int i = threadIdx.x; i < NA_T * NA_KC; i += 256) {
        const int r = i / NA_KC, k = i - r * NA_KC;
        xs[r][k] = (t0 + r < L && k0 + k < D) ? X[(i64)(off + t0 + r) * D + k0 + k] : 0.f;
        ys[r][k] = (s0 + r < L && k0 + k < D) ? Y[(i64)(off + s0 + r) * D + k0 + k] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < NA_KC; k++) {
        const float y = ys[tx][k];
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i] = fmaf(xs[4 * ty + i][k], y, acc[i]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int t = t0 + 4 * ty + i, s = s0 + tx;
      if (t < L && s < L) Rb[(i64)t * L + s] = acc[i];
    }
  }
}

// forward rows: S = tanh(raw), P = softmax(S) over the L entries of the row (warp per row; raw arrives in P)
__global__ void nodal_softmax_kernel(int B, const int* __restrict__ dia_off, const i64* __restrict__ sq_off,
                                     float* __restrict__ P, float* __restrict__ S, int rows_total, const int* __restrict__ row_dia) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows_total) return;
  const int b = row_dia[row], off = dia_off[b], L = dia_off[b + 1] - off;
  const i64 base = sq_off[b] + (i64)(row - off) * L;
  float mx = -INFINITY;
  for (int s = lane; s < L; s += 32) {
    const float v = tanhf(P[base + s]);
    S[base + s] = v;
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int s = lane; s < L; s += 32) sum += expf(S[base + s] - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int s = lane; s < L; s += 32) P[base + s] = expf(S[base + s] - mx) * inv;
}

// backward rows, in place on dA (arrives holding dP): dA = P (dP - sum_s P dP) (1 - S^2)
__global__ void nodal_softmax_bwd_kernel(int B, const int* __restrict__ dia_off, const i64* __restrict__ sq_off,
                                         const float* __restrict__ P, const float* __restrict__ S, float* __restrict__ dA,
                                         int rows_total, const int* __restrict__ row_dia) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows_total) return;
  const int b = row_dia[row], off = dia_off[b], L = dia_off[b + 1] - off;
  const i64 base = sq_off[b] + (i64)(row - off) * L;
  float dot = 0.f;
  for (int s = lane; s < L; s += 32) dot = fmaf(P[base + s], dA[base + s], dot);
  dot = warp_sum(dot);
  for (int s = lane; s < L; s += 32) {
    const float sv = S[base + s];
    dA[base + s] = P[base + s] * (dA[base + s] - dot) * (1.f - sv * sv);
  }
}

// OUT_b[r][:] (+)= sum_j W_b[r][j] V_b[j][:]            (TRANS = false)
// OUT_b[r][:] (+)= sum_j W_b[j][r] V_b[j][:]            (TRANS = true)       D <= 512 (two columns per thread)
template <bool TRANS>
__global__ void __launch_bounds__(256) nodal_apply_kernel(int D, const int* __restrict__ dia_off, const i64* __restrict__ sq_off,
                                                           const float* __restrict__ W, const float* __restrict__ V,
                                                           float* __restrict__ OUT, int accumulate) {
  __shared__ float ws[NA_T][NA_T + 1];          // ws[r][j] = weight of (output row r, contraction index j) of this tile pair
  const int b = blockIdx.y, off = dia_off[b], L = dia_off[b + 1] - off;
  const int r0 = blockIdx.x * NA_T;
  if (r0 >= L) return;
  const float* Wb = W + sq_off[b];
  const int c0 = threadIdx.x, c1 = threadIdx.x + 256;
  float acc0[NA_T], acc1[NA_T];
#pragma unroll
  for (int r = 0; r < NA_T; r++) acc0[r] = acc1[r] = 0.f;
  for (int j0 = 0; j0 < L; j0 += NA_T) {
    for (int i = threadIdx.x; i < NA_T * NA_T; i += 256) {
      const int a = i >> 5, c = i & 31;        // coalesced along the block's rows: element (a, c) of the 32 x 32 source tile
      float v = 0.f;
      if (!TRANS) { if (r0 + a < L && j0 + c < L) v = Wb[(i64)(r0 + a) * L + j0 + c]; ws[a][c] = v; }
      else { if (j0 + a < L && r0 + c < L) v = Wb[(i64)(j0 + a) * L + r0 + c]; ws[c][a] = v; }
    }
    __syncthreads();
    const int jn = min(NA_T, L - j0);
    for (int j = 0; j < jn; j++) {
      const float* vr = V + (i64)(off + j0 + j) * D;
      const float v0 = c0 < D ? __ldg(vr + c0) : 0.f;
      const float v1 = c1 < D ? __ldg(vr + c1) : 0.f;
#pragma unroll
      for (int r = 0; r < NA_T; r++) {
        const float w = ws[r][j];
        acc0[r] = fmaf(w, v0, acc0[r]);
        acc1[r] = fmaf(w, v1, acc1[r]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < NA_T; r++) {
    if (r0 + r < L) {
      float* o = OUT + (i64)(off + r0 + r) * D;
      if (c0 < D) o[c0] = accumulate ? o[c0] + acc0[r] : acc0[r];
      if (c1 < D) o[c1] = accumulate ? o[c1] + acc1[r] : acc1[r];
    }
  }
}

}  // namespace mmdfn

using namespace mmdfn;

/* P and S: sum_b L_b^2 floats (block b at sq_off[b], row-major L_b x L_b); row_dia (N): dialogue of every node row */
extern "C" int mmdfn_nodal_attn_fwd(int B, int N, int D, int Lmax, const int* dia_off, const long long* sq_off,
                                    const int* row_dia, const float* E, const float* Q, float* P, float* S, float* O,
                                    void* stream) {
  if (!dia_off || !sq_off || !row_dia || !E || !Q || !P || !S || !O) return MMDFN_ENULL;
  if (B < 0 || N < 0 || D <= 0 || D > 512 || Lmax < 0) return MMDFN_EINVAL;
  if (B == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(ceil_div(Lmax, NA_T), B);
  nodal_outer_kernel<<<grid, 256, 0, st>>>(D, dia_off, (const i64*)sq_off, Q, E, P);
  MMDFN_LAUNCH_CHECK();
  nodal_softmax_kernel<<<ceil_div(N, 8), 256, 0, st>>>(B, dia_off, (const i64*)sq_off, P, S, N, row_dia);
  MMDFN_LAUNCH_CHECK();
  nodal_apply_kernel<false><<<grid, 256, 0, st>>>(D, dia_off, (const i64*)sq_off, P, E, O, 0);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

/* dA: workspace of sum_b L_b^2 floats.  Outputs (overwritten): dQ (N, D) = dA E and dE (N, D) = P^T dO + dA^T Q -- the caller
   adds the path through the projection (dE += dQ W) and forms dW = dQ^T E, db = colsum(dQ) */
extern "C" int mmdfn_nodal_attn_bwd(int B, int N, int D, int Lmax, const int* dia_off, const long long* sq_off,
                                    const int* row_dia, const float* E, const float* Q, const float* P, const float* S,
                                    const float* dO, float* dA, float* dQ, float* dE, void* stream) {
  if (!dia_off || !sq_off || !row_dia || !E || !Q || !P || !S || !dO || !dA || !dQ || !dE) return MMDFN_ENULL;
  if (B < 0 || N < 0 || D <= 0 || D > 512 || Lmax < 0) return MMDFN_EINVAL;
  if (B == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(ceil_div(Lmax, NA_T), B);
  nodal_outer_kernel<<<grid, 256, 0, st>>>(D, dia_off, (const i64*)sq_off, dO, E, dA);              // dP
  MMDFN_LAUNCH_CHECK();
  nodal_softmax_bwd_kernel<<<ceil_div(N, 8), 256, 0, st>>>(B, dia_off, (const i64*)sq_off, P, S, dA, N, row_dia);
  MMDFN_LAUNCH_CHECK();
  nodal_apply_kernel<false><<<grid, 256, 0, st>>>(D, dia_off, (const i64*)sq_off, dA, E, dQ, 0);
  MMDFN_LAUNCH_CHECK();
  nodal_apply_kernel<true><<<grid, 256, 0, st>>>(D, dia_off, (const i64*)sq_off, P, dO, dE, 0);
  MMDFN_LAUNCH_CHECK();
  nodal_apply_kernel<true><<<grid, 256, 0, st>>>(D, dia_off, (const i64*)sq_off, dA, Q, dE, 1);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
