// k10/k11 (relation graph type): windowed edge construction and masked edge attention.
// Replaces edge_perms + the per-edge Python loop of batch_graphify (code/model.py:532-550,
// 568-611; two .nonzero() device syncs per edge) and MaskedEdgeAttention 'attn1'
// (code/model.py:449-471; host-built masks copied H2D on every call).
//
// Edge order is the canonical one: dialogue-major, then source j ascending, then target i
// ascending (the reference emits CPython-set order, which is not reproducible across Python
// versions; the bit-exact contract is on this sorted list -- SURVEY.md 8a row a10).
// Every source row j owns the contiguous edge range row_ptr[j] .. row_ptr[j+1]-1 with targets
// lo(j)..hi(j)-1, lo = max(0, j-wp), hi = min(L, j+wf+1)  (wp/wf = -1: unbounded).
#include "internal.cuh"

namespace mmdfn {

__device__ __forceinline__ int win_lo(int j, int wp) { return wp < 0 ? 0 : max(0, j - wp); }
__device__ __forceinline__ int win_hi(int j, int L, int wf) { return wf < 0 ? L : min(L, j + wf + 1); }

// one block per dialogue; thread j derives its row start by summing the (closed-form) counts of rows < j
__global__ void edges_build_kernel(int T, int B, int S, int wp, int wf, const int* __restrict__ dia_off,
                                   const i64* __restrict__ edge_off, const float* __restrict__ qmask,
                                   i64* __restrict__ edge_index, i64 E, i64* __restrict__ edge_type,
                                   i64* __restrict__ row_ptr) {
  const int b = blockIdx.x;
  const int off = dia_off[b], L = dia_off[b + 1] - off;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    i64 start = edge_off[b];
    for (int jj = 0; jj < j; jj++) start += win_hi(jj, L, wf) - win_lo(jj, wp);
    row_ptr[off + j] = start;
    // speaker of j: first p with qmask == 1 (code/model.py:591)
    int sj = 0;
    for (int p = 0; p < S; p++)
      if (qmask[((i64)j * B + b) * S + p] == 1.0f) { sj = p; break; }
    const int lo = win_lo(j, wp), hi = win_hi(j, L, wf);
    for (int i = lo; i < hi; i++) {
      int si = 0;
      for (int p = 0; p < S; p++)
        if (qmask[((i64)i * B + b) * S + p] == 1.0f) { si = p; break; }
      const i64 e = start + (i - lo);
      edge_index[e] = off + j;
      edge_index[E + e] = off + i;
      edge_type[e] = 2 * ((i64)S * sj + si) + (j >= i ? 1 : 0);
    }
  }
  if (threadIdx.x == 0 && b == B - 1) row_ptr[dia_off[B]] = edge_off[B];
}

// warp per (dialogue b, source row j < L_b): softmax over all T time steps of s[t,b,j], window mass renormalised
// with the reference's 1e-10 leak from the non-edge positions.  s_all: (T*B, ncol) rows (t,b).
__global__ void edge_attn_fwd_kernel(int T, int B, int ncol, int wp, int wf, const int* __restrict__ dia_off,
                                     const i64* __restrict__ row_ptr, const float* __restrict__ s_all,
                                     float* __restrict__ edge_norm, float* __restrict__ stat) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.y + threadIdx.y;
  const int off = dia_off[b], L = dia_off[b + 1] - off;
  if (j >= L) return;
  const int lane = threadIdx.x;
  const float* s = s_all + (i64)b * ncol + j;
  const i64 stride = (i64)B * ncol;
  float mx = -INFINITY;
  for (int i = lane; i < T; i += 32) mx = fmaxf(mx, s[i * stride]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int lo = win_lo(j, wp), hi = win_hi(j, L, wf);
  float z = 0.f, wn = 0.f;
  for (int i = lane; i < T; i += 32) {
    const float e = expf(s[i * stride] - mx);
    z += e;
    if (i >= lo && i < hi) wn += e;
  }
  z = warp_sum(z);
  wn = warp_sum(wn);
  const float d = wn + 1e-10f * (z - wn);
  const i64 e0 = row_ptr[off + j];
  for (int i = lo + lane; i < hi; i += 32) edge_norm[e0 + (i - lo)] = expf(s[i * stride] - mx) / d;
  if (lane == 0) { stat[((i64)off + j) * 2] = mx; stat[((i64)off + j) * 2 + 1] = d; }
}

// ds[t,b,j] for all t < T (rows j >= L_b and columns >= ncol_used are zero-filled by the caller's memset)
__global__ void edge_attn_bwd_kernel(int T, int B, int ncol, int wp, int wf, const int* __restrict__ dia_off,
                                     const i64* __restrict__ row_ptr, const float* __restrict__ s_all,
                                     const float* __restrict__ edge_norm, const float* __restrict__ stat,
                                     const float* __restrict__ g_norm, float* __restrict__ ds_all) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.y + threadIdx.y;
  const int off = dia_off[b], L = dia_off[b + 1] - off;
  if (j >= L) return;
  const int lane = threadIdx.x;
  const int lo = win_lo(j, wp), hi = win_hi(j, L, wf);
  const i64 e0 = row_ptr[off + j];
  float t = 0.f;
  for (int i = lo + lane; i < hi; i += 32) t = fmaf(g_norm[e0 + (i - lo)], edge_norm[e0 + (i - lo)], t);
  t = warp_sum(t);
  const float mx = stat[((i64)off + j) * 2], d = stat[((i64)off + j) * 2 + 1];
  const i64 stride = (i64)B * ncol;
  const float* s = s_all + (i64)b * ncol + j;
  float* ds = ds_all + (i64)b * ncol + j;
  for (int i = lane; i < T; i += 32) {
    float v;
    if (i >= lo && i < hi) {
      const float sc = edge_norm[e0 + (i - lo)];
      v = sc * (g_norm[e0 + (i - lo)] - t);
    } else {
      v = -t * 1e-10f * expf(s[i * stride] - mx) / d;
    }
    ds[i * stride] = v;
  }
}

// dense (B, msl, T) view of the compact edge scores and its adjoint
__global__ void edge_scores_scatter_kernel(i64 E, const i64* __restrict__ edge_index, const int* __restrict__ node_dia,
                                           const int* __restrict__ dia_off, int msl, int T,
                                           float* edge_norm, float* dense, int gather) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const i64 nj = edge_index[e], ni = edge_index[E + e];
  const int b = node_dia[nj];
  const int j = (int)(nj - dia_off[b]), i = (int)(ni - dia_off[b]);
  float* p = dense + ((i64)b * msl + j) * T + i;
  if (gather) edge_norm[e] = *p; else *p = edge_norm[e];
}

__global__ void node_dialogue_kernel(int B, const int* __restrict__ dia_off, int* __restrict__ node_dia) {
  const int b = blockIdx.x;
  for (int n = dia_off[b] + threadIdx.x; n < dia_off[b + 1]; n += blockDim.x) node_dia[n] = b;
}

}  // namespace mmdfn

using namespace mmdfn;

extern "C" int mmdfn_edges_build(int T, int B, int S, int N, int window_past, int window_future, const int* dia_off,
                                 const long long* edge_off, long long E, const float* qmask, long long* edge_index,
                                 long long* edge_type, long long* row_ptr, int* node_dia, void* stream) {
  if (!dia_off || !edge_off || !qmask || !edge_index || !edge_type || !row_ptr || !node_dia) return MMDFN_ENULL;
  if (B <= 0 || N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  edges_build_kernel<<<B, 128, 0, st>>>(T, B, S, window_past, window_future, dia_off, (const i64*)edge_off, qmask,
                                        (i64*)edge_index, (i64)E, (i64*)edge_type, (i64*)row_ptr);
  MMDFN_LAUNCH_CHECK();
  node_dialogue_kernel<<<B, 128, 0, st>>>(B, dia_off, node_dia);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// s_all = M W_att[:ncol]^T is computed by the caller with mmdfn_gemm ((T*B, ncol), ncol >= Lmax).
extern "C" int mmdfn_edge_attn_fwd(int T, int B, int Lmax, int ncol, int window_past, int window_future,
                                   const int* dia_off, const long long* row_ptr, const float* s_all,
                                   float* edge_norm, float* stat, void* stream) {
  if (!dia_off || !row_ptr || !s_all || !edge_norm || !stat) return MMDFN_ENULL;
  if (ncol < Lmax) return MMDFN_EINVAL;      // a dialogue longer than max_seq_len has no attention row (the reference raises)
  if (B <= 0 || Lmax <= 0) return 0;
  edge_attn_fwd_kernel<<<dim3(ceil_div(Lmax, 8), B), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      T, B, ncol, window_past, window_future, dia_off, (const i64*)row_ptr, s_all, edge_norm, stat);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_edge_attn_bwd(int T, int B, int Lmax, int ncol, int window_past, int window_future,
                                   const int* dia_off, const long long* row_ptr, const float* s_all,
                                   const float* edge_norm, const float* stat, const float* g_norm, float* ds_all,
                                   void* stream) {
  if (!dia_off || !row_ptr || !s_all || !edge_norm || !stat || !g_norm || !ds_all) return MMDFN_ENULL;
  cudaStream_t st = (cudaStream_t)stream;
  MMDFN_TRY(fill_zero(ds_all, (size_t)T * B * ncol * sizeof(float), st));
  if (B <= 0 || Lmax <= 0) return 0;
  edge_attn_bwd_kernel<<<dim3(ceil_div(Lmax, 8), B), dim3(32, 8), 0, st>>>(
      T, B, ncol, window_past, window_future, dia_off, (const i64*)row_ptr, s_all, edge_norm, stat, g_norm, ds_all);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// gather == 0: dense[b, j, i] = edge_norm[e] (dense pre-zeroed by this call); gather == 1: edge_norm[e] = dense[b, j, i]
extern "C" int mmdfn_edge_scores_dense(long long E, int B, int msl, int T, const long long* edge_index,
                                       const int* node_dia, const int* dia_off, float* edge_norm, float* dense,
                                       int gather, void* stream) {
  if (!edge_index || !node_dia || !dia_off || !edge_norm || !dense) return MMDFN_ENULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (!gather) MMDFN_TRY(fill_zero(dense, (size_t)B * msl * T * sizeof(float), st));
  if (E <= 0) return 0;
  edge_scores_scatter_kernel<<<(unsigned)ceil_div64(E, 256), 256, 0, st>>>((i64)E, (const i64*)edge_index, node_dia, dia_off, msl, T,
                                                                           edge_norm, dense, gather);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
