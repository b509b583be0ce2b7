// k12 (relation graph type): RGCNConv -> GraphConv of GraphNetwork (code/model.py:675-715), whose arithmetic lives in
// torch-geometric 1.4.3 (not vendored; semantics restated in oracle/mmdfn_oracle.py, parity unpinned):
//   RGCNConv : out_i = sum_{j->i} norm_e * (x_j W_{type_e}) + x_i root + bias,  W_r = sum_b att[r,b] basis[b]
//   GraphConv: out_i = sum_{j->i} (x W)_j + Linear(x_i)
// Replaces torch-scatter's atomic scatter_add: with windowed edges every node's in-edges (sources i-wf..i+wp) and
// out-edges (targets j-wp..j+wf) are contiguous ranges of the same dialogue, so both directions are gather-reduces
// (warp per node, lanes over the feature vector, no atomics).  The dense parts (x W_all, x root, weight gradients)
// run on mmdfn_gemm.
#include "internal.cuh"

namespace mmdfn {

__device__ __forceinline__ int rg_lo(int j, int wp) { return wp < 0 ? 0 : max(0, j - wp); }

struct RgGeom {
  int N, G, R, S, wp, wf;
  const int* dia_off;
  const int* node_dia;
  const int* node_spk;
  const i64* row_ptr;
};

// out[n, :] += sum_j norm[e(j,i)] * XW[j, type(j,i), :]            warp per target node n
__global__ void rgcn_gather_fwd_kernel(RgGeom g, const float* __restrict__ xw, const float* __restrict__ norm,
                                       float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.y + threadIdx.y;
  if (n >= g.N) return;
  const int lane = threadIdx.x;
  const int b = g.node_dia[n];
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const int i = n - off;
  const int jlo = g.wf < 0 ? 0 : max(0, i - g.wf), jhi = g.wp < 0 ? L - 1 : min(L - 1, i + g.wp);
  const int si = g.node_spk[n];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};          // G <= 128
  for (int j = jlo; j <= jhi; j++) {
    const i64 e = g.row_ptr[off + j] + (i - rg_lo(j, g.wp));
    const int type = 2 * (g.S * g.node_spk[off + j] + si) + (j >= i ? 1 : 0);
    const float w = norm[e];
    const float* src = xw + ((i64)(off + j) * g.R + type) * g.G;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int c = lane + 32 * q;
      if (c < g.G) acc[q] = fmaf(w, src[c], acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int c = lane + 32 * q;
    if (c < g.G) out[(i64)n * g.G + c] += acc[q];
  }
}

// warp per source node j: dXW[j, type, :] += norm[e] * dout[i, :] (row pre-zeroed) ; dnorm[e] = dout[i, :] . XW[j, type, :]
__global__ void rgcn_scatter_bwd_kernel(RgGeom g, const float* __restrict__ xw, const float* __restrict__ norm,
                                        const float* __restrict__ dout, float* __restrict__ dxw,
                                        float* __restrict__ dnorm) {
  const int n = blockIdx.x * blockDim.y + threadIdx.y;
  if (n >= g.N) return;
  const int lane = threadIdx.x;
  const int b = g.node_dia[n];
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const int j = n - off;
  const int lo = rg_lo(j, g.wp), hi = g.wf < 0 ? L : min(L, j + g.wf + 1);
  const int sj = g.node_spk[n];
  const i64 e0 = g.row_ptr[n];
  for (int i = lo; i < hi; i++) {
    const i64 e = e0 + (i - lo);
    const int type = 2 * (g.S * sj + g.node_spk[off + i]) + (j >= i ? 1 : 0);
    const float w = norm[e];
    const float* go = dout + (i64)(off + i) * g.G;
    const float* xs = xw + ((i64)n * g.R + type) * g.G;
    float* dx = dxw + ((i64)n * g.R + type) * g.G;
    float dot = 0.f;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int c = lane + 32 * q;
      if (c < g.G) {
        const float gv = go[c];
        dx[c] += w * gv;                      // only this warp touches row n; same (type) slots are revisited sequentially
        dot = fmaf(gv, xs[c], dot);
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) dnorm[e] = dot;
    __syncwarp();
  }
}

// out[n, :] (+)= sum_{j in [i-a, i+b] within the dialogue} h[j, :]   (-1 = unbounded)      warp per node
__global__ void window_sum_kernel(int N, int G, int a, int b, const int* __restrict__ dia_off,
                                  const int* __restrict__ node_dia, const float* __restrict__ h, float* __restrict__ out,
                                  int accumulate) {
  const int n = blockIdx.x * blockDim.y + threadIdx.y;
  if (n >= N) return;
  const int lane = threadIdx.x;
  const int d = node_dia[n];
  const int off = dia_off[d], L = dia_off[d + 1] - off;
  const int i = n - off;
  const int jlo = a < 0 ? 0 : max(0, i - a), jhi = b < 0 ? L - 1 : min(L - 1, i + b);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = jlo; j <= jhi; j++) {
    const float* src = h + (i64)(off + j) * G;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int c = lane + 32 * q;
      if (c < G) acc[q] += src[c];
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int c = lane + 32 * q;
    if (c < G) {
      float* o = out + (i64)n * G + c;
      *o = accumulate ? *o + acc[q] : acc[q];
    }
  }
}

// out[b][c][r] = in[b][r][c]
__global__ void transpose_batched_kernel(int rows, int cols, const float* __restrict__ in, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const i64 base = (i64)blockIdx.z * rows * cols;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int r = r0 + y, c = c0 + threadIdx.x;
    tile[y][threadIdx.x] = (r < rows && c < cols) ? in[base + (i64)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int c = c0 + y, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[base + (i64)c * rows + r] = tile[threadIdx.x][y];
  }
}

// node_spk[n] = first p with qmask == 1 (0 if none), same rule as the edge types
__global__ void node_speaker_kernel(int B, int S, const int* __restrict__ dia_off, const float* __restrict__ qmask,
                                    int* __restrict__ node_spk) {
  const int b = blockIdx.x;
  const int off = dia_off[b], L = dia_off[b + 1] - off;
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    int s = 0;
    for (int p = 0; p < S; p++)
      if (qmask[((i64)t * B + b) * S + p] == 1.0f) { s = p; break; }
    node_spk[off + t] = s;
  }
}

}  // namespace mmdfn

using namespace mmdfn;

extern "C" int mmdfn_node_speakers(int B, int S, const int* dia_off, const float* qmask, int* node_spk, void* stream) {
  if (!dia_off || !qmask || !node_spk) return MMDFN_ENULL;
  if (B <= 0) return 0;
  node_speaker_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(B, S, dia_off, qmask, node_spk);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_rgcn_aggregate_fwd(int N, int G, int R, int S, int window_past, int window_future,
                                        const int* dia_off, const int* node_dia, const int* node_spk,
                                        const long long* row_ptr, const float* xw, const float* norm, float* out,
                                        void* stream) {
  if (!dia_off || !node_dia || !node_spk || !row_ptr || !xw || !norm || !out) return MMDFN_ENULL;
  if (G <= 0 || G > 128) return MMDFN_EINVAL;
  if (N <= 0) return 0;
  RgGeom g{N, G, R, S, window_past, window_future, dia_off, node_dia, node_spk, (const i64*)row_ptr};
  rgcn_gather_fwd_kernel<<<ceil_div(N, 8), dim3(32, 8), 0, (cudaStream_t)stream>>>(g, xw, norm, out);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_rgcn_aggregate_bwd(int N, int G, int R, int S, int window_past, int window_future,
                                        const int* dia_off, const int* node_dia, const int* node_spk,
                                        const long long* row_ptr, const float* xw, const float* norm, const float* dout,
                                        float* dxw, float* dnorm, void* stream) {
  if (!dia_off || !node_dia || !node_spk || !row_ptr || !xw || !norm || !dout || !dxw || !dnorm) return MMDFN_ENULL;
  if (G <= 0 || G > 128) return MMDFN_EINVAL;
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  MMDFN_TRY(fill_zero(dxw, (size_t)N * R * G * sizeof(float), st));
  RgGeom g{N, G, R, S, window_past, window_future, dia_off, node_dia, node_spk, (const i64*)row_ptr};
  rgcn_scatter_bwd_kernel<<<ceil_div(N, 8), dim3(32, 8), 0, st>>>(g, xw, norm, dout, dxw, dnorm);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_window_sum(int N, int G, int reach_back, int reach_fwd, const int* dia_off, const int* node_dia,
                                const float* h, float* out, int accumulate, void* stream) {
  if (!dia_off || !node_dia || !h || !out) return MMDFN_ENULL;
  if (G <= 0 || G > 128) return MMDFN_EINVAL;
  if (N <= 0) return 0;
  window_sum_kernel<<<ceil_div(N, 8), dim3(32, 8), 0, (cudaStream_t)stream>>>(N, G, reach_back, reach_fwd, dia_off, node_dia, h,
                                                                          out, accumulate);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_transpose_batched(int batch, int rows, int cols, const float* in, float* out, void* stream) {
  if (!in || !out) return MMDFN_ENULL;
  if (batch <= 0 || rows <= 0 || cols <= 0) return 0;
  transpose_batched_kernel<<<dim3(ceil_div(cols, 32), ceil_div(rows, 32), batch), dim3(32, 8), 0, (cudaStream_t)stream>>>(rows, cols, in, out);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
