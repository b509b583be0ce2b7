// a13: MMGatedAttention ('general', code/model.py:757-781) -- the gated pairwise fusion of the three modality
// streams:   h_m = tanh(W_m x_m + b_m),   z_mn = sigmoid(w_mn . [x_m, x_n, x_m * x_n] + b_mn),
//            out = [ z_av h_a + (1 - z_av) h_v | z_al h_a + (1 - z_al) h_l | z_vl h_v + (1 - z_vl) h_l ].
// The three projections W_m x_m are dense GEMMs (mmdfn_gemm, with their own backward); everything else is fused here:
// one warp per utterance row computes the three gate dot products over the 3 x D features with a warp reduction, the
// tanh of the projected rows and the three mixes (forward), or all row-local gradients (backward).  The gate weight
// gradients are column sums over rows of dz_pre * [x_m, x_n, x_m * x_n]: a second, column-parallel kernel.
// HBM-bound: per row 3 D + 3 C floats in, 3 C out (forward).
#include "common.cuh"

namespace mmdfn {

constexpr int GT_WARPS = 8;

// gate g in {0: av, 1: al, 2: vl}: first / second operand stream of its features
__device__ __forceinline__ int gt_first(int g) { return g == 2 ? 1 : 0; }
__device__ __forceinline__ int gt_second(int g) { return g == 0 ? 1 : 2; }

struct GatedArgs {
  int N, D, C;
  const float* x[3];      // (N, D) a, v, l (after dropout)
  const float* P[3];      // (N, C) projections W_m x_m + b_m
  const float* w;         // (3, 3D) gate weights [av | al | vl], each [w1 (D) | w2 (D) | w3 (D)]
  const float* b;         // (3)
};

__global__ void __launch_bounds__(32 * GT_WARPS) gated_fuse_fwd_kernel(GatedArgs p, float* __restrict__ out,
                                                                        float* __restrict__ z) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * GT_WARPS + (threadIdx.x >> 5);
  if (r >= p.N) return;                                     // warp-uniform
  const int D = p.D, C = p.C;
  const float* xa = p.x[0] + (i64)r * D;
  const float* xv = p.x[1] + (i64)r * D;
  const float* xl = p.x[2] + (i64)r * D;
  const float* wav = p.w;
  const float* wal = p.w + 3 * D;
  const float* wvl = p.w + 6 * D;
  float s_av = 0.f, s_al = 0.f, s_vl = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float a = xa[k], v = xv[k], l = xl[k];
    s_av += wav[k] * a + wav[D + k] * v + wav[2 * D + k] * (a * v);
    s_al += wal[k] * a + wal[D + k] * l + wal[2 * D + k] * (a * l);
    s_vl += wvl[k] * v + wvl[D + k] * l + wvl[2 * D + k] * (v * l);
  }
  const float z_av = sigmoidf_(warp_sum(s_av) + p.b[0]);
  const float z_al = sigmoidf_(warp_sum(s_al) + p.b[1]);
  const float z_vl = sigmoidf_(warp_sum(s_vl) + p.b[2]);
  if (lane == 0) {
    z[(i64)r * 3] = z_av;
    z[(i64)r * 3 + 1] = z_al;
    z[(i64)r * 3 + 2] = z_vl;
  }
  float* o = out + (i64)r * 3 * C;
  for (int c = lane; c < C; c += 32) {
    const float ha = tanhf(p.P[0][(i64)r * C + c]);
    const float hv = tanhf(p.P[1][(i64)r * C + c]);
    const float hl = tanhf(p.P[2][(i64)r * C + c]);
    o[c] = z_av * ha + (1.f - z_av) * hv;
    o[C + c] = z_al * ha + (1.f - z_al) * hl;
    o[2 * C + c] = z_vl * hv + (1.f - z_vl) * hl;
  }
}

struct GatedBwdOut {
  float* dP[3];           // (N, C)
  float* dx[3];           // (N, D): the part of d/dx that flows through the gates (the GEMMs add theirs)
  float* dzpre;           // (N, 3): d/d(gate pre-activation), consumed by the weight-gradient kernel
};

__global__ void __launch_bounds__(32 * GT_WARPS) gated_fuse_bwd_kernel(GatedArgs p, const float* __restrict__ dout,
                                                                        const float* __restrict__ z, GatedBwdOut o) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * GT_WARPS + (threadIdx.x >> 5);
  if (r >= p.N) return;
  const int D = p.D, C = p.C;
  const float z_av = z[(i64)r * 3], z_al = z[(i64)r * 3 + 1], z_vl = z[(i64)r * 3 + 2];
  const float* g = dout + (i64)r * 3 * C;
  float t_av = 0.f, t_al = 0.f, t_vl = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float ha = tanhf(p.P[0][(i64)r * C + c]);
    const float hv = tanhf(p.P[1][(i64)r * C + c]);
    const float hl = tanhf(p.P[2][(i64)r * C + c]);
    const float g_av = g[c], g_al = g[C + c], g_vl = g[2 * C + c];
    t_av += g_av * (ha - hv);
    t_al += g_al * (ha - hl);
    t_vl += g_vl * (hv - hl);
    o.dP[0][(i64)r * C + c] = (g_av * z_av + g_al * z_al) * (1.f - ha * ha);
    o.dP[1][(i64)r * C + c] = (g_av * (1.f - z_av) + g_vl * z_vl) * (1.f - hv * hv);
    o.dP[2][(i64)r * C + c] = (g_al * (1.f - z_al) + g_vl * (1.f - z_vl)) * (1.f - hl * hl);
  }
  const float q_av = warp_sum(t_av) * z_av * (1.f - z_av);
  const float q_al = warp_sum(t_al) * z_al * (1.f - z_al);
  const float q_vl = warp_sum(t_vl) * z_vl * (1.f - z_vl);
  if (lane == 0) {
    o.dzpre[(i64)r * 3] = q_av;
    o.dzpre[(i64)r * 3 + 1] = q_al;
    o.dzpre[(i64)r * 3 + 2] = q_vl;
  }
  const float* xa = p.x[0] + (i64)r * D;
  const float* xv = p.x[1] + (i64)r * D;
  const float* xl = p.x[2] + (i64)r * D;
  const float* wav = p.w;
  const float* wal = p.w + 3 * D;
  const float* wvl = p.w + 6 * D;
  for (int k = lane; k < D; k += 32) {
    const float a = xa[k], v = xv[k], l = xl[k];
    o.dx[0][(i64)r * D + k] = q_av * (wav[k] + wav[2 * D + k] * v) + q_al * (wal[k] + wal[2 * D + k] * l);
    o.dx[1][(i64)r * D + k] = q_av * (wav[D + k] + wav[2 * D + k] * a) + q_vl * (wvl[k] + wvl[2 * D + k] * l);
    o.dx[2][(i64)r * D + k] = q_al * (wal[D + k] + wal[2 * D + k] * a) + q_vl * (wvl[D + k] + wvl[2 * D + k] * v);
  }
}

// dw[g][j] += sum_r dzpre[r][g] * feat_g(r, j) over this block's row chunk; db[g] += sum_r dzpre[r][g]
__global__ void gated_wgrad_kernel(int N, int D, const float* __restrict__ xa, const float* __restrict__ xv,
                                   const float* __restrict__ xl, const float* __restrict__ dzpre, int rows_per_block,
                                   float* __restrict__ dw, float* __restrict__ db) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  const int r0 = blockIdx.z * rows_per_block, r1 = min(N, r0 + rows_per_block);
  const float* xs[3] = {xa, xv, xl};
  const float* x1 = xs[gt_first(g)];
  const float* x2 = xs[gt_second(g)];
  if (j < 3 * D) {
    const int part = j / D, k = j - part * D;
    float acc = 0.f;
    for (int r = r0; r < r1; r++) {
      const float q = dzpre[(i64)r * 3 + g];
      const float f = part == 0 ? x1[(i64)r * D + k] : (part == 1 ? x2[(i64)r * D + k] : x1[(i64)r * D + k] * x2[(i64)r * D + k]);
      acc = fmaf(q, f, acc);
    }
    atomicAdd(dw + (i64)g * 3 * D + j, acc);
  }
  if (j == 0) {
    float acc = 0.f;
    for (int r = r0; r < r1; r++) acc += dzpre[(i64)r * 3 + g];
    atomicAdd(db + g, acc);
  }
}

// y = mask ? x * scale : 0   (uint8 keep mask; in == out allowed)
__global__ void mask_scale_kernel(i64 n, const float* __restrict__ x, const unsigned char* __restrict__ m, float scale,
                                  float* __restrict__ y) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = m[i] ? x[i] * scale : 0.f;
}

}  // namespace mmdfn

using namespace mmdfn;

extern "C" int mmdfn_mask_scale(long long n, const float* x, const unsigned char* mask, float scale, float* y,
                                void* stream) {
  if (!x || !mask || !y) return MMDFN_ENULL;
  if (n <= 0) return 0;
  mask_scale_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(n, x, mask, scale, y);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_gated_fuse_fwd(int N, int D, int C, const float* xa, const float* xv, const float* xl,
                                    const float* Pa, const float* Pv, const float* Pl, const float* w, const float* b,
                                    float* out, float* z, void* stream) {
  if (!xa || !xv || !xl || !Pa || !Pv || !Pl || !w || !b || !out || !z) return MMDFN_ENULL;
  if (N < 0 || D <= 0 || C <= 0) return MMDFN_EINVAL;
  if (N == 0) return 0;
  GatedArgs a{N, D, C, {xa, xv, xl}, {Pa, Pv, Pl}, w, b};
  gated_fuse_fwd_kernel<<<ceil_div(N, GT_WARPS), 32 * GT_WARPS, 0, (cudaStream_t)stream>>>(a, out, z);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// dw (3, 3D) and db (3) are overwritten.
extern "C" int mmdfn_gated_fuse_bwd(int N, int D, int C, const float* dout, const float* xa, const float* xv,
                                    const float* xl, const float* Pa, const float* Pv, const float* Pl, const float* w,
                                    const float* b, const float* z, float* dPa, float* dPv, float* dPl, float* dxa,
                                    float* dxv, float* dxl, float* dw, float* db, float* dzpre_ws, void* stream) {
  if (!dout || !xa || !xv || !xl || !Pa || !Pv || !Pl || !w || !b || !z || !dPa || !dPv || !dPl || !dxa || !dxv || !dxl ||
      !dw || !db || !dzpre_ws)
    return MMDFN_ENULL;
  if (N < 0 || D <= 0 || C <= 0) return MMDFN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  MMDFN_CUDA(cudaMemsetAsync(dw, 0, (size_t)9 * D * sizeof(float), st));
  MMDFN_CUDA(cudaMemsetAsync(db, 0, 3 * sizeof(float), st));
  if (N == 0) return 0;
  GatedArgs a{N, D, C, {xa, xv, xl}, {Pa, Pv, Pl}, w, b};
  GatedBwdOut o{{dPa, dPv, dPl}, {dxa, dxv, dxl}, dzpre_ws};
  gated_fuse_bwd_kernel<<<ceil_div(N, GT_WARPS), 32 * GT_WARPS, 0, st>>>(a, dout, z, o);
  MMDFN_LAUNCH_CHECK();
  const int rpb = 128;
  gated_wgrad_kernel<<<dim3(ceil_div(3 * D, 128), 3, ceil_div(N, rpb)), 128, 0, st>>>(N, D, xa, xv, xl, dzpre_ws, rpb, dw, db);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
