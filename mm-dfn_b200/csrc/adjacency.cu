// k5: block-compact multimodal adjacency (forward + backward).
// Replaces MM_GCN.create_big_adj (code/model_mm.py:122-180): instead of a dense (3N)^2
// matrix and two dense diagonal GEMMs, each dialogue owns 3 in-modal L x L blocks and 3
// cross-modal diagonals; normalisation D^-1/2 S D^-1/2 is applied in place.
//
// HBM layout (all fp32):
//   adj_blk  : sum_b 3*L_b^2 floats; block (b, m) at blk_off[b] + m*L_b^2, row-major L_b x L_b
//   adj_diag : (3, N)  pairs p = 0:(a,v) 1:(a,l) 2:(v,l); value shared by A[(m,r),(n,r)] and its transpose
//   dinv     : (3N) d^-1/2 ; rinv : (3N) 1/||x|| ; cos_blk / cos_diag : scaled cosines c' (saved for backward)
#include "gemm_tile.cuh"
#include "internal.cuh"
#include "../../include/mmdfn_b200.h"

namespace mmdfn {

constexpr int FH = 200;                 // feature width of the graph input (2*D_e)
constexpr float COS_SCALE = 0.99999f;   // code/model_mm.py:149,165
constexpr float INV_PI = 0.31830988618379067154f;

__device__ __forceinline__ float angular(float c) { return 1.0f - acosf(c) * INV_PI; }
// d/dc' of (1 - acos(c')/pi)
__device__ __forceinline__ float angular_grad(float c) { return INV_PI / sqrtf(1.0f - c * c); }

__device__ __forceinline__ int pair_of(int m, int n) { return m + n - 1; }   // (0,1)->0 (0,2)->1 (1,2)->2

// warp per utterance: row norms of the three modalities, the three cross-modal cosines,
// and the cross-modal part of the degree.
__global__ void adj_rownorm_kernel(int N, const float* __restrict__ X, float modal_weight, float* __restrict__ rinv,
                                   float* __restrict__ cos_diag, float* __restrict__ adj_diag,
                                   float* __restrict__ deg) {
  const int n = blockIdx.x * blockDim.y + threadIdx.y;
  if (n >= N) return;
  const int lane = threadIdx.x;
  float ss[3] = {0.f, 0.f, 0.f}, dt[3] = {0.f, 0.f, 0.f};
  for (int c = lane; c < FH; c += 32) {
    const float a = X[(i64)n * FH + c], v = X[((i64)N + n) * FH + c], l = X[((i64)2 * N + n) * FH + c];
    ss[0] = fmaf(a, a, ss[0]); ss[1] = fmaf(v, v, ss[1]); ss[2] = fmaf(l, l, ss[2]);
    dt[0] = fmaf(a, v, dt[0]); dt[1] = fmaf(a, l, dt[1]); dt[2] = fmaf(v, l, dt[2]);
  }
#pragma unroll
  for (int i = 0; i < 3; i++) { ss[i] = warp_sum(ss[i]); dt[i] = warp_sum(dt[i]); }
  if (lane == 0) {
    const float ia = 1.0f / sqrtf(ss[0]), iv = 1.0f / sqrtf(ss[1]), il = 1.0f / sqrtf(ss[2]);
    rinv[n] = ia; rinv[N + n] = iv; rinv[2 * N + n] = il;
    const float c0 = dt[0] * ia * iv * COS_SCALE, c1 = dt[1] * ia * il * COS_SCALE, c2 = dt[2] * iv * il * COS_SCALE;
    cos_diag[n] = c0; cos_diag[N + n] = c1; cos_diag[2 * N + n] = c2;
    const float s0 = angular(c0) * modal_weight, s1 = angular(c1) * modal_weight, s2 = angular(c2) * modal_weight;
    adj_diag[n] = s0; adj_diag[N + n] = s1; adj_diag[2 * N + n] = s2;
    deg[n] = s0 + s1; deg[N + n] = s0 + s2; deg[2 * N + n] = s1 + s2;
  }
}

struct AdjGeom {
  int B, N;
  const int* dia_off;
  const i64* blk_off;
};

// grouped Gram: one 64x64 tile of the L x L cosine block of (dialogue, modality)
__global__ void __launch_bounds__(GEMM_THREADS) adj_gram_kernel(AdjGeom g, const float* __restrict__ X,
                                                                const float* __restrict__ rinv,
                                                                float* __restrict__ cos_blk, float* __restrict__ adj_blk,
                                                                float* __restrict__ deg) {
  __shared__ __align__(16) float smem[GemmSmem<64, 64>::FLOATS];
  const int b = blockIdx.z / 3, m = blockIdx.z % 3;
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  if (m0 >= L || n0 >= L) return;
  const float* Xm = X + ((i64)m * g.N + off) * FH;
  float acc[4][4];
  zero_acc(acc);
  gemm_tile_accum<64, 64, 4, 4, false, true>(Xm, FH, Xm, FH, L, L, m0, n0, 0, FH, acc, smem);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const float* ri = rinv + (i64)m * g.N + off;
  const i64 base = g.blk_off[b] + (i64)m * L * L;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = m0 + ty * 4 + i;
    float rs = 0.f;
    if (r < L) {
      const float ir = ri[r];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = n0 + tx * 4 + j;
        if (c < L) {
          const float cs = acc[i][j] * ir * ri[c] * COS_SCALE;
          const float s = angular(cs);
          cos_blk[base + (i64)r * L + c] = cs;
          adj_blk[base + (i64)r * L + c] = s;
          rs += s;
        }
      }
    }
    // reduce over the 16 threads (tx) that share this row: they are 16 consecutive lanes
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
    if (tx == 0 && r < L) atomicAdd(deg + (i64)m * g.N + off + r, rs);
  }
}

__global__ void adj_dinv_kernel(int n3, const float* __restrict__ deg, float* __restrict__ dinv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n3) dinv[i] = 1.0f / sqrtf(deg[i]);
}

// A_hat = dinv_r * S * dinv_c, in place; blockIdx.y = (dialogue, modality)
__global__ void adj_scale_kernel(AdjGeom g, const float* __restrict__ dinv, float* __restrict__ adj_blk) {
  const int b = blockIdx.y / 3, m = blockIdx.y % 3;
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const i64 base = g.blk_off[b] + (i64)m * L * L;
  const float* di = dinv + (i64)m * g.N + off;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < (i64)L * L; idx += (i64)gridDim.x * blockDim.x) {
    const int r = (int)(idx / L), c = (int)(idx - (i64)r * L);
    adj_blk[base + idx] *= di[r] * di[c];
  }
}

__global__ void adj_scale_diag_kernel(int N, const float* __restrict__ dinv, float* __restrict__ adj_diag) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float da = dinv[n], dv = dinv[N + n], dl = dinv[2 * N + n];
  adj_diag[n] *= da * dv;
  adj_diag[N + n] *= da * dl;
  adj_diag[2 * N + n] *= dv * dl;
}

// ---------------------------------------------------------------------------------------
// y = A_hat x  (the message aggregate of GraphConvolution, code/model_GCN.py:178) and, since
// A_hat is symmetric, also its transpose product in the backward pass.
// One CTA = 64 rows x 128 (>= G) columns of one (dialogue, modality) block.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS) adj_spmm_kernel(AdjGeom g, const float* __restrict__ adj_blk,
                                                                const float* __restrict__ adj_diag,
                                                                const float* __restrict__ x, int G,
                                                                float* __restrict__ y) {
  __shared__ __align__(16) float smem[GemmSmem<64, 128>::FLOATS];
  const int b = blockIdx.z / 3, m = blockIdx.z % 3;
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 128;
  if (m0 >= L) return;
  const float* A = adj_blk + g.blk_off[b] + (i64)m * L * L;
  const float* Bx = x + ((i64)m * g.N + off) * G;
  float acc[4][8];
  zero_acc(acc);
  gemm_tile_accum<64, 128, 4, 8, false, false>(A, L, Bx, G, L, G, m0, n0, 0, L, acc, smem);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int o1 = (m == 0) ? 1 : 0, o2 = (m == 2) ? 1 : 2;       // the two other modalities
  const int p1 = pair_of(min(m, o1), max(m, o1)), p2 = pair_of(min(m, o2), max(m, o2));
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = m0 + ty * 4 + i;
    if (r >= L) continue;
    const float d1 = adj_diag[(i64)p1 * g.N + off + r], d2 = adj_diag[(i64)p2 * g.N + off + r];
    const float* x1 = x + ((i64)o1 * g.N + off + r) * G;
    const float* x2 = x + ((i64)o2 * g.N + off + r) * G;
    float* yr = y + ((i64)m * g.N + off + r) * G;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int c = n0 + tile_col<128, 8>(tx, j);
      if (c < G) yr[c] = acc[i][j] + d1 * x1[c] + d2 * x2[c];
    }
  }
}

// ---------------------------------------------------------------------------------------
// G = 100 specialisation of the aggregate (the model's graph hidden size).  One CTA = 40 rows x 100 columns of one
// (dialogue, modality) block: 250 active threads, thread (ty, tx) owns rows 4ty..4ty+3 x columns 4tx..4tx+3, so no
// lane computes padded columns.  Operands come in K chunks of up to 128 through cp.async (one chunk = the whole
// block for L <= 128, double-buffered for longer dialogues): A rows are read as warp-broadcast 128-bit loads, the
// z rows as conflict-free 128-bit loads, 64 FFMA per 8 shared loads.  288 CTAs for the 32 x 100 shard = 2 per SM.
// ---------------------------------------------------------------------------------------
constexpr int SP_ROWS = 40, SP_G = 100, SP_KCH = 128, SP_LDA = SP_KCH + 4;
constexpr int SP_STAGE_FLOATS = SP_ROWS * SP_LDA + SP_KCH * SP_G;
constexpr int SP_SMEM = 2 * SP_STAGE_FLOATS * 4;

__device__ __forceinline__ void sp_cp16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void sp_cp4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(256) adj_spmm100_kernel(AdjGeom g, const float* __restrict__ adj_blk,
                                                          const float* __restrict__ adj_diag, const float* __restrict__ x,
                                                          float* __restrict__ y) {
  extern __shared__ __align__(16) float sp_smem[];
  const int b = blockIdx.z / 3, m = blockIdx.z % 3;
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const int row0 = blockIdx.x * SP_ROWS;
  if (row0 >= L) return;
  const int tid = threadIdx.x;
  const int nrows = min(SP_ROWS, L - row0);
  const float* A = adj_blk + g.blk_off[b] + (i64)m * L * L + (i64)row0 * L;
  const float* Bx = x + ((i64)m * g.N + off) * SP_G;
  const bool a_vec = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool b_vec = (reinterpret_cast<uintptr_t>(Bx) & 15) == 0;
  const int nchunks = (L + SP_KCH - 1) / SP_KCH;

  auto load_chunk = [&](int c, int buf) {
    float* As = sp_smem + buf * SP_STAGE_FLOATS;
    float* Bs = As + SP_ROWS * SP_LDA;
    const int k0 = c * SP_KCH, kc = min(SP_KCH, L - k0), kc4 = (kc + 3) & ~3;
    if (a_vec) {
      const int q = kc >> 2;                                    // L % 4 == 0 -> kc % 4 == 0
      for (int i = tid; i < nrows * q; i += 256) {
        const int r = i / q, c4 = i - r * q;
        sp_cp16(As + r * SP_LDA + 4 * c4, A + (i64)r * L + k0 + 4 * c4);
      }
    } else {
      for (int i = tid; i < nrows * kc; i += 256) {
        const int r = i / kc, k = i - r * kc;
        sp_cp4(As + r * SP_LDA + k, A + (i64)r * L + k0 + k);
      }
      for (int i = tid; i < nrows * (kc4 - kc); i += 256) {      // zero the k tail up to a multiple of 4
        const int r = i / (kc4 - kc), k = kc + i - r * (kc4 - kc);
        As[r * SP_LDA + k] = 0.f;
      }
    }
    for (int i = tid; i < (SP_ROWS - nrows) * kc4; i += 256) {   // rows beyond the block: zeros
      const int r = nrows + i / kc4, k = i % kc4;
      As[r * SP_LDA + k] = 0.f;
    }
    const float* bsrc = Bx + (i64)k0 * SP_G;                     // kc consecutive rows of x are one contiguous run
    if (b_vec) {
      for (int i = tid; i < kc * (SP_G / 4); i += 256) sp_cp16(Bs + 4 * i, bsrc + 4 * i);
    } else {
      for (int i = tid; i < kc * SP_G; i += 256) sp_cp4(Bs + i, bsrc + i);
    }
    for (int i = tid; i < (kc4 - kc) * SP_G; i += 256) Bs[kc * SP_G + i] = 0.f;
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int tx = tid % 25, ty = tid / 25;                        // ty 0..9 (tid < 250), warps straddle at most 2 row groups
  const bool active = tid < 250;
  float acc[4][4];
  zero_acc(acc);
  load_chunk(0, 0);
  for (int c = 0; c < nchunks; c++) {
    const int buf = c & 1;
    if (c + 1 < nchunks) {
      load_chunk(c + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (active) {
      const float* As = sp_smem + buf * SP_STAGE_FLOATS + (ty * 4) * SP_LDA;
      const float* Bs = sp_smem + buf * SP_STAGE_FLOATS + SP_ROWS * SP_LDA + tx * 4;
      const int kc4 = (min(SP_KCH, L - c * SP_KCH) + 3) & ~3;
#pragma unroll 2
      for (int k = 0; k < kc4; k += 4) {
        float4 a[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = *reinterpret_cast<const float4*>(As + i * SP_LDA + k);
#pragma unroll
        for (int kk = 0; kk < 4; kk++) bb[kk] = *reinterpret_cast<const float4*>(Bs + (k + kk) * SP_G);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            acc[i][0] = fmaf(av[kk], bb[kk].x, acc[i][0]);
            acc[i][1] = fmaf(av[kk], bb[kk].y, acc[i][1]);
            acc[i][2] = fmaf(av[kk], bb[kk].z, acc[i][2]);
            acc[i][3] = fmaf(av[kk], bb[kk].w, acc[i][3]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (!active) return;
  const int o1 = (m == 0) ? 1 : 0, o2 = (m == 2) ? 1 : 2;
  const int p1 = pair_of(min(m, o1), max(m, o1)), p2 = pair_of(min(m, o2), max(m, o2));
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = row0 + ty * 4 + i;
    if (r >= L) continue;
    const float d1 = adj_diag[(i64)p1 * g.N + off + r], d2 = adj_diag[(i64)p2 * g.N + off + r];
    const float4 x1 = *reinterpret_cast<const float4*>(x + ((i64)o1 * g.N + off + r) * SP_G + tx * 4);
    const float4 x2 = *reinterpret_cast<const float4*>(x + ((i64)o2 * g.N + off + r) * SP_G + tx * 4);
    float4 o;
    o.x = acc[i][0] + d1 * x1.x + d2 * x2.x;
    o.y = acc[i][1] + d1 * x1.y + d2 * x2.y;
    o.z = acc[i][2] + d1 * x1.z + d2 * x2.z;
    o.w = acc[i][3] + d1 * x1.w + d2 * x2.w;
    *reinterpret_cast<float4*>(y + ((i64)m * g.N + off + r) * SP_G + tx * 4) = o;
  }
}

// 0 = automatic: the any-length tcgen05 kernel (spmm_tc_long.cu; validated on B200 in round 2 against fp64 products for
// L = 1..500, profiles/r02_spmm_variant_long_kernel.log: as fast as the whole-block kernel at L <= 128, 2.6-3.7x the FFMA
// fallback at L = 200..500); 1 = FFMA kernels only; 2 = the whole-block tcgen05 kernel (spmm_tc.cu) when every dialogue
// has <= 128 utterances.  1 and 2 exist for A/B timing (tools/spmm_variant.py) and the cross-kernel parity tests.
static int g_spmm_variant = 0;

int adj_spmm(int B, int N, int Lmax, const int* dia_off, const i64* blk_off, const float* adj_blk,
             const float* adj_diag, const float* x, int G, float* y, cudaStream_t st) {
  if (B <= 0 || N <= 0 || Lmax <= 0) return 0;
  AdjGeom g{B, N, dia_off, blk_off};
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (G == SP_G && aligned && Lmax <= 128 && g_spmm_variant == 2)
    return adj_spmm_tc(B, N, dia_off, blk_off, adj_blk, adj_diag, x, y, st);
  if (G == SP_G && aligned && g_spmm_variant != 1)
    return adj_spmm_tc_long(B, N, Lmax, dia_off, blk_off, adj_blk, adj_diag, x, y, st);
  if (G == SP_G && aligned) {
    static bool configured = false;
    if (!configured) {
      MMDFN_CUDA(cudaFuncSetAttribute(adj_spmm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
      configured = true;
    }
    // one operand stage when every block fits a single K chunk (3 CTAs per SM), two (double buffering) otherwise
    const int smem = (Lmax > SP_KCH ? 2 : 1) * SP_STAGE_FLOATS * 4;
    adj_spmm100_kernel<<<dim3(ceil_div(Lmax, SP_ROWS), 1, B * 3), 256, smem, st>>>(g, adj_blk, adj_diag, x, y);
    MMDFN_LAUNCH_CHECK();
    return 0;
  }
  adj_spmm_kernel<<<dim3(ceil_div(Lmax, 64), ceil_div(G, 128), B * 3), GEMM_THREADS, 0, st>>>(g, adj_blk, adj_diag, x, G, y);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// dA_blk[r,j] (+)= dhi_r . z_j   (true gradient of y = A_hat z w.r.t. the stored block entries)
__global__ void __launch_bounds__(GEMM_THREADS) adj_grad_blk_kernel(AdjGeom g, const float* __restrict__ dhi,
                                                                    const float* __restrict__ z, int G,
                                                                    float* __restrict__ p_blk, int accumulate) {
  __shared__ __align__(16) float smem[GemmSmem<64, 64>::FLOATS];
  const int b = blockIdx.z / 3, m = blockIdx.z % 3;
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  if (m0 >= L || n0 >= L) return;
  const float* A = dhi + ((i64)m * g.N + off) * G;
  const float* Bz = z + ((i64)m * g.N + off) * G;
  float acc[4][4];
  zero_acc(acc);
  gemm_tile_accum<64, 64, 4, 4, false, true>(A, G, Bz, G, L, L, m0, n0, 0, G, acc, smem);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float* P = p_blk + g.blk_off[b] + (i64)m * L * L;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = m0 + ty * 4 + i;
    if (r >= L) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int c = n0 + tx * 4 + j;
      if (c < L) {
        float* q = P + (i64)r * L + c;
        *q = accumulate ? (*q + acc[i][j]) : acc[i][j];
      }
    }
  }
}

// dA_diag[p][n] (+)= dhi_m[n].z_o[n] + dhi_o[n].z_m[n]   (the stored value serves both orientations)
__global__ void adj_grad_diag_kernel(int N, const float* __restrict__ dhi, const float* __restrict__ z, int G,
                                     float* __restrict__ p_diag, int accumulate) {
  const int n = blockIdx.x * blockDim.y + threadIdx.y;
  if (n >= N) return;
  const int lane = threadIdx.x;
  float s[3] = {0.f, 0.f, 0.f};
  for (int c = lane; c < G; c += 32) {
    const float ga = dhi[(i64)n * G + c], gv = dhi[((i64)N + n) * G + c], gl = dhi[((i64)2 * N + n) * G + c];
    const float za = z[(i64)n * G + c], zv = z[((i64)N + n) * G + c], zl = z[((i64)2 * N + n) * G + c];
    s[0] += ga * zv + gv * za;
    s[1] += ga * zl + gl * za;
    s[2] += gv * zl + gl * zv;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) s[i] = warp_sum(s[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float* q = p_diag + (i64)i * N + n;
      *q = accumulate ? (*q + s[i]) : s[i];
    }
  }
}

int adj_grad_accum(int B, int N, int Lmax, const int* dia_off, const i64* blk_off, const float* dhi,
                   const float* z, int G, float* p_blk, float* p_diag, int accumulate, cudaStream_t st) {
  if (B <= 0 || N <= 0 || Lmax <= 0) return 0;
  AdjGeom g{B, N, dia_off, blk_off};
  const int nt = ceil_div(Lmax, 64);
  adj_grad_blk_kernel<<<dim3(nt, nt, B * 3), GEMM_THREADS, 0, st>>>(g, dhi, z, G, p_blk, accumulate);
  MMDFN_LAUNCH_CHECK();
  adj_grad_diag_kernel<<<ceil_div(N, 8), dim3(32, 8), 0, st>>>(N, dhi, z, G, p_diag, accumulate);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------
// backward of the adjacency construction
// ---------------------------------------------------------------------------------------
// P <- P + P^T per block, in place (32x32 tile pairs through shared memory)
__global__ void adj_symmetrize_kernel(AdjGeom g, float* __restrict__ p_blk) {
  __shared__ float ta[32][33], tb[32][33];
  const int b = blockIdx.z / 3, m = blockIdx.z % 3;
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (tj < ti || ti * 32 >= L || tj * 32 >= L) return;
  float* P = p_blk + g.blk_off[b] + (i64)m * L * L;
  const int x = threadIdx.x;
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int ra = ti * 32 + y, ca = tj * 32 + x;
    ta[y][x] = (ra < L && ca < L) ? P[(i64)ra * L + ca] : 0.f;
    const int rb = tj * 32 + y, cb = ti * 32 + x;
    tb[y][x] = (rb < L && cb < L) ? P[(i64)rb * L + cb] : 0.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int ra = ti * 32 + y, ca = tj * 32 + x;
    if (ra < L && ca < L) P[(i64)ra * L + ca] = ta[y][x] + tb[x][y];
    if (ti != tj) {
      const int rb = tj * 32 + y, cb = ti * 32 + x;
      if (rb < L && cb < L) P[(i64)rb * L + cb] = tb[y][x] + ta[x][y];
    }
  }
}

// dd_u = -1/2 * dinv_u^2 * ( sum_j Psym[r,j] A[r,j] + sum_pairs Pdiag A_diag )   warp per row
__global__ void adj_bwd_rowterm_kernel(AdjGeom g, const float* __restrict__ p_blk, const float* __restrict__ p_diag,
                                       const float* __restrict__ adj_blk, const float* __restrict__ adj_diag,
                                       const float* __restrict__ dinv, float* __restrict__ dd) {
  const int b = blockIdx.y / 3, m = blockIdx.y % 3;
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const int r = blockIdx.x * blockDim.y + threadIdx.y;
  if (r >= L) return;
  const i64 base = g.blk_off[b] + (i64)m * L * L + (i64)r * L;
  float s = 0.f;
  for (int c = threadIdx.x; c < L; c += 32) s = fmaf(p_blk[base + c], adj_blk[base + c], s);
  s = warp_sum(s);
  if (threadIdx.x == 0) {
    const int o1 = (m == 0) ? 1 : 0, o2 = (m == 2) ? 1 : 2;
    const int p1 = pair_of(min(m, o1), max(m, o1)), p2 = pair_of(min(m, o2), max(m, o2));
    const i64 n = off + r;
    s += p_diag[(i64)p1 * g.N + n] * adj_diag[(i64)p1 * g.N + n] + p_diag[(i64)p2 * g.N + n] * adj_diag[(i64)p2 * g.N + n];
    const float di = dinv[(i64)m * g.N + n];
    dd[(i64)m * g.N + n] = -0.5f * di * di * s;
  }
}

// W[r,j] = COS_SCALE * f'(c'[r,j]) * (Psym[r,j] dinv_r dinv_j + dd_r + dd_j) * rinv_j   (in place on P)
__global__ void adj_bwd_weights_kernel(AdjGeom g, float* __restrict__ p_blk, const float* __restrict__ cos_blk,
                                       const float* __restrict__ dinv, const float* __restrict__ dd,
                                       const float* __restrict__ rinv) {
  const int b = blockIdx.y / 3, m = blockIdx.y % 3;
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const i64 base = g.blk_off[b] + (i64)m * L * L;
  const i64 nb = (i64)m * g.N + off;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < (i64)L * L; idx += (i64)gridDim.x * blockDim.x) {
    const int r = (int)(idx / L), c = (int)(idx - (i64)r * L);
    const float ds = p_blk[base + idx] * dinv[nb + r] * dinv[nb + c] + dd[nb + r] + dd[nb + c];
    p_blk[base + idx] = COS_SCALE * angular_grad(cos_blk[base + idx]) * ds * rinv[nb + c];
  }
}

// dxh (L x 200) = W (L x L) X_m (L x 200)    grouped, 64 x 64 tiles
__global__ void __launch_bounds__(GEMM_THREADS) adj_bwd_gemm_kernel(AdjGeom g, const float* __restrict__ w_blk,
                                                                    const float* __restrict__ X,
                                                                    float* __restrict__ dxh) {
  __shared__ __align__(16) float smem[GemmSmem<64, 64>::FLOATS];
  const int b = blockIdx.z / 3, m = blockIdx.z % 3;
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  if (m0 >= L) return;
  const float* A = w_blk + g.blk_off[b] + (i64)m * L * L;
  const float* Bx = X + ((i64)m * g.N + off) * FH;
  float acc[4][4];
  zero_acc(acc);
  gemm_tile_accum<64, 64, 4, 4, false, false>(A, L, Bx, FH, L, FH, m0, n0, 0, L, acc, smem);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = m0 + ty * 4 + i;
    if (r >= L) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int c = n0 + tx * 4 + j;
      if (c < FH) dxh[((i64)m * g.N + off + r) * FH + c] = acc[i][j];
    }
  }
}

// warp per utterance: add the cross-modal terms to d(x_hat), project out the radial part,
// scale by 1/||x||, (optionally) add an incoming gradient:  out = add + rinv*(dxh - xh (xh.dxh))
__global__ void adj_bwd_finish_kernel(int N, const float* __restrict__ X, const float* __restrict__ rinv,
                                      const float* __restrict__ cos_diag, const float* __restrict__ p_diag,
                                      const float* __restrict__ dinv, const float* __restrict__ dd, float modal_weight,
                                      const float* __restrict__ add, float* __restrict__ dxh_out) {
  const int n = blockIdx.x * blockDim.y + threadIdx.y;
  if (n >= N) return;
  const int lane = threadIdx.x;
  float ri[3], di[3], dq[3];
#pragma unroll
  for (int m = 0; m < 3; m++) { ri[m] = rinv[(i64)m * N + n]; di[m] = dinv[(i64)m * N + n]; dq[m] = dd[(i64)m * N + n]; }
  float qd[3];
  {
    const int pm[3] = {0, 0, 1}, pn[3] = {1, 2, 2};
#pragma unroll
    for (int p = 0; p < 3; p++) {
      const float ds = p_diag[(i64)p * N + n] * di[pm[p]] * di[pn[p]] + dq[pm[p]] + dq[pn[p]];
      qd[p] = modal_weight * COS_SCALE * angular_grad(cos_diag[(i64)p * N + n]) * ds;
    }
  }
  constexpr int PER = (FH + 31) / 32;   // 7
  float xh[3][PER], g[3][PER];
  float dotp[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < PER; k++) {
    const int c = lane + 32 * k;
#pragma unroll
    for (int m = 0; m < 3; m++) {
      xh[m][k] = (c < FH) ? X[((i64)m * N + n) * FH + c] * ri[m] : 0.f;
      g[m][k] = (c < FH) ? dxh_out[((i64)m * N + n) * FH + c] : 0.f;
    }
    // cross-modal: d xh_m += qd[p(m,o)] * xh_o
    const float xa = xh[0][k], xv = xh[1][k], xl = xh[2][k];
    g[0][k] += qd[0] * xv + qd[1] * xl;
    g[1][k] += qd[0] * xa + qd[2] * xl;
    g[2][k] += qd[1] * xa + qd[2] * xv;
#pragma unroll
    for (int m = 0; m < 3; m++) dotp[m] = fmaf(xh[m][k], g[m][k], dotp[m]);
  }
#pragma unroll
  for (int m = 0; m < 3; m++) dotp[m] = warp_sum(dotp[m]);
#pragma unroll
  for (int k = 0; k < PER; k++) {
    const int c = lane + 32 * k;
    if (c >= FH) continue;
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const i64 e = ((i64)m * N + n) * FH + c;
      float v = ri[m] * (g[m][k] - xh[m][k] * dotp[m]);
      if (add) v += add[e];
      dxh_out[e] = v;
    }
  }
}

// dense (3N x 3N) materialisation for API compatibility (not on the hot path)
__global__ void adj_densify_kernel(AdjGeom g, const float* __restrict__ adj_blk, const float* __restrict__ adj_diag,
                                   float* __restrict__ dense) {
  const int b = blockIdx.y / 3, m = blockIdx.y % 3;
  const int off = g.dia_off[b], L = g.dia_off[b + 1] - off;
  const i64 base = g.blk_off[b] + (i64)m * L * L;
  const i64 n3 = (i64)3 * g.N;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < (i64)L * L; idx += (i64)gridDim.x * blockDim.x) {
    const int r = (int)(idx / L), c = (int)(idx - (i64)r * L);
    dense[((i64)m * g.N + off + r) * n3 + (i64)m * g.N + off + c] = adj_blk[base + idx];
    if (r == c) {
      for (int o = 0; o < 3; o++) {
        if (o == m) continue;
        const int p = pair_of(min(m, o), max(m, o));
        dense[((i64)m * g.N + off + r) * n3 + (i64)o * g.N + off + r] = adj_diag[(i64)p * g.N + off + r];
      }
    }
  }
}

}  // namespace mmdfn

using namespace mmdfn;

extern "C" int mmdfn_adj_fwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* X,
                             float modal_weight, float* adj_blk, float* adj_diag, float* dinv, float* rinv,
                             float* cos_blk, float* cos_diag, float* deg_ws, void* stream) {
  if (!dia_off || !blk_off || !X || !adj_blk || !adj_diag || !dinv || !rinv || !cos_blk || !cos_diag || !deg_ws)
    return MMDFN_ENULL;
  if (B < 0 || N < 0 || Lmax < 0) return MMDFN_EINVAL;
  if (B == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  AdjGeom g{B, N, dia_off, (const i64*)blk_off};
  adj_rownorm_kernel<<<ceil_div(N, 8), dim3(32, 8), 0, st>>>(N, X, modal_weight, rinv, cos_diag, adj_diag, deg_ws);
  MMDFN_LAUNCH_CHECK();
  const int nt = ceil_div(Lmax, 64);
  adj_gram_kernel<<<dim3(nt, nt, B * 3), GEMM_THREADS, 0, st>>>(g, X, rinv, cos_blk, adj_blk, deg_ws);
  MMDFN_LAUNCH_CHECK();
  adj_dinv_kernel<<<ceil_div(3 * N, 256), 256, 0, st>>>(3 * N, deg_ws, dinv);
  MMDFN_LAUNCH_CHECK();
  const int gx = (int)(ceil_div64((i64)Lmax * Lmax, 256) < 64 ? ceil_div64((i64)Lmax * Lmax, 256) : 64);
  adj_scale_kernel<<<dim3(gx, B * 3), 256, 0, st>>>(g, dinv, adj_blk);
  MMDFN_LAUNCH_CHECK();
  adj_scale_diag_kernel<<<ceil_div(N, 256), 256, 0, st>>>(N, dinv, adj_diag);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// d_blk is overwritten (used as scratch).  dX = add + dL/dX through the adjacency.
extern "C" int mmdfn_adj_bwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* X,
                             float modal_weight, const float* adj_blk, const float* adj_diag, const float* dinv,
                             const float* rinv, const float* cos_blk, const float* cos_diag, float* d_blk,
                             const float* d_diag, const float* add, float* dX, float* dd_ws, void* stream) {
  if (!dia_off || !blk_off || !X || !adj_blk || !adj_diag || !dinv || !rinv || !cos_blk || !cos_diag || !d_blk ||
      !d_diag || !dX || !dd_ws)
    return MMDFN_ENULL;
  if (B <= 0 || N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  AdjGeom g{B, N, dia_off, (const i64*)blk_off};
  const int nt32 = ceil_div(Lmax, 32);
  adj_symmetrize_kernel<<<dim3(nt32, nt32, B * 3), dim3(32, 8), 0, st>>>(g, d_blk);
  MMDFN_LAUNCH_CHECK();
  adj_bwd_rowterm_kernel<<<dim3(ceil_div(Lmax, 8), B * 3), dim3(32, 8), 0, st>>>(g, d_blk, d_diag, adj_blk, adj_diag, dinv, dd_ws);
  MMDFN_LAUNCH_CHECK();
  const int gx = (int)(ceil_div64((i64)Lmax * Lmax, 256) < 64 ? ceil_div64((i64)Lmax * Lmax, 256) : 64);
  adj_bwd_weights_kernel<<<dim3(gx, B * 3), 256, 0, st>>>(g, d_blk, cos_blk, dinv, dd_ws, rinv);
  MMDFN_LAUNCH_CHECK();
  adj_bwd_gemm_kernel<<<dim3(ceil_div(Lmax, 64), ceil_div(FH, 64), B * 3), GEMM_THREADS, 0, st>>>(g, d_blk, X, dX);
  MMDFN_LAUNCH_CHECK();
  adj_bwd_finish_kernel<<<ceil_div(N, 4), dim3(32, 4), 0, st>>>(N, X, rinv, cos_diag, d_diag, dinv, dd_ws, modal_weight, add, dX);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_adj_spmm_set_variant(int variant) {
  if (variant < 0 || variant > 2) return MMDFN_EINVAL;
  mmdfn::g_spmm_variant = variant;
  return 0;
}

extern "C" int mmdfn_adj_spmm(int B, int N, int Lmax, const int* dia_off, const long long* blk_off,
                              const float* adj_blk, const float* adj_diag, const float* x, int G, float* y,
                              void* stream) {
  if (!dia_off || !blk_off || !adj_blk || !adj_diag || !x || !y) return MMDFN_ENULL;
  if (G <= 0) return MMDFN_EINVAL;
  return adj_spmm(B, N, Lmax, dia_off, (const i64*)blk_off, adj_blk, adj_diag, x, G, y, (cudaStream_t)stream);
}

extern "C" int mmdfn_adj_grad(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* dhi,
                              const float* z, int G, float* d_blk, float* d_diag, int accumulate, void* stream) {
  if (!dia_off || !blk_off || !dhi || !z || !d_blk || !d_diag) return MMDFN_ENULL;
  return adj_grad_accum(B, N, Lmax, dia_off, (const i64*)blk_off, dhi, z, G, d_blk, d_diag, accumulate,
                        (cudaStream_t)stream);
}

extern "C" int mmdfn_adj_densify(int B, int N, int Lmax, const int* dia_off, const long long* blk_off,
                                 const float* adj_blk, const float* adj_diag, float* dense, void* stream) {
  if (!dia_off || !blk_off || !adj_blk || !adj_diag || !dense) return MMDFN_ENULL;
  if (B <= 0 || N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  MMDFN_TRY(fill_zero(dense, (size_t)9 * N * N * sizeof(float), st));
  AdjGeom g{B, N, dia_off, (const i64*)blk_off};
  const int gx = (int)(ceil_div64((i64)Lmax * Lmax, 256) < 64 ? ceil_div64((i64)Lmax * Lmax, 256) : 64);
  adj_densify_kernel<<<dim3(gx, B * 3), 256, 0, st>>>(g, adj_blk, adj_diag, dense);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
