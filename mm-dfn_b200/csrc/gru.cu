// k2/k3: persistent GRU recurrence (forward + backward) and the 2-layer bidirectional
// orchestrator that replaces nn.GRU(200,100,num_layers=2,bidirectional=True)
// ("lstm_l" code/model.py:866,1132 and the shared "rnn_parties" :868,1082,1113,1146).
//
// Design: the input-gate GEMM is hoisted out of the time loop (one dense GEMM per layer
// over all rows).  One CTA owns NB sequences of one direction for the whole sequence:
// W_hh stays in registers (100 weights per thread, as packed pairs for fp32x2 FMAs) and
// the NB hidden states live in shared memory (forward: a row per thread + broadcast
// 128-bit loads; backward: a 25-wide slice of four outputs per thread + quad shuffles).  The T-step loop touches HBM only for the
// per-step gate rows (coalesced, prefetched before the mat-vec) and the outputs.  For the speaker-party encoder the per-step rows are fetched
// through `rowmap` (gather fused into the recurrence): the projected utterance table is
// multiplied by W_ih once per utterance instead of once per (speaker, position) slot.
#include "internal.cuh"
#include "../../include/mmdfn_b200.h"

namespace mmdfn {

// Gate nonlinearities of the time loop.  The pointwise phase is a latency chain on the step's critical path (~600 cycles
// per pass with expf / tanhf / IEEE divisions: ~150 dependent instructions); the exp2-based forms below are ~40 and
// accurate to ~2e-7 absolute (ex2.approx / rcp.approx: 2 ulp each), which the contractive recurrence does not amplify
// (checked against the fp32 oracle by the BiGRU parity tests at their unchanged tolerances).
__device__ __forceinline__ float gru_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float gru_tanh(float x) {
  const float t = __expf(-2.f * fabsf(x));                   // in (0, 1]: no overflow for any x
  return copysignf(__fdividef(1.f - t, 1.f + t), x);
}

constexpr int GH = 100;          // hidden size (D_e), fixed by the reference (code/run_train_erc.py:389)
constexpr int G3 = 300;
constexpr int GRU_THREADS = 320;

struct GruFwdArgs {
  int T, nseq;
  const float* xg;        // (rows, 600): [dir0: r z n | dir1: r z n], bias_ih included
  const int* rowmap;      // (T, nseq) row of xg per slot, -1 = zero input; nullptr = identity
  const float* b_ih[2];   // used for rowmap == -1 slots
  const float* w_hh[2];   // (300, 100)
  const float* b_hh[2];   // (300)
  float* y;               // (T, nseq, 200) [fwd | bwd]
  float* gates;           // (T, nseq, 2, 400) r z n hn, nullable
};

// Forward recurrence: thread j < 300 keeps row j of W_hh in registers (50 packed pairs: the mat-vec runs on fp32x2
// FMAs) and reads the NB hidden vectors as warp-broadcast 128-bit loads.  (The quad-sliced mapping the backward kernel
// uses was measured slower here -- 160 vs 122 us at NB = 2: the forward step ends in a serial reduce -> gate chain.)
template <int NB>
__global__ void __launch_bounds__(GRU_THREADS, 1) gru_fwd_kernel(GruFwdArgs p) {
  __shared__ __align__(16) float hs[NB][GH];
  __shared__ __align__(16) float pre[NB][4 * GH];
  const int tid = threadIdx.x;
  const int dir = blockIdx.y;
  const int s0 = blockIdx.x * NB;
  const int nb = min(NB, p.nseq - s0);
  const int j = tid < G3 ? tid : G3 - 1;     // threads 300..319 only help in the pointwise phase
  float2 w2[GH / 2];
  {
    const float* wr = p.w_hh[dir] + (i64)j * GH;
#pragma unroll
    for (int k = 0; k < GH / 2; k++) w2[k] = make_float2(wr[2 * k], wr[2 * k + 1]);
  }
  const float bh = p.b_hh[dir][j];
  const float bi = p.b_ih[dir] ? p.b_ih[dir][j] : 0.f;
  for (int i = tid; i < NB * GH; i += GRU_THREADS) (&hs[0][0])[i] = 0.f;
  __syncthreads();

  // software pipeline: gate rows of step s+1 and row indices of step s+2 are in flight while step s computes.  The
  // row index stays a raw 32-bit value until it is used one step later: any arithmetic on it here (even the widening
  // to 64 bits) makes the warp wait for the load on the spot (14 % of all stall samples in the first version).
  auto slot_of = [&](int step, int b) { return (i64)(dir ? (p.T - 1 - step) : step) * p.nseq + s0 + b; };
  auto row_of = [&](int step, int b) -> int {           // -2: no such slot, -1: zero input (bias only), >= 0: row of xg
    if (step >= p.T || b >= nb) return -2;
    const i64 slot = slot_of(step, b);
    return p.rowmap ? p.rowmap[slot] : (int)slot;
  };
  auto gate_of = [&](int row) { return row >= 0 ? p.xg[(i64)row * 600 + dir * G3 + j] : (row == -1 ? bi : 0.f); };
  float xv[NB], xn[NB];
  int rown[NB];
#pragma unroll
  for (int b = 0; b < NB; b++) {
    xv[b] = gate_of(row_of(0, b));
    rown[b] = row_of(1, b);
  }
#pragma unroll 1
  for (int step = 0; step < p.T; step++) {
    const int t = dir ? (p.T - 1 - step) : step;
#pragma unroll
    for (int b = 0; b < NB; b++) {
      xn[b] = gate_of(rown[b]);          // consumed at the top of the next step
      rown[b] = row_of(step + 2, b);
    }
    // NACC independent accumulator pairs per sequence: with one, the 50 dependent FFMA2 of a sequence form a single
    // latency chain (the compiler schedules sequence after sequence), which was the whole step time (~1040 cycles per
    // sequence); short interleaved chains make the mat-vec issue-bound instead
    constexpr int NACC = NB <= 2 ? 4 : 2;
    float2 acc2[NB][NACC];
#pragma unroll
    for (int b = 0; b < NB; b++) {
      acc2[b][0] = make_float2(bh, 0.f);
#pragma unroll
      for (int a = 1; a < NACC; a++) acc2[b][a] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < GH; k += 4) {
#pragma unroll
      for (int b = 0; b < NB; b++) {
        const float4 h4 = *reinterpret_cast<const float4*>(&hs[b][k]);
        float2& a2 = acc2[b][(k / 4) % NACC];
        a2 = __ffma2_rn(w2[k / 2], make_float2(h4.x, h4.y), a2);
        a2 = __ffma2_rn(w2[k / 2 + 1], make_float2(h4.z, h4.w), a2);
      }
    }
    float acc[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
      float2 t = acc2[b][0];
#pragma unroll
      for (int a = 1; a < NACC; a++) { t.x += acc2[b][a].x; t.y += acc2[b][a].y; }
      acc[b] = t.x + t.y;
    }
    if (tid < G3) {
#pragma unroll
      for (int b = 0; b < NB; b++) {
        if (tid < 2 * GH) {
          pre[b][tid] = xv[b] + acc[b];
        } else {
          pre[b][tid] = xv[b];
          pre[b][tid + GH] = acc[b];
        }
      }
    }
    __syncthreads();
    for (int idx = tid; idx < nb * GH; idx += GRU_THREADS) {
      const int b = idx / GH, u = idx - b * GH;
      const float r = gru_sigmoid(pre[b][u]);
      const float z = gru_sigmoid(pre[b][GH + u]);
      const float hn = pre[b][3 * GH + u];
      const float n = gru_tanh(pre[b][2 * GH + u] + r * hn);
      const float hnew = (1.0f - z) * n + z * hs[b][u];
      hs[b][u] = hnew;
      const i64 slot = (i64)t * p.nseq + s0 + b;
      p.y[slot * 200 + dir * GH + u] = hnew;
      if (p.gates) {
        float* g = p.gates + (slot * 2 + dir) * 400;
        g[u] = r; g[GH + u] = z; g[2 * GH + u] = n; g[3 * GH + u] = hn;
      }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; b++) xv[b] = xn[b];
  }
}

// Mat-vec mapping of the backward recurrence.  A thread owns a 25-wide k slice of FOUR outputs (100 weights in
// registers as 4 x 14 packed pairs, pads zero) and reads only its slice of the gate gradients (7 128-bit loads per
// sequence instead of 25); the four lanes of a quad hold the four slices of the same outputs and combine their partial
// sums with a 3-shuffle reduce-scatter, after which lane s of the quad owns finished output s (measured 171 vs 188 us).
constexpr int GSL = 28;                 // padded slice length in shared memory (25 used + 3 zero)
constexpr int GHP = 4 * GSL;            // padded 100-vector: slice s at [28 s, 28 s + 25)
__device__ __forceinline__ int gru_hpos(int u) { return (u / 25) * GSL + (u % 25); }

// sums v[i] over the 4 lanes of a quad; returns the total of row (lane & 3) in that lane
__device__ __forceinline__ float quad_reduce_scatter(const float (&v)[4], int s) {
  const bool b0 = s & 1, b1 = s & 2;
  // xor 1: lanes with bit0 = 0 keep rows 0, 2 and receive the partner's; bit0 = 1 keep rows 1, 3
  const float send0 = b0 ? v[0] : v[1], send1 = b0 ? v[2] : v[3];
  const float keep0 = b0 ? v[1] : v[0], keep1 = b0 ? v[3] : v[2];
  const float x0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 1);      // row b0       (0 or 1)
  const float x1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 1);      // row 2 + b0   (2 or 3)
  // xor 2: bit1 = 0 keeps x0, bit1 = 1 keeps x1
  const float send = b1 ? x0 : x1, keep = b1 ? x1 : x0;
  return keep + __shfl_xor_sync(0xffffffffu, send, 2);                  // row b0 + 2 b1 = s
}

struct GruBwdArgs {
  int T, nseq;
  const float* dy;      // (T, nseq, 200)
  const float* y;       // (T, nseq, 200)
  const float* gates;   // (T, nseq, 2, 400)
  const float* w_hh[2];
  float* dxg;           // (T, nseq, 2, 300)  d/d(input gates)     = [dr_pre dz_pre dn_pre]
  float* dgh;           // (T, nseq, 2, 300)  d/d(W_hh h + b_hh)   = [dr_pre dz_pre dn_pre*r]
  float* db_ih[2];      // (300) per direction, pre-zeroed: column sums of dxg, accumulated here (atomics)
  float* db_hh[2];      // (300) per direction, pre-zeroed: column sums of dgh
};

template <int NB>
__global__ void __launch_bounds__(GRU_THREADS, 1) gru_bwd_kernel(GruBwdArgs p) {
  __shared__ __align__(16) float dh[NB][GH];
  __shared__ __align__(16) float dgh[NB][3 * GHP];      // gate segment jp at [112 jp + gru_hpos(jj)], pads stay zero
  __shared__ __align__(16) float part[3][NB][GH];
  const int tid = threadIdx.x;
  const int dir = blockIdx.y;
  const int s0 = blockIdx.x * NB;
  const int nb = min(NB, p.nseq - s0);
  // W_hh^T dgh with the same quad mapping as the forward kernel: thread (jp, uq, sl) holds the 25-wide jj slice sl of
  // gate segment jp for the four outputs u = uq + 25 i; lane sl of the quad finishes output uq + 25 sl
  const bool mv_on = tid < G3;
  const int jp = mv_on ? tid / GH : 2;
  const int sl = tid & 3, uq = mv_on ? (tid % GH) >> 2 : 0;
  const int u_f = uq + 25 * sl;
  float2 w2[4][GSL / 2];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int u = uq + 25 * i;
#pragma unroll
    for (int k = 0; k < GSL / 2; k++) {
      const float* wc = p.w_hh[dir] + (i64)(jp * GH + 25 * sl) * GH + u;
      w2[i][k] = make_float2((mv_on && 2 * k < 25) ? wc[(i64)(2 * k) * GH] : 0.f,
                             (mv_on && 2 * k + 1 < 25) ? wc[(i64)(2 * k + 1) * GH] : 0.f);
    }
  }
  for (int i = tid; i < NB * GH; i += GRU_THREADS) (&dh[0][0])[i] = 0.f;
  for (int i = tid; i < 3 * NB * GH; i += GRU_THREADS) (&part[0][0][0])[i] = 0.f;
  for (int i = tid; i < NB * 3 * GHP; i += GRU_THREADS) (&dgh[0][0])[i] = 0.f;
  __syncthreads();

  constexpr int ITEMS = (NB * GH + GRU_THREADS - 1) / GRU_THREADS;
  struct Pre { float r, z, n, hn, hp, dy; };
  Pre cur[ITEMS], nxt[ITEMS];
  auto fetch = [&](int step, Pre (&o)[ITEMS]) {
    const int t = dir ? step : (p.T - 1 - step);       // reverse of the forward visiting order
    const int tp = dir ? t + 1 : t - 1;                // slot that produced h_prev
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      const int idx = tid + it * GRU_THREADS;
      o[it] = Pre{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (step < p.T && idx < nb * GH) {
        const int b = idx / GH, u = idx - b * GH;
        const i64 slot = (i64)t * p.nseq + s0 + b;
        const float* g = p.gates + (slot * 2 + dir) * 400;
        o[it].r = g[u]; o[it].z = g[GH + u]; o[it].n = g[2 * GH + u]; o[it].hn = g[3 * GH + u];
        o[it].hp = (tp >= 0 && tp < p.T) ? p.y[((i64)tp * p.nseq + s0 + b) * 200 + dir * GH + u] : 0.f;
        o[it].dy = p.dy[slot * 200 + dir * GH + u];
      }
    }
  };
  float sb[ITEMS][4];                                  // running bias-gradient sums of this thread's (b, u) items
#pragma unroll
  for (int it = 0; it < ITEMS; it++) sb[it][0] = sb[it][1] = sb[it][2] = sb[it][3] = 0.f;
  fetch(0, cur);
  for (int step = 0; step < p.T; step++) {
    const int t = dir ? step : (p.T - 1 - step);
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      const int idx = tid + it * GRU_THREADS;
      if (idx >= nb * GH) continue;
      const int b = idx / GH, u = idx - b * GH;
      const i64 slot = (i64)t * p.nseq + s0 + b;
      const float r = cur[it].r, z = cur[it].z, n = cur[it].n, hn = cur[it].hn, hp = cur[it].hp;
      const float dht = dh[b][u] + part[0][b][u] + part[1][b][u] + part[2][b][u] + cur[it].dy;
      const float dn = dht * (1.0f - z);
      const float dz = dht * (hp - n);
      const float dn_pre = dn * (1.0f - n * n);
      const float dz_pre = dz * z * (1.0f - z);
      const float dr_pre = dn_pre * hn * r * (1.0f - r);
      const float dhn_ = dn_pre * r;
      float* o = p.dxg + (slot * 2 + dir) * G3;
      o[u] = dr_pre; o[GH + u] = dz_pre; o[2 * GH + u] = dn_pre;
      float* o2 = p.dgh + (slot * 2 + dir) * G3;
      o2[u] = dr_pre; o2[GH + u] = dz_pre; o2[2 * GH + u] = dhn_;
      const int gp = gru_hpos(u);
      dgh[b][gp] = dr_pre; dgh[b][GHP + gp] = dz_pre; dgh[b][2 * GHP + gp] = dhn_;
      dh[b][u] = dht * z;
      sb[it][0] += dr_pre; sb[it][1] += dz_pre; sb[it][2] += dn_pre; sb[it][3] += dhn_;
    }
    fetch(step + 1, nxt);                              // lands while the mat-vec below runs
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; b++) {
      float2 acc2[4];
#pragma unroll
      for (int i = 0; i < 4; i++) acc2[i] = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < GSL / 4; q++) {
        const float4 d4 = *reinterpret_cast<const float4*>(&dgh[b][GHP * jp + GSL * sl + 4 * q]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          acc2[i] = __ffma2_rn(w2[i][2 * q], make_float2(d4.x, d4.y), acc2[i]);
          acc2[i] = __ffma2_rn(w2[i][2 * q + 1], make_float2(d4.z, d4.w), acc2[i]);
        }
      }
      float pv[4];
#pragma unroll
      for (int i = 0; i < 4; i++) pv[i] = acc2[i].x + acc2[i].y;
      const float acc = quad_reduce_scatter(pv, sl);
      if (mv_on) part[jp][b][u_f] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; it++) cur[it] = nxt[it];
  }
  // bias gradients: one atomic per (item, gate) per CTA
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const int idx = tid + it * GRU_THREADS;
    if (idx >= nb * GH) continue;
    const int u = idx % GH;
    atomicAdd(p.db_ih[dir] + u, sb[it][0]);
    atomicAdd(p.db_ih[dir] + GH + u, sb[it][1]);
    atomicAdd(p.db_ih[dir] + 2 * GH + u, sb[it][2]);
    atomicAdd(p.db_hh[dir] + u, sb[it][0]);
    atomicAdd(p.db_hh[dir] + GH + u, sb[it][1]);
    atomicAdd(p.db_hh[dir] + 2 * GH + u, sb[it][3]);
  }
}

// dG[rowmap[slot]] += dxg[slot]  (600 floats per slot); dG pre-zeroed
__global__ void scatter_rows_kernel(const float* __restrict__ src, const int* __restrict__ rowmap, i64 nslots,
                                    float* __restrict__ dst) {
  const i64 slot = (i64)blockIdx.x * 4 + threadIdx.y;
  if (slot >= nslots) return;
  const int row = rowmap[slot];
  if (row < 0) return;
  const float* s = src + slot * 600;
  float* d = dst + (i64)row * 600;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    for (int c = threadIdx.x; c < 150; c += 32) {            // 128-bit loads and vector reductions: 150 per row instead of 600
      const float4 v = *reinterpret_cast<const float4*>(s + 4 * c);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4 * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
  } else {
    for (int c = threadIdx.x; c < 600; c += 32) atomicAdd(d + c, s[c]);
  }
}

// y = x * mask * scale (uint8 mask); in == out allowed
__global__ void mask_mul_kernel(const float* __restrict__ x, const unsigned char* __restrict__ m, float scale, i64 n,
                                float* __restrict__ y) {
  // n is a multiple of 4 (rows of 200 floats) and all three pointers are 16-byte / 4-byte aligned: 4 elements per thread
  const i64 i = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 v = *reinterpret_cast<const float4*>(x + i);
  const uint32_t k = *reinterpret_cast<const uint32_t*>(m + i);
  float4 o;
  o.x = (k & 0xFFu) ? v.x * scale : 0.f;
  o.y = (k & 0xFF00u) ? v.y * scale : 0.f;
  o.z = (k & 0xFF0000u) ? v.z * scale : 0.f;
  o.w = (k & 0xFF000000u) ? v.w * scale : 0.f;
  *reinterpret_cast<float4*>(y + i) = o;
}

__global__ void mask_mul1_kernel(const float* __restrict__ x, const unsigned char* __restrict__ m, float scale, i64 n,
                                 float* __restrict__ y) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = m[i] ? x[i] * scale : 0.f;
}
static int mask_mul(const float* x, const unsigned char* m, float scale, i64 n, float* y, cudaStream_t st) {
  const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(m) & 3) == 0;
  if (vec) mask_mul_kernel<<<(unsigned)ceil_div64(n / 4, 256), 256, 0, st>>>(x, m, scale, n, y);
  else mask_mul1_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(x, m, scale, n, y);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// Sequences per CTA (NB).  Measured with clock64 stamps (profiles/r01_gru_phase_stamps_s2.log): a step costs about
// 260 + 750 NB cycles of mat-vec (bound by shared-memory wavefronts: every warp re-reads the NB hidden vectors as
// broadcast 128-bit loads) plus 600 cycles per pass of the pointwise phase (320 items per pass), so fewer sequences
// per CTA shorten every step as long as all CTAs are co-resident.  The text encoder and the party encoder run
// concurrently on two streams; the model sets the tile for each launch through mmdfn_gru_set_tile so that both fit in
// one wave (bench shard: text NB 4 = 16 CTAs, party NB 3 = 128 CTAs).  Automatic choice (tile 0): a launch takes at
// most ~2/3 of the SMs before it widens its tiles.
static int g_gru_tile = 0;

static int gru_pick_nb(int nseq) {
  if (g_gru_tile == 2 || g_gru_tile == 3 || g_gru_tile == 4 || g_gru_tile == 8) return g_gru_tile;
  if (2 * ceil_div(nseq, 2) <= 100) return 2;
  if (2 * ceil_div(nseq, 4) <= 100) return 4;
  return (i64)ceil_div(nseq, 8) * 2 >= 148 ? 8 : 4;
}

static int launch_gru_fwd(const GruFwdArgs& a, cudaStream_t st) {
  if (a.T <= 0 || a.nseq <= 0) return 0;
  const int nb = gru_pick_nb(a.nseq);
  if (nb == 8) {
    gru_fwd_kernel<8><<<dim3(ceil_div(a.nseq, 8), 2), GRU_THREADS, 0, st>>>(a);
  } else if (nb == 4) {
    gru_fwd_kernel<4><<<dim3(ceil_div(a.nseq, 4), 2), GRU_THREADS, 0, st>>>(a);
  } else if (nb == 3) {
    gru_fwd_kernel<3><<<dim3(ceil_div(a.nseq, 3), 2), GRU_THREADS, 0, st>>>(a);
  } else {
    gru_fwd_kernel<2><<<dim3(ceil_div(a.nseq, 2), 2), GRU_THREADS, 0, st>>>(a);
  }
  MMDFN_LAUNCH_CHECK();
  return 0;
}

static int launch_gru_bwd(const GruBwdArgs& a, cudaStream_t st) {
  if (a.T <= 0 || a.nseq <= 0) return 0;
  const int nb = gru_pick_nb(a.nseq);
  if (nb == 8) {
    gru_bwd_kernel<8><<<dim3(ceil_div(a.nseq, 8), 2), GRU_THREADS, 0, st>>>(a);
  } else if (nb == 4) {
    gru_bwd_kernel<4><<<dim3(ceil_div(a.nseq, 4), 2), GRU_THREADS, 0, st>>>(a);
  } else if (nb == 3) {
    gru_bwd_kernel<3><<<dim3(ceil_div(a.nseq, 3), 2), GRU_THREADS, 0, st>>>(a);
  } else {
    gru_bwd_kernel<2><<<dim3(ceil_div(a.nseq, 2), 2), GRU_THREADS, 0, st>>>(a);
  }
  MMDFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmdfn

using namespace mmdfn;

extern "C" int mmdfn_gru_set_tile(int nb) {
  if (nb != 0 && nb != 2 && nb != 3 && nb != 4 && nb != 8) return MMDFN_EINVAL;
  g_gru_tile = nb;
  return 0;
}

// Weight pointer table order (16 entries), matching nn.GRU's state_dict names:
//   [0..3]  weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0
//   [4..7]  the same with suffix _reverse
//   [8..11] *_l1      [12..15] *_l1_reverse
extern "C" long long mmdfn_bigru2_ws_floats(int T, int nseq, long long rows) {
  const i64 slots = (i64)T * nseq;
  // xg1 (rows*600) | y1 (slots*200) | y1d (slots*200) | gates1 (slots*800) | xg2 (slots*600) | gates2 (slots*800)
  return rows * 600 + slots * (200 + 200 + 800 + 600 + 800);
}

extern "C" int mmdfn_bigru2_fwd(int T, int nseq, long long rows, const float* x, const int* rowmap,
                                const float* const* w, const unsigned char* mask, float mask_scale, float* y2,
                                float* ws, void* stream) {
  return mmdfn_bigru2_fwd_in(200, T, nseq, rows, x, rowmap, w, mask, mask_scale, y2, ws, stream);
}

extern "C" int mmdfn_bigru2_fwd_in(int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                                   const float* const* w, const unsigned char* mask, float mask_scale, float* y2,
                                   float* ws, void* stream) {
  if (!x || !w || !y2 || !ws) return MMDFN_ENULL;
  if (T < 0 || nseq < 0 || rows < 0 || in_dim <= 0) return MMDFN_EINVAL;
  if (!rowmap && rows != (i64)T * nseq) return MMDFN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const i64 slots = (i64)T * nseq;
  if (slots == 0) return 0;
  float* xg1 = ws;
  float* y1 = xg1 + rows * 600;
  float* y1d = y1 + slots * 200;
  float* gates1 = y1d + slots * 200;
  float* xg2 = gates1 + slots * 800;
  float* gates2 = xg2 + slots * 600;
  if (rows > 2000000000LL / 600 || slots > 2000000000LL / 1) return MMDFN_ERANGE;
  // layer 0 input gates for both directions
  MMDFN_TRY(gemm_nt_pair((int)rows, 300, 300, in_dim, x, in_dim, w[0], w[4], in_dim, xg1, 600, w[2], w[6], st));
  GruFwdArgs a{T, nseq, xg1, rowmap, {w[2], w[6]}, {w[1], w[5]}, {w[3], w[7]}, y1, gates1};
  MMDFN_TRY(launch_gru_fwd(a, st));
  const float* l1in = y1;
  if (mask) {
    MMDFN_TRY(mask_mul(y1, mask, mask_scale, slots * 200, y1d, st));
    l1in = y1d;
  }
  MMDFN_TRY(gemm_nt_pair((int)slots, 300, 300, 200, l1in, 200, w[8], w[12], 200, xg2, 600, w[10], w[14], st));
  GruFwdArgs b{T, nseq, xg2, nullptr, {w[10], w[14]}, {w[9], w[13]}, {w[11], w[15]}, y2, gates2};
  MMDFN_TRY(launch_gru_fwd(b, st));
  return 0;
}

extern "C" long long mmdfn_bigru2_bwd_ws_floats(int T, int nseq, long long rows) {
  const i64 slots = (i64)T * nseq;
  // dxg1 | dgh1 | dxg0 | dgh0 (slots*600 each) | dy1 (slots*200) | dG (rows*600)
  return slots * (4 * 600 + 200) + rows * 600;
}

// One layer's weight gradients.  dgate_in: (in_rows, 600) gradient w.r.t. the input gates of the rows of `xin`
// (ld 200); dgh: (T*nseq, 600); yl: that layer's output (T, nseq, 200).  Bias gradients were accumulated by the
// recurrence kernel.  beta = 1 when the caller pre-zeroed the gradient buffers (no zero-init launches).
static int gru_layer_wgrads(int T, int nseq, i64 in_rows, const float* dgate_in, const float* xin, int in_dim, const float* dgh,
                            const float* yl, float* const* dw, int base, float beta, cudaStream_t st) {
  const i64 mprev = (i64)(T - 1) * nseq;
  float* dW_ih_f = dw[base + 0];
  float* dW_ih_b = dw[base + 4];
  if (dW_ih_b == dW_ih_f + 300 * in_dim) {
    // both directions in one GEMM: [dW_ih_f; dW_ih_b] (600, in_dim) = dgate_in^T xin
    MMDFN_TRY(gemm(true, false, 600, in_dim, (int)in_rows, 1.f, dgate_in, 600, xin, in_dim, beta, dW_ih_f, in_dim, nullptr, 0, st));
  } else {
    MMDFN_TRY(gemm(true, false, 300, in_dim, (int)in_rows, 1.f, dgate_in, 600, xin, in_dim, beta, dW_ih_f, in_dim, nullptr, 0, st));
    MMDFN_TRY(gemm(true, false, 300, in_dim, (int)in_rows, 1.f, dgate_in + 300, 600, xin, in_dim, beta, dW_ih_b, in_dim, nullptr, 0, st));
  }
  for (int d = 0; d < 2; d++) {
    float* dW_hh = dw[base + 4 * d + 1];
    // h_prev of slot t is y[t-1] (forward direction) / y[t+1] (reverse direction)
    const float* Ag = d == 0 ? dgh + (i64)nseq * 600 : dgh + 300;
    const float* Bh = d == 0 ? yl : yl + (i64)nseq * 200 + 100;
    MMDFN_TRY(gemm(true, false, 300, 100, (int)mprev, 1.f, Ag, 600, Bh, 200, beta, dW_hh, 100, nullptr, 0, st));
  }
  return 0;
}

extern "C" int mmdfn_bigru2_bwd(int T, int nseq, long long rows, const float* x, const int* rowmap,
                                const float* const* w, const unsigned char* mask, float mask_scale, const float* y2,
                                const float* dy2, const float* ws_fwd, float* dx, int accumulate_dx,
                                float* const* dw, int dw_zeroed, float* ws, void* stream) {
  return mmdfn_bigru2_bwd_in(200, T, nseq, rows, x, rowmap, w, mask, mask_scale, y2, dy2, ws_fwd, dx, accumulate_dx, dw, dw_zeroed,
                             ws, stream);
}

// Backward in two parts, so that a caller may run the second one on another stream: the DATA part (both recurrences, the
// layer-1 input gradient between them, the scatter and the input gradient) is the dependency chain of the step; the
// WEIGHT-GRADIENT part (five long contractions, ~45 % of the encoder's backward kernel time) only feeds the optimizer.
// Workspace (mmdfn_bigru2_bwd_ws_floats): dxg1 | dgh1 | dxg0 | dgh0 (slots*600 each) | dy1 (slots*200) | dG (rows*600);
// the data part fills it, the weight-gradient part reads it.
struct GruBwdWs { float *dxg1, *dgh1, *dxg0, *dgh0, *dy1, *dG; };
static GruBwdWs gru_bwd_ws(float* ws, i64 slots) {
  GruBwdWs r;
  r.dxg1 = ws;
  r.dgh1 = r.dxg1 + slots * 600;
  r.dxg0 = r.dgh1 + slots * 600;
  r.dgh0 = r.dxg0 + slots * 600;
  r.dy1 = r.dgh0 + slots * 600;
  r.dG = r.dy1 + slots * 200;
  return r;
}

// parts: bit 1 = the layer-1 section (recurrence, input gradient of layer 1, its dropout mask), bit 0 = the layer-0 section
// (recurrence, scatter, input gradient); a caller that wants the layer-1 weight gradients to overlap the layer-0 recurrence
// calls data(2), wgrad(2) on another stream, data(1), wgrad(1).
extern "C" int mmdfn_bigru2_bwd_data_part(int parts, int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                                          const float* const* w, const unsigned char* mask, float mask_scale, const float* y2,
                                          const float* dy2, const float* ws_fwd, float* dx, int accumulate_dx,
                                          float* const* dw, int dw_zeroed, float* ws, void* stream) {
  if (!x || !w || !y2 || !dy2 || !ws_fwd || !dw || !ws) return MMDFN_ENULL;
  if (in_dim <= 0 || (parts & ~3) || !parts) return MMDFN_EINVAL;
  if (!rowmap && rows != (i64)T * nseq) return MMDFN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const i64 slots = (i64)T * nseq;
  if (slots == 0) return 0;
  const float* y1 = ws_fwd + rows * 600;
  const float* y1d = y1 + slots * 200;
  const float* gates1 = y1d + slots * 200;
  const float* gates2 = gates1 + slots * 800 + slots * 600;
  const GruBwdWs b = gru_bwd_ws(ws, slots);
  if (parts & 2) {
    if (!dw_zeroed) {
      for (int i = 0; i < 16; i++) {
        if ((i & 3) >= 2) MMDFN_TRY(fill_zero(dw[i], 300 * sizeof(float), st));     // bias gradients are accumulated with atomics
      }
    }
    // ---- layer 1 ----
    GruBwdArgs l1{T, nseq, dy2, y2, gates2, {w[9], w[13]}, b.dxg1, b.dgh1, {dw[10], dw[14]}, {dw[11], dw[15]}};
    MMDFN_TRY(launch_gru_bwd(l1, st));
    // d(layer-1 input) = dgates_f W_ih_f + dgates_b W_ih_b: one contraction over the 600 gate columns
    MMDFN_TRY(gemm_nn_kpair((int)slots, 200, 300, 300, b.dxg1, 600, w[8], w[12], 200, 0.f, b.dy1, 200, st));
    if (mask) {
      MMDFN_TRY(mask_mul(b.dy1, mask, mask_scale, slots * 200, b.dy1, st));
    }
  }
  if (parts & 1) {
    // ---- layer 0 ----
    GruBwdArgs l0{T, nseq, b.dy1, y1, gates1, {w[1], w[5]}, b.dxg0, b.dgh0, {dw[2], dw[6]}, {dw[3], dw[7]}};
    MMDFN_TRY(launch_gru_bwd(l0, st));
    const float* dgate_in = b.dxg0;
    if (rowmap) {
      MMDFN_TRY(fill_zero(b.dG, (size_t)rows * 600 * sizeof(float), st));
      scatter_rows_kernel<<<(unsigned)ceil_div64(slots, 4), dim3(32, 4), 0, st>>>(b.dxg0, rowmap, slots, b.dG);
      MMDFN_LAUNCH_CHECK();
      dgate_in = b.dG;
    }
    if (dx) {
      const float beta = accumulate_dx ? 1.f : 0.f;
      MMDFN_TRY(gemm_nn_kpair((int)rows, in_dim, 300, 300, dgate_in, 600, w[0], w[4], in_dim, beta, dx, in_dim, st));
    }
  }
  return 0;
}

extern "C" int mmdfn_bigru2_bwd_data(int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                                     const float* const* w, const unsigned char* mask, float mask_scale, const float* y2,
                                     const float* dy2, const float* ws_fwd, float* dx, int accumulate_dx,
                                     float* const* dw, int dw_zeroed, float* ws, void* stream) {
  return mmdfn_bigru2_bwd_data_part(3, in_dim, T, nseq, rows, x, rowmap, w, mask, mask_scale, y2, dy2, ws_fwd, dx, accumulate_dx, dw,
                                    dw_zeroed, ws, stream);
}

extern "C" int mmdfn_bigru2_bwd_wgrad_part(int parts, int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                                           const unsigned char* mask, const float* y2, const float* ws_fwd, float* const* dw,
                                           int dw_zeroed, float* ws, void* stream) {
  if (!x || !y2 || !ws_fwd || !dw || !ws) return MMDFN_ENULL;
  if (in_dim <= 0 || (parts & ~3) || !parts) return MMDFN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const i64 slots = (i64)T * nseq;
  if (slots == 0) return 0;
  const float* y1 = ws_fwd + rows * 600;
  const float* y1d = y1 + slots * 200;
  const float* l1in = mask ? y1d : y1;
  const GruBwdWs b = gru_bwd_ws(ws, slots);
  const float wbeta = dw_zeroed ? 1.f : 0.f;
  if (parts & 2) MMDFN_TRY(gru_layer_wgrads(T, nseq, slots, b.dxg1, l1in, 200, b.dgh1, y2, dw, 8, wbeta, st));
  if (parts & 1) MMDFN_TRY(gru_layer_wgrads(T, nseq, rows, rowmap ? b.dG : b.dxg0, x, in_dim, b.dgh0, y1, dw, 0, wbeta, st));
  return 0;
}

extern "C" int mmdfn_bigru2_bwd_wgrad(int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                                      const unsigned char* mask, const float* y2, const float* ws_fwd, float* const* dw,
                                      int dw_zeroed, float* ws, void* stream) {
  return mmdfn_bigru2_bwd_wgrad_part(3, in_dim, T, nseq, rows, x, rowmap, mask, y2, ws_fwd, dw, dw_zeroed, ws, stream);
}

extern "C" int mmdfn_bigru2_bwd_in(int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                                   const float* const* w, const unsigned char* mask, float mask_scale, const float* y2,
                                   const float* dy2, const float* ws_fwd, float* dx, int accumulate_dx,
                                   float* const* dw, int dw_zeroed, float* ws, void* stream) {
  MMDFN_TRY(mmdfn_bigru2_bwd_data(in_dim, T, nseq, rows, x, rowmap, w, mask, mask_scale, y2, dy2, ws_fwd, dx, accumulate_dx, dw,
                                  dw_zeroed, ws, stream));
  return mmdfn_bigru2_bwd_wgrad(in_dim, T, nseq, rows, x, rowmap, mask, y2, ws_fwd, dw, dw_zeroed, ws, stream);
}
