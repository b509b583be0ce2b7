// Second-generation tcgen05 3xTF32 GEMM (same contract as umma_gemm.cu; this kernel takes the 16-byte-aligned cases,
// the first-generation kernel keeps the rest).  Same tile (128 x 112 per CTA, two CTAs per SM), same 3-term split, same
// MMA configuration and epilogue; what changed is how operands reach the UMMA stages (round-2 phase stamps: the old
// converter spent ~2.2 k cycles per 16-wide K chunk -- cp.async staging ring, wait_group, re-read, split -- against
// ~0.7-1.0 k of tensor time; the fused graph layer's converter, built as below, needs ~0.3-0.7 k):
//   * K-contiguous operands (A of NT / NN, B of NT): 16-byte pieces go global -> REGISTERS (three chunks ahead) ->
//     hi/lo -> stage.  No staging ring, no cp.async groups, nothing to wait for but the data itself.
//   * MN-contiguous operands (B of NN, A and B of TN), whose rows are the contraction index: each 16-row chunk is copied
//     RAW by TMA bulk copies (one per row, issued by the lanes of the otherwise idle issuer warp, completing an
//     mbarrier) into a ring and read back transposed with conflict-free 32-bit loads (lanes = consecutive columns).
//   * one mbarrier arrival per converter warp, prefetches issued after the proxy fence.
#include "umma.cuh"
#include "internal.cuh"

namespace mmdfn {

constexpr int G2_BN = 112, G2_KC = 16, G2_NS = 2;
constexpr int G2_CONV = 256, G2_THREADS = G2_CONV + 32;
constexpr int G2_LBO = 128, G2_SBO = 528;
constexpr int G2_A_PART = 16 * G2_SBO;                     // 8448
constexpr int G2_B_PART = (G2_BN / 8) * G2_SBO;            // 7392
constexpr int G2_STAGE = 2 * (G2_A_PART + G2_B_PART);      // 31680
constexpr int G2_CORR = 128, G2_TMEM = 256;
constexpr int G2_DEP = 3;                                  // chunks of a K-contiguous operand held in registers
constexpr int G2_A_SLOT = G2_KC * 128 * 4;                 // raw chunk of an MN-contiguous A: 16 rows x 128 floats
constexpr int G2_B_SLOT = G2_KC * G2_BN * 4;               // raw chunk of an MN-contiguous B: 16 rows x 112 floats

template <int MODE>
struct G2Layout {
  static constexpr bool A_KMAJ = (MODE != 2), B_KMAJ = (MODE == 0);
  static constexpr int RING = MODE == 2 ? 3 : 4;           // raw ring depth (chunks)
  static constexpr int SLOT = (A_KMAJ ? 0 : G2_A_SLOT) + (B_KMAJ ? 0 : G2_B_SLOT);
  static constexpr int SMEM = G2_NS * G2_STAGE + RING * SLOT;      // 63360 / 92032 / 109440
};

struct G2Args {
  const float* A; i64 lda;
  const float* B; i64 ldb;
  float* C; i64 ldc;
  const float* bias;
  int M, N, K;
  float alpha, beta;
  int act, splits;
};

__device__ __forceinline__ void g2_split_store(const float4 v, uint8_t* hi_dst, uint8_t* lo_dst) {
  float4 h, l;
  umma::split_tf32(v.x, h.x, l.x);
  umma::split_tf32(v.y, h.y, l.y);
  umma::split_tf32(v.z, h.z, l.z);
  umma::split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi_dst) = h;
  *reinterpret_cast<float4*>(lo_dst) = l;
}

// MODE 0: NT (A[M,K], B[N,K])   1: NN (A[M,K], B[K,N])   2: TN (A[K,M], B[K,N])
template <int MODE>
__global__ void __launch_bounds__(G2_THREADS, 2) umma_gemm2_kernel(G2Args p) {
  using LY = G2Layout<MODE>;
  constexpr bool A_KMAJ = LY::A_KMAJ, B_KMAJ = LY::B_KMAJ;
  constexpr int RING = LY::RING;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_free[G2_NS];
  __shared__ __align__(8) uint64_t bar_full[G2_NS];
  __shared__ __align__(8) uint64_t bar_raw[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * G2_BN;
  const int mrows = min(128, p.M - m0), ncols = min(G2_BN, p.N - n0);

  if (warp == 8) umma::tmem_alloc(&tmem_base_s, G2_TMEM);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < G2_NS; s++) {
      umma::mbar_init(&bar_free[s], 1);
      umma::mbar_init(&bar_full[s], G2_CONV / 32);
    }
    for (int s = 0; s < 4; s++) umma::mbar_init(&bar_raw[s], 1);
    umma::fence_barrier_init();
  }
  int kb = 0, ke = p.K;
  if (p.splits > 1) {
    const int chunk = ((p.K + p.splits - 1) / p.splits + G2_KC - 1) / G2_KC * G2_KC;
    kb = blockIdx.z * chunk;
    ke = min(p.K, kb + chunk);
  }
  const int nchunks = ke > kb ? (ke - kb + G2_KC - 1) / G2_KC : 0;
  uint8_t* stages = smem;
  uint8_t* ring = smem + G2_NS * G2_STAGE;
  constexpr uint32_t IDESC = umma::idesc_tf32(128, G2_BN);

  // raw chunk c of the MN-contiguous operand(s) -> ring slot c % RING: one bulk copy per k-row, lanes 0..15 take A's rows
  // (TN), lanes 16..31 (TN) or 0..15 (NN) B's rows
  auto issue_raw = [&](int c, int ln) {
    const int k0 = kb + c * G2_KC;
    const int rows = min(G2_KC, ke - k0);
    const uint32_t bar = umma::smem_u32(&bar_raw[c % RING]);
    uint8_t* slot = ring + (c % RING) * LY::SLOT;
    const uint32_t a_bytes = A_KMAJ ? 0u : (uint32_t)mrows * 4u, b_bytes = B_KMAJ ? 0u : (uint32_t)ncols * 4u;
    if (ln == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)rows * (a_bytes + b_bytes)) : "memory");
    if (!A_KMAJ && ln < rows)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(umma::smem_u32(slot + ln * 512)), "l"(p.A + (i64)(k0 + ln) * p.lda + m0), "r"(a_bytes), "r"(bar) : "memory");
    const int lb = A_KMAJ ? ln : ln - 16;
    if (!B_KMAJ && lb >= 0 && lb < rows)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(umma::smem_u32(slot + (A_KMAJ ? 0 : G2_A_SLOT) + lb * (G2_BN * 4))), "l"(p.B + (i64)(k0 + lb) * p.ldb + n0),
                     "r"(b_bytes), "r"(bar) : "memory");
  };

  // ---- per-thread pieces ----
  // K-contiguous operand, R rows: piece pi = tid + 256 i (< 4 R): quarter-warp qw = pi / 8 -> row group qw / 4, k-quad qw % 4
  // MN-contiguous operand, R rows: piece pi (< 4 R) -> row pi % R, k-quad pi / R (lanes = consecutive rows: conflict-free)
  constexpr int NPA = 2, NPB = 2;                            // 512 A pieces, 448 B pieces per chunk
  int a_row[NPA], a_kq[NPA], a_off[NPA];
  int b_row[NPB], b_kq[NPB], b_off[NPB];
#pragma unroll
  for (int i = 0; i < NPA; i++) {
    const int pi = tid + G2_CONV * i;
    if (A_KMAJ) { a_kq[i] = (pi >> 3) & 3; a_row[i] = (pi >> 5) * 8 + (pi & 7); }
    else { a_kq[i] = pi >> 7; a_row[i] = pi & 127; }
    a_off[i] = (a_row[i] >> 3) * G2_SBO + a_kq[i] * G2_LBO + (a_row[i] & 7) * 16;
  }
#pragma unroll
  for (int i = 0; i < NPB; i++) {
    const int pi = tid + G2_CONV * i;
    if (B_KMAJ) { b_kq[i] = (pi >> 3) & 3; b_row[i] = (pi >> 5) * 8 + (pi & 7); }
    else { b_kq[i] = pi / G2_BN; b_row[i] = pi - G2_BN * b_kq[i]; }
    b_off[i] = (pi < 4 * G2_BN) ? (b_row[i] >> 3) * G2_SBO + b_kq[i] * G2_LBO + (b_row[i] & 7) * 16 : -1;
  }
  // K-contiguous loads: element (row, k) at g[row * ld + k]; rows beyond the matrix and k beyond ke read as zero
  auto ldk = [&](const float* g, i64 ld, int row, int rows_valid, int k) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows_valid && k < ke) {
      const float* q = g + (i64)row * ld + k;
      if (k + 3 < ke) {
        v = __ldg(reinterpret_cast<const float4*>(q));
      } else {
        v.x = q[0];
        if (k + 1 < ke) v.y = q[1];
        if (k + 2 < ke) v.z = q[2];
      }
    }
    return v;
  };
  float4 ra[A_KMAJ ? G2_DEP : 1][NPA], rb[B_KMAJ ? G2_DEP : 1][NPB];
  const float* Ag = p.A + (A_KMAJ ? (i64)m0 * p.lda : 0);
  const float* Bg = p.B + (B_KMAJ ? (i64)n0 * p.ldb : 0);
  auto prefetch = [&](int c, int slot) {
    if (c >= nchunks) return;
    const int k0 = kb + c * G2_KC;
    if (A_KMAJ) {
#pragma unroll
      for (int i = 0; i < NPA; i++) ra[A_KMAJ ? slot : 0][i] = ldk(Ag, p.lda, a_row[i], mrows, k0 + 4 * a_kq[i]);
    }
    if (B_KMAJ) {
#pragma unroll
      for (int i = 0; i < NPB; i++)
        if (b_off[i] >= 0) rb[B_KMAJ ? slot : 0][i] = ldk(Bg, p.ldb, b_row[i], ncols, k0 + 4 * b_kq[i]);
    }
  };
  if (warp < 8) {
#pragma unroll
    for (int c = 0; c < G2_DEP; c++) prefetch(c, c);
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == 8) {
    // ===== MMA issuer (lane 0) + raw-chunk producer (all lanes) =====
    if (!A_KMAJ || !B_KMAJ)
      for (int c = 0; c < RING && c < nchunks; c++) issue_raw(c, lane);
    const uint64_t d0 = umma::smem_desc(umma::smem_u32(stages), G2_LBO, G2_SBO);
    const uint32_t dhi = (uint32_t)(d0 >> 32), dlo = (uint32_t)d0;
    for (int c = 0; c < nchunks; c++) {
      const int s = c % G2_NS;
      umma::mbar_wait(&bar_full[s], (uint32_t)((c / G2_NS) & 1));       // the whole warp: the MMA issue below is convergent
      umma::tc_fence_after_sync();
      if ((!A_KMAJ || !B_KMAJ) && c + RING < nchunks) issue_raw(c + RING, lane);      // slot c % RING was fully read
      const uint32_t o = dlo + (uint32_t)s * (G2_STAGE >> 4);
      const int kleft = ke - (kb + c * G2_KC);
      const int ksteps = kleft >= G2_KC ? G2_KC / 8 : (kleft + 7) / 8;
      for (int j = 0; j < ksteps; j++) {
        const uint32_t oj = o + (uint32_t)j * ((2 * G2_LBO) >> 4);
        umma::kstep3_elect(tmem, tmem + G2_CORR, dhi, oj, oj + (G2_A_PART >> 4), oj + ((2 * G2_A_PART) >> 4),
                           oj + ((2 * G2_A_PART + G2_B_PART) >> 4), IDESC, (c > 0 || j > 0) ? 1u : 0u);
      }
      umma::mma_commit_elect(&bar_free[s]);
    }
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tmem_dealloc(tmem, G2_TMEM);
    return;
  }

  // ===== converters (warps 0-7) =====
  for (int c0 = 0; c0 < nchunks; c0 += G2_DEP) {
#pragma unroll
    for (int u = 0; u < G2_DEP; u++) {
      const int c = c0 + u;
      if (c < nchunks) {
        const int s = c % G2_NS;
        if (!A_KMAJ || !B_KMAJ) umma::mbar_wait(&bar_raw[c % RING], (uint32_t)((c / RING) & 1));
        if (c >= G2_NS) umma::mbar_wait(&bar_free[s], (uint32_t)(((c / G2_NS) - 1) & 1));
        uint8_t* st = stages + s * G2_STAGE;
        const uint8_t* slot = ring + (c % RING) * LY::SLOT;
        const int kleft = ke - (kb + c * G2_KC);             // valid k rows in this chunk (>= 1)
        // ---- A
#pragma unroll
        for (int i = 0; i < NPA; i++) {
          float4 v;
          if (A_KMAJ) {
            v = ra[A_KMAJ ? u : 0][i];
          } else {
            const float* q = reinterpret_cast<const float*>(slot) + (4 * a_kq[i]) * 128 + a_row[i];
            const int jl = kleft - 4 * a_kq[i];
            const bool ok = a_row[i] < mrows;
            v = make_float4(ok && jl > 0 ? q[0] : 0.f, ok && jl > 1 ? q[128] : 0.f, ok && jl > 2 ? q[256] : 0.f,
                            ok && jl > 3 ? q[384] : 0.f);
          }
          g2_split_store(v, st + a_off[i], st + G2_A_PART + a_off[i]);
        }
        // ---- B
#pragma unroll
        for (int i = 0; i < NPB; i++) {
          if (b_off[i] < 0) continue;
          float4 v;
          if (B_KMAJ) {
            v = rb[B_KMAJ ? u : 0][i];
          } else {
            const float* q = reinterpret_cast<const float*>(slot + (A_KMAJ ? 0 : G2_A_SLOT)) + (4 * b_kq[i]) * G2_BN + b_row[i];
            const int jl = kleft - 4 * b_kq[i];
            const bool ok = b_row[i] < ncols;
            v = make_float4(ok && jl > 0 ? q[0] : 0.f, ok && jl > 1 ? q[G2_BN] : 0.f, ok && jl > 2 ? q[2 * G2_BN] : 0.f,
                            ok && jl > 3 ? q[3 * G2_BN] : 0.f);
          }
          g2_split_store(v, st + 2 * G2_A_PART + b_off[i], st + 2 * G2_A_PART + G2_B_PART + b_off[i]);
        }
        umma::warp_arrive_full(&bar_full[s]);
        prefetch(c + G2_DEP, u);                             // after the proxy fence: it waits for outstanding loads
      }
    }
  }
  if (nchunks > 0) {
    const int last = nchunks - 1;
    umma::mbar_wait(&bar_free[last % G2_NS], (uint32_t)((last / G2_NS) & 1));
  }
  umma::tc_fence_after_sync();

  // ---- epilogue: thread = output row (TMEM lane 32*(warp%4)+lane); the two warp groups split the columns ----
  const int row = m0 + (warp & 3) * 32 + lane;
  const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const bool c_vec = (p.splits <= 1) && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
  constexpr int HALF = 64;
  const int cb_begin = (warp < 4) ? 0 : HALF, cb_end = (warp < 4) ? HALF : G2_BN;
  const bool c_old = (p.beta != 0.f) && (p.splits <= 1) && (row < p.M);
  float4 cold[4];
  auto fetch_c = [&](int cb) {
    if (!c_old || cb >= cb_end || n0 + cb >= p.N) return;
    const float* crow = p.C + (i64)row * p.ldc + n0 + cb;
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      const int n = n0 + cb + 4 * q4;
      if (c_vec && n + 3 < p.N) {
        t = *reinterpret_cast<const float4*>(crow + 4 * q4);
      } else {
        if (n < p.N) t.x = crow[4 * q4];
        if (n + 1 < p.N) t.y = crow[4 * q4 + 1];
        if (n + 2 < p.N) t.z = crow[4 * q4 + 2];
        if (n + 3 < p.N) t.w = crow[4 * q4 + 3];
      }
      cold[q4] = t;
    }
  };
  fetch_c(cb_begin);
#pragma unroll 1
  for (int cb = cb_begin; cb < cb_end; cb += 16) {
    if (n0 + cb >= p.N) break;                      // warp-uniform
    float v[16];
    float cprev[16];
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {
      cprev[4 * q4] = cold[q4].x; cprev[4 * q4 + 1] = cold[q4].y; cprev[4 * q4 + 2] = cold[q4].z; cprev[4 * q4 + 3] = cold[q4].w;
    }
    fetch_c(cb + 16);
    if (nchunks > 0) {
      float w[16];
      umma::tmem_ld16x2(taddr + cb, taddr + G2_CORR + cb, v, w);
#pragma unroll
      for (int q = 0; q < 16; q++) v[q] += w[q];
    } else {
#pragma unroll
      for (int q = 0; q < 16; q++) v[q] = 0.f;
    }
    if (row >= p.M) continue;
    float* crow = p.C + (i64)row * p.ldc + n0 + cb;
    if (p.splits > 1) {
      const bool r_vec = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((n0 & 3) == 0);
#pragma unroll
      for (int q4 = 0; q4 < 16; q4 += 4) {
        if (r_vec && n0 + cb + q4 + 3 < p.N) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + q4), "f"(p.alpha * v[q4]),
                       "f"(p.alpha * v[q4 + 1]), "f"(p.alpha * v[q4 + 2]), "f"(p.alpha * v[q4 + 3]) : "memory");
        } else {
#pragma unroll
          for (int q = 0; q < 4; q++)
            if (n0 + cb + q4 + q < p.N) atomicAdd(crow + q4 + q, p.alpha * v[q4 + q]);
        }
      }
      continue;
    }
#pragma unroll
    for (int q4 = 0; q4 < 16; q4 += 4) {
      float o[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int n = n0 + cb + q4 + q;
        float t = p.alpha * v[q4 + q];
        if (n < p.N) {
          if (p.beta != 0.f) t = fmaf(p.beta, cprev[q4 + q], t);
          if (p.bias) t += p.bias[n];
        }
        if (p.act == 1) t = fmaxf(t, 0.f);
        o[q] = t;
      }
      if (c_vec && n0 + cb + q4 + 3 < p.N) {
        *reinterpret_cast<float4*>(crow + q4) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (n0 + cb + q4 + q < p.N) crow[q4 + q] = o[q];
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
}

template <int MODE>
static int launch_g2(const G2Args& p, cudaStream_t st) {
  using LY = G2Layout<MODE>;
  static bool configured = false;
  if (!configured) {
    MMDFN_CUDA(cudaFuncSetAttribute(umma_gemm2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, LY::SMEM));
    configured = true;
  }
  dim3 grid(ceil_div(p.M, 128), ceil_div(p.N, G2_BN), p.splits > 1 ? p.splits : 1);
  umma_gemm2_kernel<MODE><<<grid, G2_THREADS, LY::SMEM, st>>>(p);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// true when this kernel can take the problem: every operand 16-byte aligned with leading dimensions that are
// multiples of 4 floats, and -- for MN-contiguous operands, whose rows travel as bulk copies -- M / N multiples of 4
bool umma_gemm2_eligible(bool ta, bool tb, int M, int N, int K, const float* A, i64 lda, const float* B, i64 ldb) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al(A) || !al(B) || (lda & 3) || (ldb & 3) || K <= 0) return false;
  if (ta && (M & 3)) return false;                 // TN: A rows are M-contiguous
  if (!tb && (N & 3)) return false;                // NN / TN: B rows are N-contiguous
  return true;
}

int umma_gemm2(bool ta, bool tb, int M, int N, int K, float alpha, const float* A, i64 lda, const float* B, i64 ldb,
               float beta, float* C, i64 ldc, const float* bias, int act, int splits, cudaStream_t st) {
  G2Args p{A, lda, B, ldb, C, ldc, bias, M, N, K, alpha, beta, act, splits};
  if (!ta && tb) return launch_g2<0>(p, st);
  if (!ta && !tb) return launch_g2<1>(p, st);
  return launch_g2<2>(p, st);
}

}  // namespace mmdfn
