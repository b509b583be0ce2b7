// Third-generation tcgen05 3xTF32 GEMM (same contract as umma_gemm.cu / umma_gemm2.cu: C = act(alpha op(A) op(B) + beta C +
// bias), fp32-level accuracy through the 3-term tf32 split with separate main / correction accumulators).
//
// What limited the first two generations (profiles/r02_umma_gemm_gen2.log: 30-60 TFLOP/s on the path's shapes): BOTH
// operands were split into hi/lo by CUDA cores and written to shared memory, from where the tensor core read them back
// (per 16-wide K chunk of a 128 x 112 tile: 32 KB written, 48 KB read; two shallow stages per CTA), and the epilogue wrote
// one 16-byte piece per output row per instruction.  This generation follows the fused graph layer (gcn_layer2.cu):
//   * the A operand never touches shared memory: thread = tile row reads its own 4 k-values of a chunk straight from
//     global memory (K-contiguous A: one 16-byte load; M-contiguous A of the TN form: four coalesced 32-bit loads),
//     splits them in registers and tcgen05.st's hi / lo into a ring of TENSOR-MEMORY chunks; the MMAs take A from there
//     (tcgen05.mma [d], [a_tmem], b_desc).  Shared memory carries only B.
//   * one wide CTA per SM: 128 rows x BN columns with BN = N split evenly into tiles of <= 192 columns (N = 300 -> 2 x 160,
//     200 -> 2 x 112, 400 -> 3 x 144, 600 -> 4 x 160, 100 -> 112), so the A conversion and its global traffic are
//     amortised over up to 192 columns instead of 112.
//   * 16 converter warps in 4 GROUPS (one warp per TMEM lane quarter each); group g owns the chunks c = g (mod 4) entirely
//     (its 128 threads convert the chunk's 128 x 16 A block and all B pieces).  A chunk's chain -- split, tcgen05.st,
//     wait::st, proxy fence, mbarrier arrive -- is ~1-2 k cycles of latency for one warp; with four groups in flight on
//     four different chunks it is paid once per FOUR chunks (the first version, in which every warp touched every
//     chunk, measured ~1.9 k cycles per chunk: no faster than generation 2).  A 17th warp issues the MMAs convergently.
//   * a K-contiguous A (NT / NN forms) reaches its row threads through shared memory: a producer warp issues ONE tensor-map
//     TMA copy (cp.async.bulk.tensor.2d, 128 rows x 32 floats, 128-byte swizzle; no L1 tag traffic, no registers, zero
//     fill outside the matrix) per 32-wide K panel into a 4-deep raw ring, and the row thread reads its own 64 bytes with
//     conflict-free 128-bit loads.  (One bulk copy per ROW was tried first: 128-byte copies are far below the copy
//     engine's efficient size, the kernel ran 2-3x slower.)  (Reading the row straight from global memory
//     -- 32 lanes, 32 different cache lines per instruction -- kept L1 at 54 % busy and the warps on the long scoreboard:
//     profiles/r02_ncu_umma_gemm3_v1.csv.)  The M-contiguous A of the TN form is read with coalesced 32-bit loads.
//   * stages = (TMEM A chunk, smem B chunk) pairs, 4..6 deep (what 512 TMEM columns leave after the two accumulators);
//     a group prefetches its NEXT chunk (four chunks ahead) into registers right after handing a chunk over.
//   * epilogue through shared memory: every warp drains 32 x 32 panels TMEM -> registers -> a padded scratch tile and
//     writes them back with lanes along the row: 128-byte coalesced stores (or vector reductions for split-K) instead of
//     32 scattered 16-byte pieces.
#include <cuda.h>
#include "umma.cuh"
#include "internal.cuh"

namespace mmdfn {

constexpr int G3_KC = 16;
constexpr int G3_CONVW = 16, G3_CONV = 32 * G3_CONVW, G3_THREADS = G3_CONV + 64;     // + MMA warp + raw-A producer warp
constexpr int G3_LBO = 128, G3_SBO = 528;
// stages = converter groups: chunk c and chunk c + 4 are converted by the same group, so a stage (shared-memory B chunk +
// tensor-memory A chunk) is only ever rewritten by the warps that wrote it before, after the mbarrier that the MMAs'
// tcgen05.commit completes -- program order plus that barrier, which also keeps compute-sanitizer's racecheck (it does
// not follow the commit -> mbarrier edge across warps) free of false reports.  Deeper rings (6) measured no faster: a
// group's iteration (~4 k cycles) is far longer than its chunk's MMAs (~0.5 k).
constexpr int G3_BN_MAX = 192, G3_NS_MAX = 4;
constexpr int G3_GROUPS = 4;                               // converter groups (4 warps = 128 threads each)
// raw panels of a K-contiguous A: 2 chunks (32 floats = 128 B) of each of the 128 tile rows = one 16 KB box of a 2-D
// tensor map (ONE cp.async.bulk.tensor per panel; rows / columns outside the matrix arrive as zeros).  SWIZZLE_128B:
// the 16-byte piece j of row r sits at piece j ^ (r & 7), so the pieces that the 32 lanes (= 32 rows) of a warp read in
// one instruction fall into distinct bank groups.  Panels are 1024-byte aligned.
constexpr int G3_PROW = 128, G3_PANEL = 128 * G3_PROW, G3_NP = 4;
constexpr int G3_SCRATCH = 32 * 33 * 4;                    // per-warp epilogue panel (padded: conflict-free both ways)

struct G3Args {
  const float* A; i64 lda;
  const float* B; i64 ldb;
  float* C; i64 ldc;
  const float* bias;
  int M, N, K;
  float alpha, beta;
  int act, splits;
  int bn, ns;                                              // column tile (multiple of 16, <= 192), stages
  // operand pairs (one launch instead of two): NT form with nsplit > 0: output columns [0, nsplit) come from B (bias),
  // columns [nsplit, N) from B2 (bias2); NN form with ksplit > 0: contraction rows [0, ksplit) of B, the rest from B2
  const float* B2;
  const float* bias2;
  float* C2;                                               // column pair: where the second half's columns go (its column 0), with beta2
  float beta2;
  int nsplit, ksplit;
  int dbg;                                                 // profiling aid: 1 = skip the MMAs (conversion pace only)
  long long* stamps;                                       // profiling aid: clock64 stamps of CTA 0's converter warp 0 (8 per chunk it owns)
};

__device__ __forceinline__ void g3_tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
                 "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
                 "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
                 "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void g3_tmem_ld8x2(uint32_t ta, uint32_t tb, float (&v)[8]) {
  uint32_t r[8], q[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "r"(tb) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]) + __uint_as_float(q[i]);
}

// x = hi + lo with hi = x rounded to tf32 (nearest, ties away: add half an ulp of the 13 dropped bits to the magnitude,
// clear them) and lo = x - hi exact.  Same values as cvt.rna.tf32.f32 for finite inputs, in two integer operations --
// the conversion instruction expands to five with its NaN / infinity handling, and this kernel is issue-bound.
__device__ __forceinline__ void g3_split(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}
__device__ __forceinline__ void g3_split4(const float4 v, float4& h, float4& l) {
  g3_split(v.x, h.x, l.x);
  g3_split(v.y, h.y, l.y);
  g3_split(v.z, h.z, l.z);
  g3_split(v.w, h.w, l.w);
}

// MODE 0: NT (A[M,K], B[N,K])   1: NN (A[M,K], B[K,N])   2: TN (A[K,M], B[K,N])
template <int MODE>
__global__ void __launch_bounds__(G3_THREADS, 1) umma_gemm3_kernel(const __grid_constant__ CUtensorMap tmA, G3Args p) {
  constexpr bool A_KMAJ = (MODE != 2), B_KMAJ = (MODE == 0);
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[G3_NS_MAX];
  __shared__ __align__(8) uint64_t bar_free[G3_NS_MAX];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ __align__(8) uint64_t raw_full[G3_NP];
  __shared__ __align__(8) uint64_t raw_free[G3_NP];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.bn, ns = p.ns;
  // column tile: with an NT operand pair the grid's y axis walks the tiles of the first half, then of the second
  int ty = blockIdx.y, ncap = p.N;
  const float* Bsrc = p.B;
  const float* biasp = p.bias;
  float* Cdst = p.C;
  float beta = p.beta;
  if (p.nsplit > 0) {
    const int t1 = (p.nsplit + BN - 1) / BN;
    if (ty >= t1) { ty -= t1; Bsrc = p.B2; biasp = p.bias2; ncap = p.N - p.nsplit; Cdst = p.C2; beta = p.beta2; }
    else ncap = p.nsplit;
  }
  const int m0 = blockIdx.x * 128, n0 = ty * BN;             // n0: column within this half's operand
  const int mrows = min(128, p.M - m0), ncols = min(BN, ncap - n0);
  const int bpart = (BN >> 3) * G3_SBO, stage_bytes = 2 * bpart;
  // K-contiguous A: G3_NP raw panels (1024-byte aligned: swizzle atom), then the B stages
  uint8_t* const raw = smem + ((1024u - (umma::smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* const stages = A_KMAJ ? raw + G3_NP * G3_PANEL : smem;
  const uint32_t tm_corr = (uint32_t)BN, tm_a = (uint32_t)(2 * BN);

  if (warp == G3_CONVW) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < G3_NS_MAX; s++) {
      umma::mbar_init(&bar_full[s], G3_CONVW / G3_GROUPS);
      umma::mbar_init(&bar_free[s], 1);
    }
    umma::mbar_init(&bar_done, 1);
    for (int s = 0; s < G3_NP; s++) {
      umma::mbar_init(&raw_full[s], 1);
      umma::mbar_init(&raw_free[s], 2 * G3_CONVW / G3_GROUPS);     // the two groups that convert the panel's two chunks
    }
    umma::fence_barrier_init();
  }
  int kb = 0, ke = p.K;
  if (p.splits > 1) {
    const int chunk = ((p.K + p.splits - 1) / p.splits + G3_KC - 1) / G3_KC * G3_KC;
    kb = blockIdx.z * chunk;
    ke = min(p.K, kb + chunk);
  }
  const int nchunks = ke > kb ? (ke - kb + G3_KC - 1) / G3_KC : 0;
  // K order: every CTA walks its 32-wide K panels in a ROTATED order (start panel = a function of the tile index).  All
  // CTAs of a column tile read the same B rows; in lockstep they requested the same few cache lines at the same time and
  // the loads queued at those L2 lines (in-kernel stamps: ~2.7 k cycles to get five 16-byte loads per thread issued).
  // Chunk slot c -> chunk kchunk(c); an odd chunk count gets one phantom slot (its MMAs are skipped).
  const int npan = (nchunks + 1) >> 1, nslots = 2 * npan;
  const int prot = npan > 0 ? (int)((blockIdx.x + 7u * blockIdx.y + 3u * blockIdx.z) % (unsigned)npan) : 0;
  auto kchunk = [&](int c) {
    int ap = (c >> 1) + prot;
    if (ap >= npan) ap -= npan;
    return 2 * ap + (c & 1);
  };

  // ---- per-thread operand pieces ----
  const int q = warp & 3, g = warp >> 2;                     // converter warp: TMEM lane quarter, group
  const int arow = 32 * q + lane;                            // tile row owned by this thread (= its index within the group)
  // B: piece pi = arow + 128 i (< 4 BN): (row n, k-quad kq).  K-contiguous B: a quarter-warp covers 8 rows of one k-quad
  // (16-byte loads); N-contiguous B: lanes = consecutive rows (coalesced 32-bit loads, four k rows per piece)
  constexpr int NPB = (4 * G3_BN_MAX) / 128;                 // 6
  int b_row[NPB], b_kq[NPB], b_off[NPB];
#pragma unroll
  for (int i = 0; i < NPB; i++) {
    const int pi = arow + 128 * i;
    if (B_KMAJ) { b_kq[i] = (pi >> 3) & 3; b_row[i] = (pi >> 5) * 8 + (pi & 7); }
    else { b_kq[i] = pi / BN; b_row[i] = pi - BN * b_kq[i]; }
    b_off[i] = (warp < G3_CONVW && pi < 4 * BN) ? (b_row[i] >> 3) * G3_SBO + b_kq[i] * G3_LBO + (b_row[i] & 7) * 16 : -1;
  }
  // element (row, k..k+3) of a K-contiguous operand; rows beyond the matrix and k beyond ke read as zero
  auto ldk = [&](const float* gp, i64 ld, int row, int rows_valid, int k) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows_valid && k < ke) {
      const float* s = gp + (i64)row * ld + k;
      if (k + 3 < ke) {
        v = __ldg(reinterpret_cast<const float4*>(s));
      } else {
        v.x = s[0];
        if (k + 1 < ke) v.y = s[1];
        if (k + 2 < ke) v.z = s[2];
      }
    }
    return v;
  };
  // element (k..k+3, col) of an MN-contiguous operand (rows = contraction index)
  auto ldm = [&](const float* gp, i64 ld, int col, int cols_valid, int k) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < cols_valid) {
      const float* s = gp + (i64)k * ld + col;
      if (k < ke) v.x = __ldg(s);
      if (k + 1 < ke) v.y = __ldg(s + ld);
      if (k + 2 < ke) v.z = __ldg(s + 2 * ld);
      if (k + 3 < ke) v.w = __ldg(s + 3 * ld);
    }
    return v;
  };
  const float* Ag = A_KMAJ ? p.A + (i64)m0 * p.lda : p.A + m0;
  const float* Bg = B_KMAJ ? Bsrc + (i64)n0 * p.ldb : Bsrc + n0;
  // row k of an N-contiguous B (an operand pair along K switches to B2 at ksplit, a multiple of 4)
  auto brow = [&](int k) -> const float* {
    return (p.ksplit > 0 && k >= p.ksplit) ? p.B2 + n0 + (i64)(k - p.ksplit) * p.ldb : Bg + (i64)k * p.ldb;
  };
  // whole-chunk fast path: base pointers of this thread's pieces at k = kb (advanced by 16 k per chunk), validity flags
  const float* aptr = Ag + (i64)kb * p.lda + arow;                       // M-contiguous A only
  const bool aok = arow < mrows;
  const float* bptr[NPB];
  bool bok[NPB];
#pragma unroll
  for (int i = 0; i < NPB; i++) {
    bok[i] = b_off[i] >= 0 && b_row[i] < ncols;
    bptr[i] = B_KMAJ ? Bg + (i64)b_row[i] * p.ldb + kb + 4 * b_kq[i] : nullptr;
  }
  const i64 astep = (i64)G3_KC * p.lda;
  float4 ra[A_KMAJ ? 1 : 4], rb[NPB];
  auto prefetch = [&](int c) {
    if (c >= nslots) return;
    const int cc = kchunk(c);
    const int k0 = kb + cc * G3_KC;
    if (k0 + G3_KC <= ke) {
      // every k of the chunk is inside the contraction range: no per-element checks
      if (!A_KMAJ) {
        const float* a = aptr + (i64)cc * astep;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (aok) {
            v.x = __ldg(a + (i64)(4 * j) * p.lda);
            v.y = __ldg(a + (i64)(4 * j + 1) * p.lda);
            v.z = __ldg(a + (i64)(4 * j + 2) * p.lda);
            v.w = __ldg(a + (i64)(4 * j + 3) * p.lda);
          }
          ra[A_KMAJ ? 0 : j] = v;
        }
      }
#pragma unroll
      for (int i = 0; i < NPB; i++) {
        if (b_off[i] < 0) continue;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bok[i] && p.dbg != 7) {
          if (B_KMAJ) {
            v = __ldg(reinterpret_cast<const float4*>(bptr[i] + cc * G3_KC));
          } else {
            const float* b = brow(k0 + 4 * b_kq[i]) + b_row[i];
            v.x = __ldg(b);
            v.y = __ldg(b + p.ldb);
            v.z = __ldg(b + 2 * p.ldb);
            v.w = __ldg(b + 3 * p.ldb);
          }
        }
        rb[i] = v;
      }
      return;
    }
    if (!A_KMAJ) {
#pragma unroll
      for (int j = 0; j < 4; j++) ra[A_KMAJ ? 0 : j] = ldm(Ag, p.lda, arow, mrows, k0 + 4 * j);
    }
#pragma unroll
    for (int i = 0; i < NPB; i++)
      if (b_off[i] >= 0)
        rb[i] = B_KMAJ ? ldk(Bg, p.ldb, b_row[i], ncols, k0 + 4 * b_kq[i])
                       : ldm(brow(k0 + 4 * b_kq[i]) - (i64)(k0 + 4 * b_kq[i]) * p.ldb, p.ldb, b_row[i], ncols, k0 + 4 * b_kq[i]);
  };
  if (warp < G3_CONVW) prefetch(g);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == G3_CONVW) {
    // ===== MMA issuer: the whole warp runs the loop on uniform values, one elected lane issues =====
    const uint32_t idesc = umma::idesc_tf32(128, BN);
    const uint64_t d0 = umma::smem_desc(umma::smem_u32(stages), G3_LBO, G3_SBO);
    const uint32_t d0_hi = (uint32_t)(d0 >> 32), d0_lo = (uint32_t)d0;
    int s = 0;
    uint32_t par = 0;
    uint32_t started = 0;
    for (int c = 0; c < nslots; c++) {
      umma::mbar_wait(&bar_full[s], par);
      umma::tc_fence_after_sync();
      const int kleft = ke - (kb + kchunk(c) * G3_KC);
      const int ksteps = kleft >= G3_KC ? 2 : kleft > 0 ? (kleft + 7) >> 3 : 0;       // 0: the phantom slot of an odd chunk count
      const uint32_t acol = tm_a + 32u * (uint32_t)s;
      for (int j = 0; j < ksteps; j++) {
        const uint32_t dlo = d0_lo + (uint32_t)((s * stage_bytes + j * 2 * G3_LBO) >> 4);
        const uint64_t b_hi = ((uint64_t)d0_hi << 32) | dlo;
        const uint64_t b_lo = ((uint64_t)d0_hi << 32) | (dlo + (uint32_t)(bpart >> 4));
        const uint32_t acc = started;
        started = 1u;
        if (p.dbg == 1) continue;
        umma::mma_tf32_ta_elect(tmem, tmem + acol + 8u * j, b_hi, idesc, acc);
        umma::mma_tf32_ta_elect(tmem + tm_corr, tmem + acol + 16u + 8u * j, b_hi, idesc, acc);
        umma::mma_tf32_ta_elect(tmem + tm_corr, tmem + acol + 8u * j, b_lo, idesc, 1u);
      }
      umma::mma_commit_elect(&bar_free[s]);
      if (++s == ns) { s = 0; par ^= 1u; }
    }
    umma::mma_commit_elect(&bar_done);
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tmem_dealloc(tmem, 512);
    return;
  }

  if (warp == G3_CONVW + 1) {
    // ===== raw-A producer (K-contiguous A only): one bulk copy per tile row per panel =====
    if (A_KMAJ) {
      for (int pp = 0; pp < npan; pp++) {
        const int slot = pp % G3_NP, use = pp / G3_NP;
        if (use > 0) umma::mbar_wait(&raw_free[slot], (uint32_t)((use - 1) & 1));
        const int kp = kb + G3_KC * kchunk(2 * pp);
        const uint32_t bar = umma::smem_u32(&raw_full[slot]);
        if (lane == 0) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)G3_PANEL) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(umma::smem_u32(raw + slot * G3_PANEL)), "l"(&tmA), "r"(kp), "r"(m0), "r"(bar) : "memory");
        }
        __syncwarp();
      }
    }
    umma::tc_fence_before_sync();
    __syncthreads();
    return;
  }

  // ===== converters (warps 0-15) =====
  const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
  const bool stamp_on = p.stamps != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
#define G3_STAMP(slot) do { if (stamp_on && c < 32) p.stamps[8 * (c >> 2) + (slot)] = clock64(); } while (0)
  for (int c = g; c < nslots; c += G3_GROUPS) {
    const int use = c / ns, s = c - use * ns;
    G3_STAMP(0);
    if (use > 0) {
      umma::mbar_wait(&bar_free[s], (uint32_t)((use - 1) & 1));              // the MMAs of chunk c - ns have retired
      umma::tc_fence_after_sync();
    }
    // ---- A: this row's 16 k-values -> hi / lo -> TMEM chunk s (hi columns 0..15, lo columns 16..31)
    {
      float4 av[4];
      if (A_KMAJ) {
        const int pp = c >> 1, slot = pp % G3_NP;
        G3_STAMP(1);
        umma::mbar_wait(&raw_full[slot], (uint32_t)((pp / G3_NP) & 1));
        G3_STAMP(2);
        // rows / columns outside the matrix were zero-filled by the copy engine and a K split ends on a chunk boundary:
        // nothing to mask
        const uint8_t* src = raw + slot * G3_PANEL + arow * G3_PROW;
        const int pj = (c & 1) * 4, sw = (p.dbg == 6) ? 0 : (arow & 7);   // first 16-byte piece of this chunk, swizzle of this row
#pragma unroll
        for (int j = 0; j < 4; j++) av[j] = *reinterpret_cast<const float4*>(src + (((pj + j) ^ sw) << 4));
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(&raw_free[slot])) : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) av[j] = ra[A_KMAJ ? 0 : j];
      }
      float hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        g3_split(av[j].x, hi[4 * j + 0], lo[4 * j + 0]);
        g3_split(av[j].y, hi[4 * j + 1], lo[4 * j + 1]);
        g3_split(av[j].z, hi[4 * j + 2], lo[4 * j + 2]);
        g3_split(av[j].w, hi[4 * j + 3], lo[4 * j + 3]);
      }
      const uint32_t ta = tlane + tm_a + 32u * (uint32_t)s;
      if (p.dbg != 5) {
        g3_tmem_st16(ta, hi);
        g3_tmem_st16(ta + 16u, lo);
      } else if (hi[3] + lo[7] + hi[9] + lo[15] + hi[0] + lo[1] + hi[12] + lo[4] == 123.456f) {
        g3_tmem_st16(ta, hi);                                 // timing probe: never true, keeps the conversion alive
      }
      G3_STAMP(3);
    }
    // ---- B: pieces -> hi / lo parts of smem stage s
    uint8_t* st = stages + s * stage_bytes;
#pragma unroll
    for (int i = 0; i < NPB; i++) {
      if (b_off[i] < 0) continue;
      float4 h, l;
      g3_split4(rb[i], h, l);
      *reinterpret_cast<float4*>(st + b_off[i]) = h;
      *reinterpret_cast<float4*>(st + bpart + b_off[i]) = l;
    }
    G3_STAMP(4);
    // the next chunk's global loads go out BEFORE the hand-off: issued after the proxy fence they sat behind it for ~2.7 k
    // cycles (in-kernel stamps), a third of a group's iteration
    prefetch(c + G3_GROUPS);
    G3_STAMP(5);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    umma::tc_fence_before_sync();
    G3_STAMP(6);
    umma::warp_arrive_full(&bar_full[s]);
    G3_STAMP(7);
  }
#undef G3_STAMP
  umma::mbar_wait(&bar_done, 0);
  umma::tc_fence_after_sync();
  asm volatile("bar.sync 1, %0;" ::"n"(G3_CONV) : "memory");      // the scratch below aliases stages written by other warps

  // ---- epilogue: 32 x 32 panels, TMEM -> registers -> padded scratch -> lanes along the row ----
  float* scr = reinterpret_cast<float*>(smem + warp * G3_SCRATCH);
  const bool vec_red = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cdst) & 15) == 0) && ((n0 & 3) == 0);
  const int npanels = (ncols + 31) >> 5;
  for (int pn = g; pn < npanels; pn += 4) {
    const int c0 = 32 * pn;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      if (c0 + 8 * b < BN) {                                  // warp-uniform
        float v[8];
        if (nchunks > 0) {
          g3_tmem_ld8x2(tlane + (uint32_t)(c0 + 8 * b), tlane + tm_corr + (uint32_t)(c0 + 8 * b), v);
        } else {
#pragma unroll
          for (int e = 0; e < 8; e++) v[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; e++) scr[lane * 33 + 8 * b + e] = v[e];
      }
    }
    __syncwarp();
    if (p.splits > 1) {
      // split-K: 8 lanes x 4 columns cover a row's 32 columns, 4 rows per instruction, vector reductions
      const int cq = 4 * (lane & 7), rsub = lane >> 3;
      const int n = n0 + c0 + cq;
#pragma unroll 4
      for (int r4 = 0; r4 < 32; r4 += 4) {
        const int r = r4 + rsub, row = m0 + 32 * q + r;
        if (row < p.M && n < ncap && c0 + cq < BN) {
          float* cp = Cdst + (i64)row * p.ldc + n;
          const float* sv = scr + r * 33 + cq;
          if (vec_red && n + 3 < ncap) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(p.alpha * sv[0]), "f"(p.alpha * sv[1]),
                         "f"(p.alpha * sv[2]), "f"(p.alpha * sv[3]) : "memory");
          } else {
#pragma unroll
            for (int e = 0; e < 4; e++)
              if (n + e < ncap) atomicAdd(cp + e, p.alpha * sv[e]);
          }
        }
      }
    } else {
      const int n = n0 + c0 + lane;
      const bool cok = (c0 + lane < BN) && n < ncap;
      const float bv = (cok && biasp) ? biasp[n] : 0.f;
      const int rmax = min(32, p.M - (m0 + 32 * q));
      float* cp = Cdst + (i64)(m0 + 32 * q) * p.ldc + n;
      if (cok) {
        if (beta != 0.f) {
#pragma unroll 8
          for (int r = 0; r < rmax; r++) {
            float t = fmaf(beta, cp[(i64)r * p.ldc], p.alpha * scr[r * 33 + lane]) + bv;
            if (p.act == 1) t = fmaxf(t, 0.f);
            cp[(i64)r * p.ldc] = t;
          }
        } else {
#pragma unroll 8
          for (int r = 0; r < rmax; r++) {
            float t = p.alpha * scr[r * 33 + lane] + bv;
            if (p.act == 1) t = fmaxf(t, 0.f);
            cp[(i64)r * p.ldc] = t;
          }
        }
      }
    }
    __syncwarp();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
}

static int g_g3_dbg = 0;          // profiling aid (mmdfn_gemm_tc_set_variant 30 + x): 1 = no MMAs, 2 / 3 / 4 = column tile <= 112 / 128 / 144
static long long* g_g3_stamps = nullptr;
void umma_gemm3_set_debug(int v) { g_g3_dbg = v; }
void umma_gemm3_set_stamps(long long* device_buf) { g_g3_stamps = device_buf; }

static int g3_pick_bn(int N) {
  const int bmax = g_g3_dbg == 2 ? 112 : g_g3_dbg == 3 ? 128 : g_g3_dbg == 4 ? 144 : G3_BN_MAX;
  const int nt = ceil_div(N, bmax);
  int bn = ceil_div(ceil_div(N, nt), 16) * 16;
  return bn < 16 ? 16 : bn;
}

typedef CUresult (*G3EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static G3EncodeFn g3_encode_fn() {
  static G3EncodeFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (G3EncodeFn)q;
  }
  return fn;
}

template <int MODE>
static int launch_g3(const G3Args& p, cudaStream_t st) {
  static bool configured = false;
  const int stage = 2 * (p.bn / 8) * G3_SBO;
  int smem = p.ns * stage + (MODE != 2 ? G3_NP * G3_PANEL + 1024 : 0);
  if (smem < G3_CONVW * G3_SCRATCH) smem = G3_CONVW * G3_SCRATCH;
  if (!configured) {
    MMDFN_CUDA(cudaFuncSetAttribute(umma_gemm3_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured = true;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (MODE != 2) {
    // A (M, K) row-major, row stride lda: dimension 0 = k (contiguous), dimension 1 = row; box = 32 floats x 128 rows
    G3EncodeFn enc = g3_encode_fn();
    if (!enc) return MMDFN_EINVAL;
    const cuuint64_t gdim[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
    const cuuint64_t gstr[1] = {(cuuint64_t)p.lda * 4};
    const cuuint32_t box[2] = {32, 128};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.A), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMDFN_EINVAL;
  }
  const int ytiles = p.nsplit > 0 ? ceil_div(p.nsplit, p.bn) + ceil_div(p.N - p.nsplit, p.bn) : ceil_div(p.N, p.bn);
  dim3 grid(ceil_div(p.M, 128), ytiles, p.splits > 1 ? p.splits : 1);
  umma_gemm3_kernel<MODE><<<grid, G3_THREADS, smem, st>>>(tm, p);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// K-contiguous operands are read with 16-byte loads: they need 16-byte aligned bases and leading dimensions that are
// multiples of 4 floats; MN-contiguous operands are read with 32-bit loads (no constraint)
bool umma_gemm3_eligible(bool ta, bool tb, int M, int N, int K, const float* A, i64 lda, const float* B, i64 ldb) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (K <= 0 || M <= 0 || N <= 0) return false;
  if (!ta && (!al(A) || (lda & 3))) return false;
  if (tb && (!al(B) || (ldb & 3))) return false;
  return true;
}

// number of K splits this kernel wants for a problem (1 = none): the output has too few tiles for one wave
int umma_gemm3_splits(int M, int N, int K, bool plain_epilogue) {
  const int bn = g3_pick_bn(N);
  const i64 tiles = (i64)ceil_div(M, 128) * ceil_div(N, bn);
  if (!plain_epilogue || tiles * 2 > 148 || K < 512) return 1;
  i64 s = 148 / tiles;
  const i64 smax = ceil_div(K, 128);
  if (s > smax) s = smax;
  return s < 1 ? 1 : (int)s;
}

int umma_gemm3(bool ta, bool tb, int M, int N, int K, float alpha, const float* A, i64 lda, const float* B, i64 ldb,
               float beta, float* C, i64 ldc, const float* bias, int act, int splits, cudaStream_t st) {
  G3Args p{A, lda, B, ldb, C, ldc, bias, M, N, K, alpha, beta, act, splits, 0, 0, nullptr, nullptr, nullptr, 0.f, 0, 0, g_g3_dbg, g_g3_stamps};
  p.bn = g3_pick_bn(N);
  p.ns = (512 - 2 * p.bn) / 32;
  if (p.ns > G3_NS_MAX) p.ns = G3_NS_MAX;
  while (p.ns > 2 && p.ns * 2 * (p.bn / 8) * G3_SBO + (ta ? 0 : G3_NP * G3_PANEL + 1024) > 226 * 1024) p.ns--;
  if (!ta && tb) return launch_g3<0>(p, st);
  if (!ta && !tb) return launch_g3<1>(p, st);
  return launch_g3<2>(p, st);
}

// C[:, :N1] = A B1^T + bias1 and C[:, N1:N1+N2] = A B2^T + bias2 in ONE launch (the two directions' input-gate products
// of a GRU layer: same A, two (300, 200) weights).  Caller checked umma_gemm3_eligible for (A, B1) and B2's alignment.
int umma_gemm3_nt_pair(int M, int N1, int N2, int K, const float* A, i64 lda, const float* B1, const float* B2, i64 ldb,
                       float* C, i64 ldc, const float* bias1, const float* bias2, cudaStream_t st) {
  G3Args p{A, lda, B1, ldb, C, ldc, bias1, M, N1 + N2, K, 1.f, 0.f, 0, 1, 0, 0, B2, bias2, C + N1, 0.f, N1, 0, g_g3_dbg, g_g3_stamps};
  p.bn = g3_pick_bn(N1 > N2 ? N1 : N2);
  p.ns = (512 - 2 * p.bn) / 32;
  if (p.ns > G3_NS_MAX) p.ns = G3_NS_MAX;
  while (p.ns > 2 && p.ns * 2 * (p.bn / 8) * G3_SBO + G3_NP * G3_PANEL + 1024 > 226 * 1024) p.ns--;
  return launch_g3<0>(p, st);
}

// C1 = op(A) B1 + beta1 C1 and C2 = op(A) B2 + beta2 C2 in ONE launch for the NN (ta = false) and TN (ta = true) forms: the
// same A against two N-contiguous operands of N columns each, two separate outputs with the same row stride (the LSTM
// gate's input / recurrent gradients and weight gradients of a graph layer).  `splits` as from umma_gemm3_splits for
// (M, 2 N, K); outputs pre-scaled by the caller when splits > 1.
int umma_gemm3_npair(bool ta, int M, int N, int K, const float* A, i64 lda, const float* B1, const float* B2, i64 ldb,
                     float beta1, float* C1, float beta2, float* C2, i64 ldc, int splits, cudaStream_t st) {
  G3Args p{A, lda, B1, ldb, C1, ldc, nullptr, M, 2 * N, K, 1.f, beta1, 0, splits, 0, 0, B2, nullptr, C2, beta2, N, 0, g_g3_dbg, g_g3_stamps};
  p.bn = g3_pick_bn(N);
  p.ns = (512 - 2 * p.bn) / 32;
  if (p.ns > G3_NS_MAX) p.ns = G3_NS_MAX;
  while (p.ns > 2 && p.ns * 2 * (p.bn / 8) * G3_SBO + (ta ? 0 : G3_NP * G3_PANEL + 1024) > 226 * 1024) p.ns--;
  return ta ? launch_g3<2>(p, st) : launch_g3<1>(p, st);
}

// C = A[:, :K1] B1 + A[:, K1:K1+K2] B2 (+ beta C) in ONE launch, B1 (K1, N) and B2 (K2, N) row-major with the same row
// stride (the input gradient of a bidirectional layer: dx = dgates_f W_f + dgates_b W_b).  K1 must be a multiple of 4.
int umma_gemm3_nn_kpair(int M, int N, int K1, int K2, const float* A, i64 lda, const float* B1, const float* B2, i64 ldb,
                        float beta, float* C, i64 ldc, cudaStream_t st) {
  G3Args p{A, lda, B1, ldb, C, ldc, nullptr, M, N, K1 + K2, 1.f, beta, 0, 1, 0, 0, B2, nullptr, nullptr, 0.f, 0, K1, g_g3_dbg, g_g3_stamps};
  p.bn = g3_pick_bn(N);
  p.ns = (512 - 2 * p.bn) / 32;
  if (p.ns > G3_NS_MAX) p.ns = G3_NS_MAX;
  while (p.ns > 2 && p.ns * 2 * (p.bn / 8) * G3_SBO + G3_NP * G3_PANEL + 1024 > 226 * 1024) p.ns--;
  return launch_g3<1>(p, st);
}

}  // namespace mmdfn
