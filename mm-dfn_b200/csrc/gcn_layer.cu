// k6 fused: one launch = one whole GraphConvolution layer (code/model_GCN.py:176-189 inside the GCNII_lyc loop,
// :461-472) for every (dialogue, modality) block of the batch, on the tensor cores (tcgen05.mma kind::tf32, 3-term split
// = fp32-level accuracy, accumulators in TMEM).  Per 128-row tile of a block the CTA chains two products and a fused
// epilogue without leaving the SM:
//
//   phase A   T = A_hat[tile rows, :] . Zin[block rows]   (+ the two cross-modal diagonal terms)      "message aggregate"
//   phase B   U = T . Mw                                   (Mw = 100 x 100, pre-split image, see below)
//   forward   z_out = dropout(relu(U + R)) (+ q);  flags = [relu and keep]     (R = h0 . Mbot, one GEMM for all layers)
//   backward  d_in  = U (+ add);  T is also written out (dW_top = theta Zin^T T needs it)
//
// Algebra.  The reference layer is  u = theta [hi | h0] W + (1-theta) ((1-alpha) hi + alpha h0),  hi = A_hat zin.
// With  Mtop = theta W[0:100] + (1-theta)(1-alpha) I  and  Mbot = theta W[100:200] + (1-theta) alpha I  this is
// u = hi Mtop + h0 Mbot: the alpha/theta mixes fold into the weight operand, the h0 term (layer-invariant input) becomes
// ONE GEMM for all layers (R_all = h0 [Mbot_1 | .. | Mbot_K]) and the fused kernel's second product has K = 100, not 200.
// Backward (A_hat symmetric): du = dz * flags * scale;  t = A_hat du;  d zin = t Mtop^T  -- the same two chained products
// with the transposed weight image;  dMtop = hi^T du = zin^T t, so `hi` is never stored.
//
// Operands.  A_hat rows are K-contiguous: 16-byte pieces global -> registers (4 chunks ahead) -> hi/lo -> UMMA K-major
// stage.  Zin's rows are the contraction index: 16-row chunks arrive RAW through TMA bulk copies into an 8-slot ring that
// lives in the (not yet needed) T-tile region and are read back transposed, conflict-free.  T goes TMEM -> registers ->
// row-major tile in shared memory (cross terms added there) and is re-read as the A operand of phase B; the weight
// operand of phase B is a PRE-SPLIT image in the stage layout (built once per step by gcn_prep_kernel), copied with
// plain 16-byte loads prefetched one k-step ahead.  8 converter warps + 1 issuer warp, a 4-stage operand ring of single
// 8-wide k-steps (the converter -> tensor-core hand-off latency, ~1.1 k cycles, hides behind three queued k-steps), 112.6 KB of
// shared memory and 256 TMEM columns per CTA -> two CTAs per SM, so one tile's loads/epilogues overlap the other's MMAs.
#include "umma.cuh"
#include "internal.cuh"
#include "gcn_layer.cuh"
#include <math.h>

namespace mmdfn {

constexpr int GL_G = 100, GL_BN = 112, GL_ROWS = 128;
constexpr int GL_TILE = GL_ROWS * GL_G * 4;                // 51200 B: T tile / output tile / (before that) the raw z ring
constexpr int GL_ZROWS = 16;                               // rows per raw-ring slot
constexpr int GL_ZSLOT = GL_ZROWS * GL_G * 4;              // 6400 B
constexpr int GL_ZR = GL_TILE / GL_ZSLOT;                  // 8 ring slots
constexpr int GL_CONV = 256, GL_THREADS = GL_CONV + 32;
constexpr int GL_CORR = 128, GL_TMEM = 256;
constexpr int GL_LBO = 128;                                // bytes between the core matrices of consecutive k-quads

// Operand-stage geometry for a K chunk of KC (8 or 16) columns and NS stages.  An 8-row group holds KC/4 core matrices
// (128 B each) and is padded by 16 B: with a power-of-two group stride the tensor core's operand fetch was measured at
// ~255 cycles per 128x112x8 MMA (r02 phase stamps), with the padded stride at ~117 (umma_gemm.cu's layout).
template <int KC, int NS>
struct GLGeo {
  static constexpr int SBO = (KC / 4) * 128 + 16;
  static constexpr int A_PART = 16 * SBO;
  static constexpr int B_PART = (GL_BN / 8) * SBO;
  static constexpr int STAGE = 2 * (A_PART + B_PART);
  static constexpr int SMEM = GL_TILE + NS * STAGE;
  static constexpr int WCHUNKS = (GL_G + KC - 1) / KC;     // K chunks of the 100-deep second product
  static constexpr int KPAD = WCHUNKS * KC;
  static constexpr int IMG_CHUNK = 2 * B_PART / 4;         // floats per chunk of a weight image (hi part | lo part, stage layout)
  static constexpr int IMG = WCHUNKS * IMG_CHUNK;
  static constexpr int NPA = KC / 8;                       // A pieces per converter thread per chunk
  static constexpr int NPB = (GL_BN * (KC / 4) + GL_CONV - 1) / GL_CONV;
  static constexpr int WP = (2 * B_PART / 16 + GL_CONV - 1) / GL_CONV;
  static constexpr int ADEPTH = 64 / KC;                   // chunks of A_hat held in registers ahead of their conversion
  static_assert(ADEPTH % NS == 0 || NS % ADEPTH == 0 || true, "");
};
// default configuration (see gcn_layer_variant): 16-wide chunks, 2 stages
constexpr int GL_KC0 = 16, GL_NS0 = 2;
constexpr int GL_KC1 = 8, GL_NS1 = 3;

__device__ __forceinline__ void gl_split_store(const float4 v, uint8_t* hi_dst, uint8_t* lo_dst) {
  float4 h, l;
  gl_split(v.x, h.x, l.x);
  gl_split(v.y, h.y, l.y);
  gl_split(v.z, h.z, l.z);
  gl_split(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi_dst) = h;
  *reinterpret_cast<float4*>(lo_dst) = l;
}

// FWD = true: forward epilogue (R, relu, mask, flags, +q); false: backward epilogue (+ add)
template <bool FWD, int KC, int NS>
__global__ void __launch_bounds__(GL_THREADS, 2) gcn_layer_kernel(GcnLayerArgs p) {
  using GEO = GLGeo<KC, NS>;
  constexpr int SBO = GEO::SBO, A_PART = GEO::A_PART, B_PART = GEO::B_PART, STAGE = GEO::STAGE;
  constexpr int WCHUNKS = GEO::WCHUNKS, NPA = GEO::NPA, NPB = GEO::NPB, WP = GEO::WP, ADEPTH = GEO::ADEPTH;
  constexpr int KQ = KC / 4;                                 // k-quads (16-byte pieces) per row per chunk
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_free[NS];
  __shared__ __align__(8) uint64_t bar_full[NS];
  __shared__ __align__(8) uint64_t bar_z[GL_ZR];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / 3, m = blockIdx.x % 3;
  const int off = p.dia_off[b], L = p.dia_off[b + 1] - off;
  const int r0 = blockIdx.y * GL_ROWS;
  if (r0 >= L) return;                                       // whole CTA, nothing allocated yet
  const int nrows = min(GL_ROWS, L - r0);
  const float* A = p.adj_blk + p.blk_off[b] + (i64)m * L * L + (i64)r0 * L;
  const float* Z = p.zin + ((i64)m * p.N + off) * p.ldz;
  const int nchA = (L + KC - 1) / KC;                        // K chunks of phase A
  const int nzc = (L + GL_ZROWS - 1) / GL_ZROWS;              // 16-row raw chunks of the z block
  const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
  int dbg_n = 0;
#define GL_STAMP() do { if (dbg_on && dbg_n < 60) p.dbg[dbg_n++] = clock64(); } while (0)
  GL_STAMP();                                                // [0] entry

  float* tile = reinterpret_cast<float*>(smem);              // raw z ring first, T tile / output tile later
  uint8_t* stages = smem + GL_TILE;
  const bool z_dense = (p.ldz == GL_G);
  // 16-row chunk zc of the z block -> ring slot zc % GL_ZR.  Dense rows: ONE bulk copy (lane 0).  Strided rows (a column
  // block of a wider matrix): one 400-byte bulk copy per row, issued by lanes 0..15 of the calling warp.
  auto issue_z = [&](int zc, int ln) {
    const int j0 = zc * GL_ZROWS;
    const int rows = min(L, j0 + GL_ZROWS) - j0;
    const uint32_t bar = umma::smem_u32(&bar_z[zc % GL_ZR]);
    float* dst = tile + (zc % GL_ZR) * (GL_ZROWS * GL_G);
    if (ln == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)rows * GL_G * 4) : "memory");
    if (z_dense) {
      if (ln == 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(umma::smem_u32(dst)), "l"(Z + (i64)j0 * GL_G), "r"((uint32_t)rows * GL_G * 4), "r"(bar) : "memory");
    } else if (ln < rows) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(umma::smem_u32(dst + ln * GL_G)), "l"(Z + (i64)(j0 + ln) * p.ldz), "r"((uint32_t)GL_G * 4), "r"(bar) : "memory");
    }
  };
  if (warp == 8) umma::tmem_alloc(&tmem_base_s, GL_TMEM);
  if (tid == 0) {
    for (int s = 0; s < NS; s++) {
      umma::mbar_init(&bar_free[s], 1);
      umma::mbar_init(&bar_full[s], GL_CONV / 32);
    }
    for (int s = 0; s < GL_ZR; s++) umma::mbar_init(&bar_z[s], 1);
    umma::fence_barrier_init();
  }
  constexpr uint32_t IDESC = umma::idesc_tf32(128, GL_BN);

  // A pieces of this thread (16 bytes each, NPA per chunk): piece index pi = tid + 256 i, quarter-warp qw = pi / 8 ->
  // row group qw / KQ, k-quad qw % KQ; the 8 lanes of a quarter-warp write one whole core matrix (128 contiguous bytes)
  int row_a[NPA], kq_a[NPA], o_a[NPA];
#pragma unroll
  for (int i = 0; i < NPA; i++) {
    const int qw = (tid + GL_CONV * i) >> 3;
    kq_a[i] = qw % KQ;
    row_a[i] = (qw / KQ) * 8 + (tid & 7);
    o_a[i] = (qw / KQ) * SBO + kq_a[i] * GL_LBO + (tid & 7) * 16;
  }
  const bool a_vec = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  auto load_a = [&](int c, float4 (&va)[NPA]) {
#pragma unroll
    for (int i = 0; i < NPA; i++) {
      const int k = c * KC + 4 * kq_a[i];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_a[i] < nrows && k < L) {
        const float* q = A + (i64)row_a[i] * L + k;
        if (a_vec) {
          v = __ldg(reinterpret_cast<const float4*>(q));
        } else {
          v.x = q[0];
          if (k + 1 < L) v.y = q[1];
          if (k + 2 < L) v.z = q[2];
          if (k + 3 < L) v.w = q[3];
        }
      }
      va[i] = v;
    }
  };
  const int o1 = (m == 0) ? 1 : 0, o2 = (m == 2) ? 1 : 2;
  // cross-modal diagonal entries of this tile's rows (read in the T-tile pass)
  const float* dg1 = p.adj_diag + (i64)(min(m, o1) + max(m, o1) - 1) * p.N + off + r0;
  const float* dg2 = p.adj_diag + (i64)(min(m, o2) + max(m, o2) - 1) * p.N + off + r0;
  float4 va[ADEPTH][NPA];
  if (warp < 8) {
#pragma unroll
    for (int u = 0; u < ADEPTH; u++)
      if (u < nchA) load_a(u, va[u]);
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  GL_STAMP();                                                // [1] set-up done

  if (warp == 8) {
    // ===== MMA issuer (lane 0) + z producer (lanes 0..15) =====
    for (int zc = 0; zc < GL_ZR && zc < nzc; zc++) issue_z(zc, lane);
    const int total = nchA + WCHUNKS;
    const uint64_t d0 = umma::smem_desc(umma::smem_u32(stages), GL_LBO, SBO);
    const uint32_t dhi = (uint32_t)(d0 >> 32), dlo = (uint32_t)d0;
    for (int g = 0; g < total; g++) {
      const int s = g % NS;
      umma::mbar_wait(&bar_full[s], (uint32_t)((g / NS) & 1));           // the whole warp: the MMA issue below is convergent
      umma::tc_fence_after_sync();
      if (lane == 0 && p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && g < 60) p.dbg[64 + g] = clock64();   // issuer: chunk g full
      // chunk g read the LAST rows of raw chunk zc: every converter has finished with that ring slot (it arrived on
      // bar_full after its reads) -> refill it with the raw chunk one ring revolution ahead
      if (g < nchA && ((g + 1) * KC) % GL_ZROWS == 0) {
        const int zc = ((g + 1) * KC) / GL_ZROWS - 1;
        if (zc + GL_ZR < nzc) issue_z(zc + GL_ZR, lane);
      }
      __syncwarp();
      const uint32_t o = dlo + (uint32_t)s * (STAGE >> 4);
#pragma unroll
      for (int j = 0; j < KC / 8; j++) {
        const uint32_t oj = o + (uint32_t)j * ((2 * GL_LBO) >> 4);
        // each phase starts a fresh accumulator
        umma::kstep3_elect(tmem, tmem + GL_CORR, dhi, oj, oj + (A_PART >> 4), oj + ((2 * A_PART) >> 4), oj + ((2 * A_PART + B_PART) >> 4),
                           IDESC, ((g > 0 && g != nchA) || j > 0) ? 1u : 0u);
      }
      umma::mma_commit_elect(&bar_free[s]);
    }
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tmem_dealloc(tmem, GL_TMEM);
    return;
  }

  // ===== converters (warps 0-7) =====
  // ---------------- phase A: T = A_hat[tile rows, :] * Zin[block] ----------------
  // B pieces of this thread: piece index pi = tid + 256 i < 112 KQ -> feature column pi % 112, k-quad pi / 112
  int c_b[NPB], kq_b[NPB], o_b[NPB];
#pragma unroll
  for (int i = 0; i < NPB; i++) {
    const int pi = tid + GL_CONV * i;
    kq_b[i] = pi / GL_BN;
    c_b[i] = pi - GL_BN * kq_b[i];
    o_b[i] = (pi < GL_BN * KQ) ? (c_b[i] >> 3) * SBO + kq_b[i] * GL_LBO + (c_b[i] & 7) * 16 : -1;
  }
  for (int c0 = 0; c0 < nchA; c0 += ADEPTH) {
#pragma unroll
    for (int u = 0; u < ADEPTH; u++) {
      const int c = c0 + u;
      if (c < nchA) {
        const int s = c % NS;
        const int zc = (c * KC) / GL_ZROWS;
        umma::mbar_wait(&bar_z[zc % GL_ZR], (uint32_t)((zc / GL_ZR) & 1));
        if (dbg_on && c < 60) p.dbg[192 + c] = clock64();                            // converter: z landed
        if (c >= NS) umma::mbar_wait(&bar_free[s], (uint32_t)(((c / NS) - 1) & 1));
        if (dbg_on && c < 60) p.dbg[128 + c] = clock64();                            // converter: stage free
        uint8_t* st = stages + s * STAGE;
        const bool fine = dbg_on && c == 3;
        if (fine) p.dbg[240] = clock64();
#pragma unroll
        for (int i = 0; i < NPA; i++) gl_split_store(va[u][i], st + o_a[i], st + A_PART + o_a[i]);
        if (fine) p.dbg[241] = clock64();
#pragma unroll
        for (int i = 0; i < NPB; i++) {
          if (o_b[i] >= 0) {
            const int jr = (c * KC) % GL_ZROWS + 4 * kq_b[i];      // first row of the piece inside the 16-row slot
            const int jleft = L - zc * GL_ZROWS - jr;              // valid rows from there on
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c_b[i] < GL_G) {                                   // columns 100..111 pad N; rows beyond L are masked
              const float* q = tile + (zc % GL_ZR) * (GL_ZROWS * GL_G) + jr * GL_G + c_b[i];
              if (jleft > 0) v.x = q[0];
              if (jleft > 1) v.y = q[GL_G];
              if (jleft > 2) v.z = q[2 * GL_G];
              if (jleft > 3) v.w = q[3 * GL_G];
            }
            gl_split_store(v, st + 2 * A_PART + o_b[i], st + 2 * A_PART + B_PART + o_b[i]);
          }
        }
        if (fine) p.dbg[242] = clock64();
        umma::fence_proxy_async_smem();
        if (fine) p.dbg[243] = clock64();
        __syncwarp();
        if (fine) p.dbg[244] = clock64();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(&bar_full[s])) : "memory");
        if (fine) p.dbg[245] = clock64();
        // prefetch AFTER the hand-off: fence.proxy.async waits for the thread's outstanding global loads (measured:
        // ~1 k cycles per chunk when the next pieces were requested before it)
        if (c + ADEPTH < nchA) load_a(c + ADEPTH, va[u]);
      }
    }
  }
  GL_STAMP();                                                // [2] phase A converted

  // weight-image pieces of phase-B chunk 0 and the first batch of cross-modal rows are requested before the wait
  const float4* wimg4 = reinterpret_cast<const float4*>(p.wimg);
  float4 wb[WP];
  auto load_w = [&](int c) {
#pragma unroll
    for (int u = 0; u < WP; u++) {
      const int i = tid + u * GL_CONV;
      if (i < 2 * B_PART / 16) wb[u] = __ldg(wimg4 + (i64)c * (2 * B_PART / 16) + i);
    }
  };
  load_w(0);
  const float* x1 = p.zin + ((i64)o1 * p.N + off + r0) * p.ldz;
  const float* x2 = p.zin + ((i64)o2 * p.N + off + r0) * p.ldz;
  const int total4 = nrows * (GL_G / 4);
  constexpr int EB = 4;
  float4 a1[EB], a2[EB];
#pragma unroll
  for (int u = 0; u < EB; u++) {
    const int i = tid + u * GL_CONV;
    if (i < total4) {
      const int r = i / (GL_G / 4), c4 = i - r * (GL_G / 4);
      a1[u] = __ldg(reinterpret_cast<const float4*>(x1 + (i64)r * p.ldz) + c4);
      a2[u] = __ldg(reinterpret_cast<const float4*>(x2 + (i64)r * p.ldz) + c4);
    }
  }
  {
    const int last = nchA - 1;
    umma::mbar_wait(&bar_free[last % NS], (uint32_t)((last / NS) & 1));
  }
  umma::tc_fence_after_sync();
  GL_STAMP();                                                // [3] phase A MMAs done

  // TMEM -> row-major tile (thread = row; warps 0-3 take columns 0..63, warps 4-7 columns 64..99).  Every converter
  // passed the last full barrier before the final commit could fire, so nobody still reads the raw ring.
  auto tmem_to_tile = [&]() {
    const int r = (warp & 3) * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int cb_begin = (warp < 4) ? 0 : 64, cb_end = (warp < 4) ? 64 : GL_BN;
#pragma unroll 1
    for (int cb = cb_begin; cb < cb_end; cb += 16) {
      if (cb >= GL_G) break;
      float v[16], w[16];
      umma::tmem_ld16x2(taddr + cb, taddr + GL_CORR + cb, v, w);
#pragma unroll
      for (int q4 = 0; q4 < 16; q4 += 4) {
        if (cb + q4 < GL_G)
          *reinterpret_cast<float4*>(tile + r * GL_G + cb + q4) =
              make_float4(v[q4] + w[q4], v[q4 + 1] + w[q4 + 1], v[q4 + 2] + w[q4 + 2], v[q4 + 3] + w[q4 + 3]);
      }
    }
  };
  tmem_to_tile();
  umma::tc_fence_before_sync();                              // the TMEM reads are ordered before phase B's first MMA
  asm volatile("bar.sync 1, 256;" ::: "memory");
  {
    // cross-modal terms (coalesced pass over the tile) and the optional T output
    float4* tile4 = reinterpret_cast<float4*>(tile);
    float* tout = p.t_out ? p.t_out + ((i64)m * p.N + off + r0) * p.ldt : nullptr;
#pragma unroll 1
    for (int base = tid; base < total4; base += EB * GL_CONV) {
      if (base != tid) {
#pragma unroll
        for (int u = 0; u < EB; u++) {
          const int i = base + u * GL_CONV;
          if (i < total4) {
            const int r = i / (GL_G / 4), c4 = i - r * (GL_G / 4);
            a1[u] = __ldg(reinterpret_cast<const float4*>(x1 + (i64)r * p.ldz) + c4);
            a2[u] = __ldg(reinterpret_cast<const float4*>(x2 + (i64)r * p.ldz) + c4);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < EB; u++) {
        const int i = base + u * GL_CONV;
        if (i < total4) {
          const int r = i / (GL_G / 4), c4 = i - r * (GL_G / 4);
          const float e1 = __ldg(dg1 + r), e2 = __ldg(dg2 + r);
          const float4 t = tile4[i];
          const float4 o = make_float4(t.x + e1 * a1[u].x + e2 * a2[u].x, t.y + e1 * a1[u].y + e2 * a2[u].y,
                                       t.z + e1 * a1[u].z + e2 * a2[u].z, t.w + e1 * a1[u].w + e2 * a2[u].w);
          tile4[i] = o;
          if (tout) *(reinterpret_cast<float4*>(tout + (i64)r * p.ldt) + c4) = o;
        }
      }
    }
    // rows nrows..127 of the tile feed the A operand of phase B: zero them (their results are never stored)
    for (int i = total4 + tid; i < GL_ROWS * (GL_G / 4); i += GL_CONV) tile4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  GL_STAMP();                                                // [4] T tile complete

  // ---------------- phase B: U = T * Mw ----------------
#pragma unroll 1
  for (int c = 0; c < WCHUNKS; c++) {
    const int g = nchA + c, s = g % NS;
    // A pieces from the T tile (k >= 100 reads as zero)
    float4 ta[NPA];
#pragma unroll
    for (int i = 0; i < NPA; i++) {
      const int k = c * KC + 4 * kq_a[i];
      ta[i] = (k < GL_G) ? *reinterpret_cast<const float4*>(tile + row_a[i] * GL_G + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (dbg_on && g < 60) p.dbg[192 + g] = clock64();
    if (g >= NS) umma::mbar_wait(&bar_free[s], (uint32_t)(((g / NS) - 1) & 1));
    if (dbg_on && g < 60) p.dbg[128 + g] = clock64();
    uint8_t* st = stages + s * STAGE;
    const bool fine = dbg_on && c == 3;
    if (fine) p.dbg[248] = clock64();
#pragma unroll
    for (int i = 0; i < NPA; i++) gl_split_store(ta[i], st + o_a[i], st + A_PART + o_a[i]);
    if (fine) p.dbg[249] = clock64();
#pragma unroll
    for (int u = 0; u < WP; u++) {
      const int i = tid + u * GL_CONV;
      if (i < 2 * B_PART / 16) *reinterpret_cast<float4*>(st + 2 * A_PART + i * 16) = wb[u];
    }
    if (fine) p.dbg[250] = clock64();
    umma::fence_proxy_async_smem();
    if (fine) p.dbg[251] = clock64();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(&bar_full[s])) : "memory");
    if (fine) p.dbg[252] = clock64();
    if (c + 1 < WCHUNKS) load_w(c + 1);                      // after the fence (see phase A)
  }
  GL_STAMP();                                                // [5] phase B converted

  // ---------------- epilogue ----------------
  // operands of the fused epilogue: first batch requested before the wait for the last MMA
  const i64 row0 = (i64)m * p.N + off + r0;
  float4 e1v[EB], e2v[EB];
  uint32_t mk[EB];
  auto load_epi = [&](int base) {
#pragma unroll
    for (int u = 0; u < EB; u++) {
      const int i = base + u * GL_CONV;
      if (i < total4) {
        const int r = i / (GL_G / 4), c4 = i - r * (GL_G / 4);
        if (FWD) {
          e1v[u] = __ldg(reinterpret_cast<const float4*>(p.r + (row0 + r) * p.ldr) + c4);
          if (p.q) e2v[u] = __ldg(reinterpret_cast<const float4*>(p.q + (row0 + r) * GL_G) + c4);
          if (p.mask) mk[u] = __ldg(reinterpret_cast<const uint32_t*>(p.mask + (row0 + r) * GL_G) + c4);
        } else {
          if (p.add) e1v[u] = __ldg(reinterpret_cast<const float4*>(p.add + (row0 + r) * GL_G) + c4);
        }
      }
    }
  };
  load_epi(tid);
  {
    const int last = nchA + WCHUNKS - 1;
    umma::mbar_wait(&bar_free[last % NS], (uint32_t)((last / NS) & 1));
  }
  umma::tc_fence_after_sync();
  GL_STAMP();                                                // [6] phase B MMAs done
  tmem_to_tile();                                            // every converter finished reading the T tile before the
  asm volatile("bar.sync 1, 256;" ::: "memory");             // last full barrier, i.e. before the last commit could fire
  {
    const float4* tile4 = reinterpret_cast<const float4*>(tile);
#pragma unroll 1
    for (int base = tid; base < total4; base += EB * GL_CONV) {
      if (base != tid) load_epi(base);
#pragma unroll
      for (int u = 0; u < EB; u++) {
        const int i = base + u * GL_CONV;
        if (i < total4) {
          const int r = i / (GL_G / 4), c4 = i - r * (GL_G / 4);
          const float4 t = tile4[i];
          float o[4] = {t.x, t.y, t.z, t.w};
          if (FWD) {
            const float rr[4] = {e1v[u].x, e1v[u].y, e1v[u].z, e1v[u].w};
            uint32_t fl = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              float v = o[j] + rr[j];
              bool on = v > 0.f;
              if (p.mask) on = on && ((mk[u] >> (8 * j)) & 0xFFu);
              o[j] = on ? v * p.scale : 0.f;
              fl |= (on ? 1u : 0u) << (8 * j);
            }
            if (p.q) { o[0] += e2v[u].x; o[1] += e2v[u].y; o[2] += e2v[u].z; o[3] += e2v[u].w; }
            *(reinterpret_cast<uint32_t*>(p.flags + (row0 + r) * GL_G) + c4) = fl;
          } else if (p.add) {
            o[0] += e1v[u].x; o[1] += e1v[u].y; o[2] += e1v[u].z; o[3] += e1v[u].w;
          }
          *(reinterpret_cast<float4*>(p.out + (row0 + r) * p.ldo) + c4) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
  GL_STAMP();                                                // [7] stored
  if (dbg_on) p.dbg[63] = dbg_n;
#undef GL_STAMP
  umma::tc_fence_before_sync();
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------
// Per-step weight preparation for up to GL_PREP_MAX layers per launch: folded matrices and their pre-split images.
//   mtop_all / mbot_all : (100, 100 K) row-major, column block l = Mtop_l / Mbot_l
//   img_f[l] : phase-B operand of the forward   (B(n, k) = Mtop_l[k][n])
//   img_b[l] : phase-B operand of the backward  (B(n, k) = Mtop_l[n][k])
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GL_PREP_MAX = 32;
struct GcnPrepArgs {
  const float* w[GL_PREP_MAX];        // convs[l].weight (200, 100)
  float theta[GL_PREP_MAX];
  float alpha;
  int l0, nl, K;
  int kc, sbo, b_part, kpad, img;     // image geometry of the layer kernel's configuration (GLGeo)
  float* mtop_all;
  float* mbot_all;
  float* img_f;                       // K images
  float* img_b;
};

__global__ void gcn_prep_kernel(GcnPrepArgs p) {
  const int li = blockIdx.y, l = p.l0 + li;
  const float* W = p.w[li];
  const float th = p.theta[li], c1 = (1.f - th) * (1.f - p.alpha), c2 = (1.f - th) * p.alpha;
  const i64 ldm = (i64)GL_G * p.K;
  const int img_chunk = 2 * p.b_part / 4;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < GL_BN * p.kpad; idx += gridDim.x * blockDim.x) {
    const int n = idx / p.kpad, k = idx - n * p.kpad;        // image element (n, k)
    float top_kn = 0.f, top_nk = 0.f;
    if (n < GL_G && k < GL_G) {
      top_kn = th * W[k * GL_G + n] + (k == n ? c1 : 0.f);
      top_nk = th * W[n * GL_G + k] + (k == n ? c1 : 0.f);
      p.mtop_all[(i64)n * ldm + (i64)l * GL_G + k] = top_nk;
      p.mbot_all[(i64)n * ldm + (i64)l * GL_G + k] = th * W[(GL_G + n) * GL_G + k] + (k == n ? c2 : 0.f);
    }
    const int ch = k / p.kc, kq = (k % p.kc) >> 2;
    const int o = ch * img_chunk + ((n >> 3) * p.sbo + kq * GL_LBO + (n & 7) * 16 + (k & 3) * 4) / 4;
    float hi, lo;
    gl_split(top_kn, hi, lo);
    p.img_f[(i64)l * p.img + o] = hi;
    p.img_f[(i64)l * p.img + o + p.b_part / 4] = lo;
    gl_split(top_nk, hi, lo);
    p.img_b[(i64)l * p.img + o] = hi;
    p.img_b[(i64)l * p.img + o + p.b_part / 4] = lo;
  }
}

// dconvW[l] (=/+=) theta_l [dMtop_l ; dMbot_l]
struct GcnUnfoldArgs {
  float* dw[GL_PREP_MAX];
  float theta[GL_PREP_MAX];
  int l0, K, accumulate;
  const float* dmtop_all;
  const float* dmbot_all;
};

__global__ void gcn_unfold_kernel(GcnUnfoldArgs p) {
  const int li = blockIdx.y, l = p.l0 + li;
  const i64 ldm = (i64)GL_G * p.K;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * GL_G * GL_G; idx += gridDim.x * blockDim.x) {
    const int k = idx / GL_G, n = idx - k * GL_G;
    const float g = k < GL_G ? p.dmtop_all[(i64)k * ldm + (i64)l * GL_G + n] : p.dmbot_all[(i64)(k - GL_G) * ldm + (i64)l * GL_G + n];
    float* d = p.dw[li] + idx;
    *d = p.accumulate ? *d + p.theta[li] * g : p.theta[li] * g;
  }
}

static long long* g_gl_dbg = nullptr;
// First-generation configuration: 0 = 16-wide K chunks, 2 operand stages; 1 = 8-wide chunks, 3 stages.  Process-global
// timing knob (tools/): the weight images are laid out for the configuration that was current when gcn_layer_prep built
// them.  g_gl_gen2: use the second-generation kernel (gcn_layer2.cu, same 16-wide image) wherever it is eligible.
static int g_gl_variant = 0;
static bool g_gl_gen2 = true;

long long gcn_layer_img_floats() { return g_gl_variant == 0 ? GLGeo<GL_KC0, GL_NS0>::IMG : GLGeo<GL_KC1, GL_NS1>::IMG; }

// theta_l = ln(lamda / l + 1) in Python float (double) arithmetic, l = 1-based layer index   (code/model_GCN.py:177)
static inline double theta_of(double lamda, int l0) { return log(lamda / (double)(l0 + 1) + 1.0); }

template <class GEO, int KC>
static void prep_geo(GcnPrepArgs& a) {
  a.kc = KC; a.sbo = GEO::SBO; a.b_part = GEO::B_PART; a.kpad = GEO::KPAD; a.img = GEO::IMG;
}

int gcn_layer_prep(int K, const float* const* convW, double lamda, double alpha, float* mtop_all, float* mbot_all,
                   float* img_f, float* img_b, cudaStream_t st) {
  for (int l0 = 0; l0 < K; l0 += GL_PREP_MAX) {
    GcnPrepArgs a;
    a.nl = K - l0 < GL_PREP_MAX ? K - l0 : GL_PREP_MAX;
    for (int i = 0; i < a.nl; i++) {
      a.w[i] = convW[l0 + i];
      a.theta[i] = (float)theta_of(lamda, l0 + i);
    }
    a.alpha = (float)alpha;
    a.l0 = l0;
    a.K = K;
    if (g_gl_variant == 0) prep_geo<GLGeo<GL_KC0, GL_NS0>, GL_KC0>(a);
    else prep_geo<GLGeo<GL_KC1, GL_NS1>, GL_KC1>(a);
    // zero-fill the images first: the 16 pad bytes of every 8-row group are copied into the operand stage as they are
    MMDFN_CUDA(cudaMemsetAsync(img_f + (i64)l0 * a.img, 0, (size_t)a.nl * a.img * sizeof(float), st));
    MMDFN_CUDA(cudaMemsetAsync(img_b + (i64)l0 * a.img, 0, (size_t)a.nl * a.img * sizeof(float), st));
    a.mtop_all = mtop_all;
    a.mbot_all = mbot_all;
    a.img_f = img_f;
    a.img_b = img_b;
    gcn_prep_kernel<<<dim3(13, a.nl), 256, 0, st>>>(a);
    MMDFN_LAUNCH_CHECK();
  }
  return 0;
}

int gcn_layer_unfold(int K, float* const* dconvW, double lamda, const float* dmtop_all, const float* dmbot_all,
                     int accumulate, cudaStream_t st) {
  for (int l0 = 0; l0 < K; l0 += GL_PREP_MAX) {
    GcnUnfoldArgs a;
    const int nl = K - l0 < GL_PREP_MAX ? K - l0 : GL_PREP_MAX;
    for (int i = 0; i < nl; i++) {
      a.dw[i] = dconvW[l0 + i];
      a.theta[i] = (float)theta_of(lamda, l0 + i);
    }
    a.l0 = l0;
    a.K = K;
    a.accumulate = accumulate;
    a.dmtop_all = dmtop_all;
    a.dmbot_all = dmbot_all;
    gcn_unfold_kernel<<<dim3(20, nl), 256, 0, st>>>(a);
    MMDFN_LAUNCH_CHECK();
  }
  return 0;
}

template <bool FWD, int KC, int NS>
static int launch_layer_cfg(const GcnLayerArgs& a, int Lmax, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    MMDFN_CUDA(cudaFuncSetAttribute(gcn_layer_kernel<FWD, KC, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, GLGeo<KC, NS>::SMEM));
    configured = true;
  }
  gcn_layer_kernel<FWD, KC, NS><<<dim3(a.B * 3, ceil_div(Lmax, GL_ROWS)), GL_THREADS, GLGeo<KC, NS>::SMEM, st>>>(a);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

template <bool FWD>
static int launch_layer(const GcnLayerArgs& a, int Lmax, cudaStream_t st) {
  if (g_gl_gen2 && g_gl_variant == 0 && gcn_layer2_eligible(Lmax)) return gcn_layer2_launch(FWD, a, Lmax, st);
  return g_gl_variant == 0 ? launch_layer_cfg<FWD, GL_KC0, GL_NS0>(a, Lmax, st) : launch_layer_cfg<FWD, GL_KC1, GL_NS1>(a, Lmax, st);
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int gcn_layer_fwd(int B, int N, int Lmax, const int* dia_off, const i64* blk_off, const float* adj_blk,
                  const float* adj_diag, const float* zin, const float* wimg, const float* r, i64 ldr, const float* q,
                  const unsigned char* mask, float scale, unsigned char* flags, float* out, i64 ldo, cudaStream_t st) {
  if (B <= 0 || N <= 0 || Lmax <= 0) return 0;
  if (!al16(zin) || !al16(wimg) || !al16(r) || !al16(out) || (q && !al16(q)) || (mask && (reinterpret_cast<uintptr_t>(mask) & 3)) ||
      (reinterpret_cast<uintptr_t>(flags) & 3) || (ldr & 3) || (ldo & 3))
    return MMDFN_EINVAL;
  GcnLayerArgs a{};
  a.B = B; a.N = N; a.dia_off = dia_off; a.blk_off = blk_off; a.adj_blk = adj_blk; a.adj_diag = adj_diag;
  a.zin = zin; a.ldz = GL_G; a.wimg = wimg; a.t_out = nullptr; a.ldt = 0;
  a.r = r; a.ldr = ldr; a.q = q; a.mask = mask; a.scale = mask ? scale : 1.f; a.flags = flags;
  a.add = nullptr; a.out = out; a.ldo = ldo; a.dbg = g_gl_dbg;
  return launch_layer<true>(a, Lmax, st);
}

int gcn_layer_bwd(int B, int N, int Lmax, const int* dia_off, const i64* blk_off, const float* adj_blk,
                  const float* adj_diag, const float* du, i64 ldu, const float* wimg_t, float* t_out, i64 ldt,
                  const float* add, float* out, cudaStream_t st) {
  if (B <= 0 || N <= 0 || Lmax <= 0) return 0;
  if (!al16(du) || !al16(wimg_t) || !al16(t_out) || !al16(out) || (add && !al16(add)) || (ldu & 3) || (ldt & 3)) return MMDFN_EINVAL;
  GcnLayerArgs a{};
  a.B = B; a.N = N; a.dia_off = dia_off; a.blk_off = blk_off; a.adj_blk = adj_blk; a.adj_diag = adj_diag;
  a.zin = du; a.ldz = ldu; a.wimg = wimg_t; a.t_out = t_out; a.ldt = ldt;
  a.r = nullptr; a.ldr = 0; a.q = nullptr; a.mask = nullptr; a.scale = 1.f; a.flags = nullptr;
  a.add = add; a.out = out; a.ldo = GL_G; a.dbg = g_gl_dbg;
  return launch_layer<false>(a, Lmax, st);
}

}  // namespace mmdfn

extern "C" long long mmdfn_gcn_layer_img_floats(void) { return mmdfn::gcn_layer_img_floats(); }

extern "C" int mmdfn_gcn_layer_set_variant(int v) {
  if (v < 0 || v > 2) return MMDFN_EINVAL;
  mmdfn::g_gl_variant = v == 1 ? 1 : 0;
  mmdfn::g_gl_gen2 = v == 0;
  return 0;
}

extern "C" int mmdfn_gcn_layer_prep(int K, const float* const* convW, double lamda, double alpha, float* mtop_all,
                                    float* mbot_all, float* img_f, float* img_b, void* stream) {
  if (K < 0) return MMDFN_EINVAL;
  if (K == 0) return 0;
  if (!convW || !mtop_all || !mbot_all || !img_f || !img_b) return MMDFN_ENULL;
  return mmdfn::gcn_layer_prep(K, convW, lamda, alpha, mtop_all, mbot_all, img_f, img_b, (cudaStream_t)stream);
}

extern "C" int mmdfn_gcn_layer_fwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off,
                                   const float* adj_blk, const float* adj_diag, const float* zin, const float* wimg,
                                   const float* r, long long ldr, const float* q, const unsigned char* mask,
                                   float mask_scale, unsigned char* flags, float* out, long long ldo, void* stream) {
  if (!dia_off || !blk_off || !adj_blk || !adj_diag || !zin || !wimg || !r || !flags || !out) return MMDFN_ENULL;
  return mmdfn::gcn_layer_fwd(B, N, Lmax, dia_off, (const mmdfn::i64*)blk_off, adj_blk, adj_diag, zin, wimg, r, ldr, q, mask,
                              mask_scale, flags, out, ldo, (cudaStream_t)stream);
}

extern "C" int mmdfn_gcn_layer_bwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off,
                                   const float* adj_blk, const float* adj_diag, const float* du, long long ldu,
                                   const float* wimg_t, float* t_out, long long ldt, const float* add, float* out,
                                   void* stream) {
  if (!dia_off || !blk_off || !adj_blk || !adj_diag || !du || !wimg_t || !t_out || !out) return MMDFN_ENULL;
  return mmdfn::gcn_layer_bwd(B, N, Lmax, dia_off, (const mmdfn::i64*)blk_off, adj_blk, adj_diag, du, ldu, wimg_t, t_out, ldt,
                              add, out, (cudaStream_t)stream);
}

// profiling aid: 256 x int64 device buffer receiving clock64() stamps of CTA (0,0): [0..62] phases ([63] = count),
// [64+g] issuer saw k-step g full, [128+g] converter saw its stage free, [192+g] converter reached the wait (nullptr = off)
extern "C" int mmdfn_gcn_layer_set_debug(long long* device_buf) {
  mmdfn::g_gl_dbg = device_buf;
  return 0;
}
