// k6 fused, second generation: the same contract as gcn_layer.cu (one launch = one whole GraphConvolution layer,
// code/model_GCN.py:176-189 inside the GCNII_lyc loop :461-472), rebuilt as a PERSISTENT warp-specialised kernel, one CTA
// per SM, for dialogues of up to 128 utterances:
//
//   phase A   T = A_hat[block] . Zin[block]   (+ the two cross-modal diagonal terms)
//   phase B   U = T . Mw
//   forward   z_out = dropout(relu(U + R)) (+ q), flags;      backward   d_in = U (+ add), T rows written out
//
// What changed against the first generation (round-2 phase stamps: 32 k cycles per tile; ~117 cycles per 128x112x8 MMA
// because the issuing lane rebuilt descriptors inside a lane-0 branch; every row operand fetched by latency-bound LDG
// batches; a tile's phases strictly one after the other):
//   * the A operand of both products lives in TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): the raw A_hat block
//     arrives by ONE TMA bulk copy, each thread reads its own row (conflict-free 128-bit reads), splits it into tf32
//     hi/lo and tcgen05.st's it; the phase-A result T goes TMEM -> registers (+ cross terms) -> hi/lo -> TMEM without
//     touching shared memory.  Only the B operand is fetched from shared memory by the tensor core.
//   * every row operand (cross-modal z rows, R, q, add, the keep mask; the out / T rows and the flag bytes on the way
//     back) moves by TMA bulk copies through two row buffers, issued by a producer warp: no LSU traffic, no exposed
//     latency in the compute warps; the pre-split weight image streams into the operand ring by bulk copies as well
//     (one 14.8 KB copy per 16-wide K chunk).
//   * the MMA warp runs its loop convergently with uniform operands and one elected lane issues: ~54 cycles per MMA
//     (= the tensor time, tools/umma_rate.py) instead of ~117.
//   * everything is pipelined per 16-column K chunk: phase B starts on T chunk c as soon as the hop has produced it
//     (after the accumulators were read out completely -- they are also phase B's destination), the NEXT tile's A_hat
//     chunk c is written as soon as phase B has consumed T chunk c, its z chunks are converted while phase B still
//     runs, and its phase A starts the moment the epilogue has drained the accumulators into registers.
//   * roles: 8 row warps (thread = tile row x column half: hop, epilogue), 4 A warps (A_hat -> TMEM), 4 z warps (z ->
//     transposed hi/lo ring stages: 4 x 4 blocks, four coalesced 128-bit loads, register transpose), one MMA warp, one
//     row-operand producer (loads AND stores), one weight-chunk producer.  The CTA walks tiles t = blockIdx.x,
//     += gridDim.x.
// Accuracy is unchanged (3-term tf32 split, separate main / correction accumulators in TMEM).
// Measured (B200, profiles/r02_gcn_layer2_*): 13.4 us per launch at the 32-dialogue bench shard (first generation 25.7),
// 56 us at 256 dialogues (96), 97 us at 512 (179); per-tile critical loop ~13.7 k cycles (hop + phase B 5.9 k, epilogue
// 4.4 k, out store -> next cross-row load through the shared row buffers 3.3 k).
#include "umma.cuh"
#include "internal.cuh"
#include "gcn_layer.cuh"

namespace mmdfn {
namespace {

constexpr int L2_G = 100, L2_BN = 112;
constexpr int L2_LBO = 128, L2_SBO = 528;
constexpr int L2_BPART = (L2_BN / 8) * L2_SBO;             // 7392 B: one B part (hi or lo) of a 16-wide K chunk
constexpr int L2_STAGE = 2 * L2_BPART;                     // 14784 B = one chunk of the pre-split weight image
constexpr int L2_WCHUNKS = 7;                              // K chunks of the 100-deep second product (13 k-steps of 8)
// Phase B consumes its K chunks in the order 0 3 1 4 2 5 6: the hop's two column halves (chunks 0-2 / 3-6) finish their
// chunks at the same pace, so alternating between them lets the tensor core follow the hop instead of waiting for one half.
__host__ __device__ constexpr int l2_bchunk(int i) { return i == 6 ? 6 : (i >> 1) + 3 * (i & 1); }
constexpr int L2_ROWW = 8, L2_CONVW = 8;
constexpr int L2_W_MMA = 16, L2_W_PA = 17, L2_W_PW = 18;
constexpr int L2_THREADS = 19 * 32;
constexpr uint32_t L2_TM_AH = 0, L2_TM_AL = 128, L2_TM_DM = 256, L2_TM_DC = 384, L2_TMEM = 512;
constexpr int L2_HDR = 512;
constexpr int L2_MAXNS = 6;
constexpr int L2_SMEM_MAX = 232448;

enum {
  BAR_FULL = 0, BAR_FREE = L2_MAXNS, BAR_A_RDY = 2 * L2_MAXNS, BAR_T_RDY = BAR_A_RDY + 8, BAR_T_USED = BAR_T_RDY + 7,
  BAR_ARAW_FULL = BAR_T_USED + 7, BAR_ARAW_FREE, BAR_MMA_A, BAR_MMA_B, BAR_D_FREE, BAR_T_DRAINED, BAR_X_FULL, BAR_Y_FULL, BAR_XY_FREE, BAR_COUNT
};
static_assert(8 * BAR_COUNT + 4 <= L2_HDR, "barrier header");

struct L2Lay { int araw, x, y, mask, ring, total; };
__host__ __device__ inline int l2_up(int v) { return (v + 127) & ~127; }
__host__ __device__ inline L2Lay l2_layout(int lmax, int ns) {
  L2Lay l;
  l.araw = L2_HDR;
  l.x = l.araw + l2_up(4 * lmax * lmax + 16);
  l.y = l.x + l2_up(400 * lmax);
  l.mask = l.y + l2_up(400 * lmax);
  l.ring = l.mask + l2_up(100 * lmax);
  l.total = l.ring + ns * L2_STAGE;
  return l;
}

// ---- PTX helpers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(umma::smem_u32(dst)), "l"(src), "r"(bytes), "r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(umma::smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void row_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 8 consecutive fp32 columns of this thread's lane from two accumulators, one wait
__device__ __forceinline__ void tmem_ld8x2(uint32_t ta, uint32_t tb, float (&v)[8], float (&w)[8]) {
  uint32_t r[8], q[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "r"(tb) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; i++) { v[i] = __uint_as_float(r[i]); w[i] = __uint_as_float(q[i]); }
}
__device__ __forceinline__ void tmem_st8(uint32_t ta, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(ta), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// split 8 values and store them as the A operand (hi / lo) at column `col` of this thread's lane
__device__ __forceinline__ void split_st8(uint32_t tlane, uint32_t col, const float (&v)[8]) {
  float hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; e++) gl_split(v[e], hi[e], lo[e]);
  tmem_st8(tlane + L2_TM_AH + col, hi);
  tmem_st8(tlane + L2_TM_AL + col, lo);
}

struct Tile { int b, m, off, L; };
// keep mask / flag bytes of a tile (100 L contiguous bytes at byte offset 100 row0) can move by TMA bulk copies when that
// range is 16-byte aligned and a multiple of 16 bytes; otherwise the row warps copy them with plain 32-bit accesses
__device__ __forceinline__ bool mask_by_tma(const GcnLayerArgs& p, i64 row0, int L) {
  return ((L & 3) == 0) && ((row0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.flags) & 15) == 0) &&
         (p.mask == nullptr || (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0);
}
__device__ __forceinline__ Tile tile_of(const GcnLayerArgs& p, int t) {
  Tile x;
  x.b = t / 3;
  x.m = t - 3 * x.b;
  x.off = p.dia_off[x.b];
  x.L = p.dia_off[x.b + 1] - x.off;
  return x;
}

template <bool FWD>
__global__ void __launch_bounds__(L2_THREADS, 1) gcn_layer2_kernel(GcnLayerArgs p, int ntiles, int lmax, int ns) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * BAR_COUNT);
  const L2Lay ly = l2_layout(lmax, ns);
  float* araw = reinterpret_cast<float*>(smem + ly.araw);
  float* X = reinterpret_cast<float*>(smem + ly.x);
  float* Y = reinterpret_cast<float*>(smem + ly.y);
  uint32_t* mask_s = reinterpret_cast<uint32_t*>(smem + ly.mask);
  uint8_t* ring = smem + ly.ring;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == L2_W_MMA) umma::tmem_alloc(tmem_slot, L2_TMEM);
  if (tid == 0) {
    for (int s = 0; s < L2_MAXNS; s++) {
      umma::mbar_init(&bars[BAR_FULL + s], 4);                // a z chunk: the 4 warps of one converter group; a weight chunk: 3 + expect_tx
      umma::mbar_init(&bars[BAR_FREE + s], 1);
    }
    for (int c = 0; c < 8; c++) umma::mbar_init(&bars[BAR_A_RDY + c], 4);
    for (int c = 0; c < 7; c++) {
      umma::mbar_init(&bars[BAR_T_RDY + c], 4);
      umma::mbar_init(&bars[BAR_T_USED + c], 1);
    }
    umma::mbar_init(&bars[BAR_ARAW_FULL], 1);
    umma::mbar_init(&bars[BAR_ARAW_FREE], 4);
    umma::mbar_init(&bars[BAR_MMA_A], 1);
    umma::mbar_init(&bars[BAR_MMA_B], 1);
    umma::mbar_init(&bars[BAR_D_FREE], L2_ROWW);
    umma::mbar_init(&bars[BAR_T_DRAINED], L2_ROWW);
    umma::mbar_init(&bars[BAR_X_FULL], 1);
    umma::mbar_init(&bars[BAR_Y_FULL], 1);
    umma::mbar_init(&bars[BAR_XY_FREE], 1);
    umma::fence_barrier_init();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // profiling aid (mmdfn_gcn_layer_set_debug): clock64 stamps of CTA 0, 32 slots per tile for its first 8 tiles
  const bool dbg_cta = p.dbg != nullptr && blockIdx.x == 0;
#define L2_STAMP(cond, slot) do { if (dbg_cta && (cond) && it < 8) p.dbg[32 * it + (slot)] = clock64(); } while (0)

  if (warp < L2_ROWW) {
    // ============================== row warps: thread = tile row; cross-term hop + fused epilogue ==============================
    // Column split: half 0 owns columns 0..47 (K chunks 0..2 of phase B), half 1 columns 48..103 (chunks 3..6), so that
    // every 16-column chunk of T is complete -- and phase B may consume it -- as soon as ONE half has passed it.
    const int q4 = warp & 3, half = warp >> 2;
    const int row = q4 * 32 + lane;
    const int cbeg = half ? 48 : 0, nblk = half ? 7 : 6;
    const uint32_t tlane = tmem + ((uint32_t)(q4 * 32) << 16);
    uint32_t xyc = 0;
    int it = 0;
    // geometry and the two cross-modal diagonal entries of this thread's row are fetched one tile ahead (during the
    // previous tile's epilogue), so that the hop can start the moment phase A completes
    auto diag_of = [&](const Tile& x, float& d1, float& d2) {
      const int o1 = (x.m == 0) ? 1 : 0, o2 = (x.m == 2) ? 1 : 2;
      d1 = d2 = 0.f;
      if (row < x.L) {
        d1 = __ldg(p.adj_diag + (i64)(min(x.m, o1) + max(x.m, o1) - 1) * p.N + x.off + row);
        d2 = __ldg(p.adj_diag + (i64)(min(x.m, o2) + max(x.m, o2) - 1) * p.N + x.off + row);
      }
    };
    Tile tn = tile_of(p, min((int)blockIdx.x, ntiles - 1));
    float e1n, e2n;
    diag_of(tn, e1n, e2n);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
      const Tile tl = tn;
      const float e1 = e1n, e2 = e2n;
      const int L = tl.L, m = tl.m;
      const i64 row0 = (i64)m * p.N + tl.off;
      const bool rv = row < L;
      const int nw = 25 * L;                                   // 32-bit words of the tile's keep mask / flags
      const bool mtma = mask_by_tma(p, row0, L);
      if (FWD && p.mask && !mtma) {
        // 25 L <= 3200 words over 256 threads: at most 13 per thread, all loads issued before the first store
        const uint32_t* mg = reinterpret_cast<const uint32_t*>(p.mask + row0 * L2_G);
        uint32_t mv[13];
#pragma unroll
        for (int i = 0; i < 13; i++) mv[i] = (tid + 256 * i < nw) ? __ldg(mg + tid + 256 * i) : 0u;
#pragma unroll
        for (int i = 0; i < 13; i++)
          if (tid + 256 * i < nw) mask_s[tid + 256 * i] = mv[i];
      }
      float* xr = X + row * L2_G;
      const float* yr = Y + row * L2_G;
      L2_STAMP(tid == 0, 0);
      // ---------------- hop: T = D (+ cross terms) -> hi/lo -> A operand of phase B, chunk by chunk ----------------
      umma::mbar_wait(&bars[BAR_MMA_A], (uint32_t)(it & 1));
      umma::tc_fence_after_sync();
      L2_STAMP(tid == 0, 1);
      umma::mbar_wait(&bars[BAR_X_FULL], xyc & 1);
      umma::mbar_wait(&bars[BAR_Y_FULL], xyc & 1);
      xyc++;
      L2_STAMP(tid == 0, 2);
      // the accumulators are also the destination of phase B: read this thread's part of T out completely first
      float tv[56];
#pragma unroll
      for (int blk = 0; blk < 7; blk++) {
        if (blk < nblk) {
          const int c0 = cbeg + 8 * blk;
          float v[8], w[8];
          tmem_ld8x2(tlane + L2_TM_DM + c0, tlane + L2_TM_DC + c0, v, w);
#pragma unroll
          for (int e = 0; e < 8; e++) tv[8 * blk + e] = v[e] + w[e];
        }
      }
      umma::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_T_DRAINED]);
#pragma unroll
      for (int blk = 0; blk < 7; blk++) {
        if (blk < nblk) {
          const int c0 = cbeg + 8 * blk;
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; e++) v[e] = tv[8 * blk + e];
          if (rv) {
#pragma unroll
            for (int h4 = 0; h4 < 2; h4++) {
              const int c = c0 + 4 * h4;
              if (c < L2_G) {
                const float4 a = *reinterpret_cast<const float4*>(xr + c);
                const float4 bq = *reinterpret_cast<const float4*>(yr + c);
                v[4 * h4 + 0] += e1 * a.x + e2 * bq.x;
                v[4 * h4 + 1] += e1 * a.y + e2 * bq.y;
                v[4 * h4 + 2] += e1 * a.z + e2 * bq.z;
                v[4 * h4 + 3] += e1 * a.w + e2 * bq.w;
                if (!FWD) *reinterpret_cast<float4*>(xr + c) = make_float4(v[4 * h4], v[4 * h4 + 1], v[4 * h4 + 2], v[4 * h4 + 3]);
              }
            }
          }
          split_st8(tlane, (uint32_t)c0, v);
          if ((blk & 1) || blk == nblk - 1) {                  // a 16-column chunk of T is complete in this warp's lanes
            tmem_wait_st();
            umma::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_T_RDY + (c0 >> 4)]);
          }
        }
      }
      if (!FWD) umma::fence_proxy_async_smem();
      L2_STAMP(tid == 0, 3);
      row_bar_sync();                                           // every row warp is done with the cross rows
      if (tid == 0) mbar_arrive(&bars[BAR_XY_FREE]);          // the row-operand producer stores the T rows (backward) and reloads
      // ---------------- epilogue ----------------
      L2_STAMP(tid == 0, 4);
      umma::mbar_wait(&bars[BAR_MMA_B], (uint32_t)(it & 1));
      umma::tc_fence_after_sync();
      L2_STAMP(tid == 0, 5);
      float u[56];
#pragma unroll
      for (int blk = 0; blk < 7; blk++) {
        const int c0 = cbeg + 8 * blk;
        if (blk < nblk && c0 < L2_G) {
          float v[8], w[8];
          tmem_ld8x2(tlane + L2_TM_DM + c0, tlane + L2_TM_DC + c0, v, w);
#pragma unroll
          for (int e = 0; e < 8; e++) u[8 * blk + e] = v[e] + w[e];
        } else {
#pragma unroll
          for (int e = 0; e < 8; e++) u[8 * blk + e] = 0.f;
        }
      }
      umma::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_D_FREE]);            // accumulators drained: the next tile's phase A may start
      L2_STAMP(tid == 0, 6);
      if (t + (int)gridDim.x < ntiles) {
        tn = tile_of(p, t + gridDim.x);
        diag_of(tn, e1n, e2n);
      }
      umma::mbar_wait(&bars[BAR_X_FULL], xyc & 1);
      umma::mbar_wait(&bars[BAR_Y_FULL], xyc & 1);
      xyc++;
      L2_STAMP(tid == 0, 7);
      if (rv) {
#pragma unroll
        for (int blk = 0; blk < 7; blk++) {
#pragma unroll
          for (int h4 = 0; h4 < 2; h4++) {
            const int c = cbeg + 8 * blk + 4 * h4;
            if (blk < nblk && c < L2_G) {
              float o[4] = {u[8 * blk + 4 * h4], u[8 * blk + 4 * h4 + 1], u[8 * blk + 4 * h4 + 2], u[8 * blk + 4 * h4 + 3]};
              if (FWD) {
                const float4 r4 = *reinterpret_cast<const float4*>(xr + c);
                const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
                const uint32_t mk = p.mask ? mask_s[row * 25 + (c >> 2)] : 0x01010101u;       // keep bytes are 0 / 1
                uint32_t fl = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                  // relu, then keep * scale as ONE factor; the result is +0 or positive, so "kept and positive" is "bits != 0"
                  const float ks = (float)((mk >> (8 * j)) & 0xFFu) * p.scale;
                  o[j] = fmaxf(o[j] + rr[j], 0.f) * ks;
                  fl |= min(__float_as_uint(o[j]), 1u) << (8 * j);
                }
                if (p.q) {
                  const float4 q4v = *reinterpret_cast<const float4*>(yr + c);
                  o[0] += q4v.x; o[1] += q4v.y; o[2] += q4v.z; o[3] += q4v.w;
                }
                mask_s[row * 25 + (c >> 2)] = fl;
              } else if (p.add) {
                const float4 a4 = *reinterpret_cast<const float4*>(yr + c);
                o[0] += a4.x; o[1] += a4.y; o[2] += a4.z; o[3] += a4.w;
              }
              *reinterpret_cast<float4*>(xr + c) = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      }
      L2_STAMP(tid == 0, 8);
      umma::fence_proxy_async_smem();
      row_bar_sync();
      L2_STAMP(tid == 0, 9);
      if (FWD && !mtma) {
        uint32_t* fg = reinterpret_cast<uint32_t*>(p.flags + row0 * L2_G);
        for (int w = tid; w < nw; w += 256) fg[w] = mask_s[w];
      }
      if (tid == 0) mbar_arrive(&bars[BAR_XY_FREE]);          // the row-operand producer stores the out rows
      L2_STAMP(tid == 0, 10);
    }
  } else if (warp < L2_ROWW + 4) {
    // ============================== A warps (4): A_hat block -> TMEM ==============================
    // thread = tile row; one 16-column K chunk = two 8-column blocks; chunks go two at a time so that ONE tcgen05.wait::st
    // covers four stores.  Chunk c of the NEXT tile may be written as soon as this tile's phase B has consumed T chunk c.
    const int q4 = warp & 3, atid = tid - 32 * L2_ROWW;
    const int row = q4 * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(q4 * 32) << 16);
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
      const Tile tl = tile_of(p, t);
      const int L = tl.L, m = tl.m;
      const int nchA = (L + 15) >> 4, k8n = (L + 7) >> 3;
      const float* A = p.adj_blk + p.blk_off[tl.b] + (i64)m * L * L;
      const int leadw = (int)((reinterpret_cast<uintptr_t>(A) & 15) >> 2);
      const float* arow = araw + leadw + row * L;
      const bool vec = (leadw == 0) && ((L & 3) == 0);
      L2_STAMP(atid == 0, 12);
      umma::mbar_wait(&bars[BAR_ARAW_FULL], (uint32_t)(it & 1));
      L2_STAMP(atid == 0, 13);
#pragma unroll 1
      for (int c2 = 0; c2 < nchA; c2 += 2) {
        const int cend = min(c2 + 2, nchA);
        if (it > 0) {
          for (int c = c2; c < cend; c++)
            if (c < L2_WCHUNKS) umma::mbar_wait(&bars[BAR_T_USED + c], (uint32_t)((it - 1) & 1));
          umma::tc_fence_after_sync();
        }
        if (c2 == 0) L2_STAMP(atid == 0, 14);
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          const int j8 = 2 * c2 + jj;
          if (j8 < k8n) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) v[e] = 0.f;
            if (row < L) {
              const int k = 8 * j8;
              if (vec) {
                if (k < L) {
                  const float4 a = *reinterpret_cast<const float4*>(arow + k);
                  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                }
                if (k + 4 < L) {
                  const float4 a = *reinterpret_cast<const float4*>(arow + k + 4);
                  v[4] = a.x; v[5] = a.y; v[6] = a.z; v[7] = a.w;
                }
              } else {
#pragma unroll
                for (int e = 0; e < 8; e++)
                  if (k + e < L) v[e] = arow[k + e];
              }
            }
            split_st8(tlane, (uint32_t)(8 * j8), v);
          }
        }
        tmem_wait_st();
        umma::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0)
          for (int c = c2; c < cend; c++) mbar_arrive(&bars[BAR_A_RDY + c]);
      }
      __syncwarp();
      if (lane == 0) {
        for (int c = nchA; c < 8; c++) mbar_arrive(&bars[BAR_A_RDY + c]);      // every barrier completes once per tile (parity = tile index)
        mbar_arrive(&bars[BAR_ARAW_FREE]);
      }
      L2_STAMP(atid == 0, 15);
    }
  } else if (warp < L2_ROWW + L2_CONVW) {
    // ============================== z warps (4): z block -> B operand of phase A ==============================
    // B(n = feature, k = block row) = z[k][n].  A thread loads a 4 x 4 block (rows k..k+3, columns 4 cg..4 cg+3) with four
    // coalesced 128-bit loads (two chunks ahead), transposes it in registers and stores four hi and four lo pieces.
    const int gt = tid - 32 * (L2_ROWW + 4);
    const int zkq = gt / 28, zcg = gt - 28 * zkq;              // gt < 112: block (column group zcg, k-quad zkq)
    const bool zact = gt < 112, zcol = zact && zcg < 25;
    const int zo = (zcg >> 1) * L2_SBO + zkq * L2_LBO + (zcg & 1) * 64;
    uint32_t gc = 0;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
      const Tile tl = tile_of(p, t);
      const int L = tl.L, m = tl.m;
      const int nchA = (L + 15) >> 4;
      const float* zp = p.zin + ((i64)m * p.N + tl.off + 4 * zkq) * p.ldz + 4 * zcg;
      float4 zr[2][4];
      auto load_z = [&](int c, float4 (&d)[4]) {
        const float* q = zp + (i64)(16 * c) * p.ldz;
        const int k0 = 16 * c + 4 * zkq;
#pragma unroll
        for (int j = 0; j < 4; j++)
          d[j] = (zcol && k0 + j < L) ? __ldg(reinterpret_cast<const float4*>(q + (i64)j * p.ldz)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      load_z(0, zr[0]);
      if (1 < nchA) load_z(1, zr[1]);
#pragma unroll 1
      for (int c2 = 0; c2 < nchA; c2 += 2) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int c = c2 + u;
          if (c < nchA) {
            float4 h[4], l[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const float a0 = i == 0 ? zr[u][0].x : i == 1 ? zr[u][0].y : i == 2 ? zr[u][0].z : zr[u][0].w;
              const float a1 = i == 0 ? zr[u][1].x : i == 1 ? zr[u][1].y : i == 2 ? zr[u][1].z : zr[u][1].w;
              const float a2 = i == 0 ? zr[u][2].x : i == 1 ? zr[u][2].y : i == 2 ? zr[u][2].z : zr[u][2].w;
              const float a3 = i == 0 ? zr[u][3].x : i == 1 ? zr[u][3].y : i == 2 ? zr[u][3].z : zr[u][3].w;
              gl_split(a0, h[i].x, l[i].x);
              gl_split(a1, h[i].y, l[i].y);
              gl_split(a2, h[i].z, l[i].z);
              gl_split(a3, h[i].w, l[i].w);
            }
            if (c + 2 < nchA) load_z(c + 2, zr[u]);
            const uint32_t g = gc + (uint32_t)c;
            const int s = (int)(g % (uint32_t)ns);
            if (g >= (uint32_t)ns) umma::mbar_wait(&bars[BAR_FREE + s], ((g / (uint32_t)ns) - 1u) & 1u);
            if (zact) {
              uint8_t* st = ring + s * L2_STAGE + zo;
#pragma unroll
              for (int i = 0; i < 4; i++) {
                *reinterpret_cast<float4*>(st + 16 * i) = h[i];
                *reinterpret_cast<float4*>(st + L2_BPART + 16 * i) = l[i];
              }
            }
            umma::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_FULL + s]);
          }
        }
      }
      L2_STAMP(gt == 0, 16);
      gc += (uint32_t)(nchA + L2_WCHUNKS);
    }
  } else if (warp == L2_W_MMA) {
    // ============================== MMA issuer ==============================
    // The issuing thread is the bottleneck of a 128 x 112 x 8 MMA stream (tools/umma_rate.py: 54 cycles per MMA from a
    // bare loop = the tensor time; ~117 when the instruction sits in a lane-0 branch and descriptors are rebuilt between
    // MMAs).  The whole warp therefore runs this loop convergently on uniform values, one elected lane issues, and a
    // k-step is three MMAs separated by 32-bit adds only.
    {
      constexpr uint32_t IDESC = umma::idesc_tf32(128, L2_BN);
      const uint64_t d0 = umma::smem_desc(umma::smem_u32(ring), L2_LBO, L2_SBO);
      const uint32_t d0_hi = (uint32_t)(d0 >> 32), d0_lo = (uint32_t)d0;
      auto kstep = [&](uint32_t dlo, uint32_t kcol, uint32_t acc) {
        const uint64_t b_hi = ((uint64_t)d0_hi << 32) | dlo;
        const uint64_t b_lo = ((uint64_t)d0_hi << 32) | (dlo + (L2_BPART >> 4));
        umma::mma_tf32_ta_elect(tmem + L2_TM_DM, tmem + L2_TM_AH + kcol, b_hi, IDESC, acc);
        umma::mma_tf32_ta_elect(tmem + L2_TM_DC, tmem + L2_TM_AL + kcol, b_hi, IDESC, acc);
        umma::mma_tf32_ta_elect(tmem + L2_TM_DC, tmem + L2_TM_AH + kcol, b_lo, IDESC, 1u);
      };
      uint32_t gc = 0;
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
        const Tile tl = tile_of(p, t);
        const int nchA = (tl.L + 15) >> 4, ksA = (tl.L + 7) >> 3;
        const uint32_t tpar = (uint32_t)(it & 1);
        if (it > 0) umma::mbar_wait(&bars[BAR_D_FREE], (uint32_t)((it - 1) & 1));
        L2_STAMP(lane == 0, 19);
        for (int ph = 0; ph < 2; ph++) {
          const int nch = ph == 0 ? nchA : L2_WCHUNKS, ks = ph == 0 ? ksA : 13;
          uint32_t s = gc % (uint32_t)ns, par = (gc / (uint32_t)ns) & 1u;
          if (ph == 1) umma::mbar_wait(&bars[BAR_T_DRAINED], tpar);                     // the hop has read T out of the accumulators
          for (int i = 0; i < nch; i++) {
            const int c = ph == 0 ? i : l2_bchunk(i);
            umma::mbar_wait(&bars[(ph == 0 ? BAR_A_RDY : BAR_T_RDY) + c], tpar);       // A operand chunk c is in TMEM
            umma::mbar_wait(&bars[BAR_FULL + s], par);
            umma::tc_fence_after_sync();
            if (i == 0) L2_STAMP(lane == 0, ph == 0 ? 20 : 23);
            if (i == nch - 1) L2_STAMP(lane == 0, ph == 0 ? 21 : 24);
            const uint32_t dlo = d0_lo + s * (L2_STAGE >> 4);
            kstep(dlo, (uint32_t)(16 * c), i > 0 ? 1u : 0u);
            if (ks - 2 * c > 1) kstep(dlo + ((2 * L2_LBO) >> 4), (uint32_t)(16 * c + 8), 1u);
            umma::mma_commit_elect(&bars[BAR_FREE + s]);
            if (ph == 1) umma::mma_commit_elect(&bars[BAR_T_USED + c]);
            if (++s == (uint32_t)ns) { s = 0; par ^= 1u; }
          }
          gc += (uint32_t)nch;
          umma::mma_commit_elect(&bars[ph == 0 ? BAR_MMA_A : BAR_MMA_B]);
        }
      }
    }
  } else if (warp == L2_W_PA) {
    // ============================== TMA producer: raw A_hat block and the row operands ==============================
    auto issue_araw = [&](int t) {
      const Tile tl = tile_of(p, t);
      if (lane == 0) {
        const float* A = p.adj_blk + p.blk_off[tl.b] + (i64)tl.m * tl.L * tl.L;
        const uintptr_t a = reinterpret_cast<uintptr_t>(A);
        const uint32_t lead = (uint32_t)(a & 15);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a - lead);
        const uint32_t total = lead + 4u * (uint32_t)tl.L * (uint32_t)tl.L;
        const uint32_t mainb = total & ~15u;
        // the last (< 16) bytes of an unaligned block end: plain loads, made visible by the arrive below
        for (uint32_t o = mainb > lead ? mainb : lead; o < total; o += 4)
          *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(araw) + o) = __ldg(reinterpret_cast<const float*>(src + o));
        if (mainb) {
          mbar_expect_tx(&bars[BAR_ARAW_FULL], mainb);
          bulk_g2s(araw, src, mainb, &bars[BAR_ARAW_FULL]);
        } else {
          mbar_arrive(&bars[BAR_ARAW_FULL]);
        }
      }
    };
    // L rows of 100 floats (row stride ld) -> dst, completing `bar`; src == nullptr: nothing to load, just complete
    auto issue_rows = [&](float* dst, const float* src, i64 ld, int L, uint64_t* bar) {
      if (src == nullptr) {
        if (lane == 0) mbar_arrive(bar);
        return;
      }
      if (lane == 0) mbar_expect_tx(bar, 400u * (uint32_t)L);
      __syncwarp();
      if (ld == L2_G) {
        if (lane == 0) bulk_g2s(dst, src, 400u * (uint32_t)L, bar);
      } else {
        for (int r = lane; r < L; r += 32) bulk_g2s(dst + r * L2_G, src + (i64)r * ld, 400u, bar);
      }
    };
    // L rows of the X buffer -> dst (row stride ld); returns when the copy engine has READ the buffer
    auto store_rows = [&](float* dst, i64 ld, int L) {
      if (ld == L2_G) {
        if (lane == 0) bulk_s2g(dst, X, 400u * (uint32_t)L);
      } else {
        for (int r = lane; r < L; r += 32) bulk_s2g(dst + (i64)r * ld, X + r * L2_G, 400u);
      }
      bulk_commit_wait_read();
      __syncwarp();
    };
    uint32_t xyw = 0;
    int it = 0;
    i64 prev_row0 = 0;
    int prev_L = 0;
    if ((int)blockIdx.x < ntiles) issue_araw(blockIdx.x);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
      const Tile tl = tile_of(p, t);
      const int L = tl.L, m = tl.m;
      const i64 row0 = (i64)m * p.N + tl.off;
      const int o1 = (m == 0) ? 1 : 0, o2 = (m == 2) ? 1 : 2;
      // the previous tile's result rows leave from X; then the cross-modal rows of this tile's hop arrive
      if (it > 0) {
        umma::mbar_wait(&bars[BAR_XY_FREE], xyw & 1);
        xyw++;
        if (FWD && lane == 0 && mask_by_tma(p, prev_row0, prev_L)) bulk_s2g(p.flags + prev_row0 * L2_G, mask_s, 100u * (uint32_t)prev_L);
        store_rows(p.out + prev_row0 * p.ldo, p.ldo, prev_L);
      }
      issue_rows(X, p.zin + ((i64)o1 * p.N + tl.off) * p.ldz, p.ldz, L, &bars[BAR_X_FULL]);
      issue_rows(Y, p.zin + ((i64)o2 * p.N + tl.off) * p.ldz, p.ldz, L, &bars[BAR_Y_FULL]);
      // the next tile's A_hat block, as soon as the converters have read this one
      if (t + (int)gridDim.x < ntiles) {
        umma::mbar_wait(&bars[BAR_ARAW_FREE], (uint32_t)(it & 1));
        issue_araw(t + gridDim.x);
      }
      // after the hop: (backward) the T rows leave from X; then the operands of the epilogue arrive
      umma::mbar_wait(&bars[BAR_XY_FREE], xyw & 1);
      xyw++;
      if (FWD) {
        if (p.mask && mask_by_tma(p, row0, L)) {               // the keep bytes ride on the Y barrier (its own expect_tx arrival)
          if (lane == 0) {
            asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(umma::smem_u32(&bars[BAR_Y_FULL])), "r"(100u * (uint32_t)L) : "memory");
            bulk_g2s(mask_s, p.mask + row0 * L2_G, 100u * (uint32_t)L, &bars[BAR_Y_FULL]);
          }
          __syncwarp();
        }
        issue_rows(X, p.r + row0 * p.ldr, p.ldr, L, &bars[BAR_X_FULL]);
        issue_rows(Y, p.q ? p.q + row0 * L2_G : nullptr, L2_G, L, &bars[BAR_Y_FULL]);
      } else {
        store_rows(p.t_out + row0 * p.ldt, p.ldt, L);
        issue_rows(X, nullptr, 0, L, &bars[BAR_X_FULL]);
        issue_rows(Y, p.add ? p.add + row0 * L2_G : nullptr, L2_G, L, &bars[BAR_Y_FULL]);
      }
      prev_row0 = row0;
      prev_L = L;
    }
    if (it > 0) {
      umma::mbar_wait(&bars[BAR_XY_FREE], xyw & 1);
      if (FWD && lane == 0 && mask_by_tma(p, prev_row0, prev_L)) bulk_s2g(p.flags + prev_row0 * L2_G, mask_s, 100u * (uint32_t)prev_L);
      store_rows(p.out + prev_row0 * p.ldo, p.ldo, prev_L);
    }
    __syncwarp();
  } else {
    // ============================== TMA producer: weight-image chunks of phase B ==============================
    // Walks EVERY ring use in order (also the z chunks the converters fill) and observes each stage's "free" completion,
    // so a parity wait can never alias an older phase; a weight chunk is issued as soon as its stage has been released.
    if (lane == 0) {
      uint32_t gc = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const Tile tl = tile_of(p, t);
        const int nchA = (tl.L + 15) >> 4;
        for (int c = 0; c < nchA + L2_WCHUNKS; c++) {
          const int s = (int)(gc % (uint32_t)ns);
          if (gc >= (uint32_t)ns) umma::mbar_wait(&bars[BAR_FREE + s], ((gc / (uint32_t)ns) - 1u) & 1u);
          if (c >= nchA) {
            mbar_arrive_cnt(&bars[BAR_FULL + s], 3);
            mbar_expect_tx(&bars[BAR_FULL + s], L2_STAGE);
            bulk_g2s(ring + s * L2_STAGE, p.wimg + (i64)l2_bchunk(c - nchA) * (L2_STAGE / 4), L2_STAGE, &bars[BAR_FULL + s]);
          }
          gc++;
        }
      }
    }
    __syncwarp();
  }
#undef L2_STAMP
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == L2_W_MMA) umma::tmem_dealloc(tmem, L2_TMEM);
}

int pick_stages(int Lmax) {
  int ns = L2_MAXNS;
  while (ns >= 2 && l2_layout(Lmax, ns).total > L2_SMEM_MAX) ns--;
  return ns;
}

}  // namespace

bool gcn_layer2_eligible(int Lmax) { return Lmax >= 1 && Lmax <= 128 && pick_stages(Lmax) >= 2; }

int gcn_layer2_launch(bool fwd, const GcnLayerArgs& a, int Lmax, cudaStream_t st) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    MMDFN_CUDA(cudaGetDevice(&dev));
    MMDFN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MMDFN_CUDA(cudaFuncSetAttribute(gcn_layer2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L2_SMEM_MAX));
    MMDFN_CUDA(cudaFuncSetAttribute(gcn_layer2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L2_SMEM_MAX));
  }
  const int ns = pick_stages(Lmax);
  if (ns < 2) return MMDFN_EINVAL;
  const int ntiles = 3 * a.B;
  const int grid = ntiles < sms ? ntiles : sms;
  const int smem = l2_layout(Lmax, ns).total;
  if (fwd) gcn_layer2_kernel<true><<<grid, L2_THREADS, smem, st>>>(a, ntiles, Lmax, ns);
  else gcn_layer2_kernel<false><<<grid, L2_THREADS, smem, st>>>(a, ntiles, Lmax, ns);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmdfn
