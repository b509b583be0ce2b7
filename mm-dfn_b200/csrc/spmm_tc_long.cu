// The default aggregate kernel for G = 100 (any dialogue length).  Validated on B200 in round 2
// (profiles/r02_spmm_variant_long_kernel.log, tests/test_gpu_parity.py::test_aggregate_kernels_*).
//
// k6 on the tensor cores for dialogues of ANY length: the generalisation of spmm_tc.cu (same 3xTF32 tcgen05 scheme, same
// operand layouts, same epilogue) from "one CTA = one whole (dialogue, modality) block of <= 128 utterances" to
// "one CTA = one 128-row tile of a block, contraction streamed over the block's L columns":
//   * grid (3 B, ceil(Lmax / 128)); CTAs whose row tile starts beyond their dialogue's length exit at once;
//   * the z block no longer fits in shared memory (L = 500: 200 KB), so its 16-row chunks travel through a 4-slot raw ring
//     filled by TMA bulk copies.  Slot reuse needs no extra barrier: the issuer thread refills slot c % 4 right after it
//     has seen the "stage full" barrier of chunk c, i.e. after every converter finished reading that slot;
//   * rows of the last chunk beyond L are masked in the transposed read (the ring slot holds stale data there);
//   * the output tile is staged in the operand ring (free once the last MMA has completed) and leaves through one bulk
//     store of the tile's rows.
// At the measured FFMA fallback (9-13 % of the HBM roof at L = 200..500) this is the fix for
// BASELINE config 5 (500-utterance dialogues), whose aggregate is compute-bound on FFMA (AI 31 flop/B).
#include "umma.cuh"
#include "internal.cuh"

namespace mmdfn {

constexpr int SL_G = 100, SL_BN = 112, SL_KC = 16, SL_ROWS = 128;
constexpr int SL_LBO = 128, SL_SBO = 512;
constexpr int SL_A_PART = 16 * SL_SBO;                     // 8192 B
constexpr int SL_B_PART = (SL_BN / 8) * SL_SBO;            // 7168 B
constexpr int SL_STAGE = 2 * (SL_A_PART + SL_B_PART);      // 30720 B
constexpr int SL_ZR = 4;                                   // raw z ring slots (chunks in flight)
constexpr int SL_ZSLOT = SL_KC * SL_G * 4;                 // 6400 B
constexpr int SL_SMEM = SL_ZR * SL_ZSLOT + 2 * SL_STAGE;   // 87040 B -> two CTAs per SM
constexpr int SL_CONV = 256, SL_THREADS = SL_CONV + 32;
constexpr int SL_CORR = 128, SL_TMEM = 256;
static_assert(2 * SL_STAGE >= SL_ROWS * SL_G * 4, "the output tile is staged in the operand ring");

struct SpmmLongArgs {
  int B, N;
  const int* dia_off;
  const i64* blk_off;
  const float* adj_blk;
  const float* adj_diag;
  const float* x;
  float* y;
};

__device__ __forceinline__ int sl_pair_of(int m, int n) { return m + n - 1; }

__device__ __forceinline__ void sl_split(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

__global__ void __launch_bounds__(SL_THREADS, 2) adj_spmm_tc_long_kernel(SpmmLongArgs p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_free[2];
  __shared__ __align__(8) uint64_t bar_full[2];
  __shared__ __align__(8) uint64_t bar_z[SL_ZR];
  __shared__ uint32_t tmem_base_s;
  __shared__ float dsm[2][SL_ROWS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / 3, m = blockIdx.x % 3;
  const int off = p.dia_off[b], L = p.dia_off[b + 1] - off;
  const int r0 = blockIdx.y * SL_ROWS;
  if (r0 >= L) return;                                       // whole CTA: nothing allocated yet
  const int nrows = min(SL_ROWS, L - r0);
  const float* A = p.adj_blk + p.blk_off[b] + (i64)m * L * L + (i64)r0 * L;   // first row of this tile
  const float* Z = p.x + ((i64)m * p.N + off) * SL_G;
  const int nchunks = (L + SL_KC - 1) / SL_KC;

  float* zring = reinterpret_cast<float*>(smem);
  uint8_t* stages = smem + SL_ZR * SL_ZSLOT;
  auto issue_z = [&](int c) {                                // chunk c -> ring slot c % SL_ZR (single thread)
    const int j0 = c * SL_KC;
    const uint32_t bytes = (uint32_t)(min(L, j0 + SL_KC) - j0) * SL_G * 4;
    const uint32_t bar = umma::smem_u32(&bar_z[c % SL_ZR]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(umma::smem_u32(zring + (c % SL_ZR) * (SL_KC * SL_G))), "l"(Z + (i64)j0 * SL_G), "r"(bytes), "r"(bar)
                 : "memory");
  };
  if (warp == 8) umma::tmem_alloc(&tmem_base_s, SL_TMEM);
  if (tid == 0) {
    for (int s = 0; s < 2; s++) {
      umma::mbar_init(&bar_free[s], 1);
      umma::mbar_init(&bar_full[s], SL_CONV / 32);
    }
    for (int s = 0; s < SL_ZR; s++) umma::mbar_init(&bar_z[s], 1);
    umma::fence_barrier_init();
    issue_z(0);
  }
  constexpr uint32_t IDESC = umma::idesc_tf32(128, SL_BN);

  // A pieces of this thread: row group warp + 8 i (i = 0, 1), row lane & 7, k-quad lane >> 3
  const int r_in = lane & 7, kq_a = lane >> 3;
  const bool a_vec = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  auto load_a = [&](int c, float4 (&va)[2]) {
    const int k = c * SL_KC + 4 * kq_a;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int row = (warp + 8 * i) * 8 + r_in;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < nrows && k < L) {
        const float* q = A + (i64)row * L + k;
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
  if (tid < 2 * SL_ROWS) {
    const int r = tid & (SL_ROWS - 1), which = tid >> 7, o = which ? o2 : o1;
    dsm[which][r] = (r < nrows) ? __ldg(p.adj_diag + (i64)sl_pair_of(min(m, o), max(m, o)) * p.N + off + r0 + r) : 0.f;
  }
  float4 va[SL_ZR][2];                                       // A pieces of the next SL_ZR chunks (same depth as the z ring)
  if (warp < 8) {
#pragma unroll
    for (int u = 0; u < SL_ZR; u++)
      if (u < nchunks) load_a(u, va[u]);
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == 8) {
    // ===== MMA issuer + z producer =====
    if (lane == 0)
      for (int c = 1; c < SL_ZR && c < nchunks; c++) issue_z(c);
    __syncwarp();
    {
      const uint64_t d0 = umma::smem_desc(umma::smem_u32(stages), SL_LBO, SL_SBO);
      const uint32_t dhi = (uint32_t)(d0 >> 32), dlo = (uint32_t)d0;
      for (int c = 0; c < nchunks; c++) {
        const int s = c & 1;
        umma::mbar_wait(&bar_full[s], (uint32_t)((c >> 1) & 1));
        umma::tc_fence_after_sync();
        // every converter has finished reading ring slot c % SL_ZR (it arrived on bar_full after its reads): refill it
        if (lane == 0 && c + SL_ZR < nchunks) issue_z(c + SL_ZR);
        __syncwarp();
        const uint32_t o = dlo + (uint32_t)s * (SL_STAGE >> 4);
        const int kleft = L - c * SL_KC;
        const int ksteps = kleft >= SL_KC ? SL_KC / 8 : (kleft + 7) / 8;
        for (int j = 0; j < ksteps; j++) {
          const uint32_t oj = o + (uint32_t)j * ((2 * SL_LBO) >> 4);
          umma::kstep3_elect(tmem, tmem + SL_CORR, dhi, oj, oj + (SL_A_PART >> 4), oj + ((2 * SL_A_PART) >> 4),
                             oj + ((2 * SL_A_PART + SL_B_PART) >> 4), IDESC, (c > 0 || j > 0) ? 1u : 0u);
        }
        umma::mma_commit_elect(&bar_free[s]);
      }
    }
    __syncwarp();
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tmem_dealloc(tmem, SL_TMEM);
    return;
  }

  // ===== converters (warps 0-7) =====
  const int c_b = lane + 32 * (warp & 3);                    // feature column of this thread's B pieces
  for (int c0 = 0; c0 < nchunks; c0 += SL_ZR) {
#pragma unroll
    for (int u = 0; u < SL_ZR; u++) {                        // u = ring slot = c % SL_ZR (c0 is a multiple of SL_ZR)
      const int c = c0 + u;
      if (c < nchunks) {
        const int s = u & 1;                                 // = c & 1
        umma::mbar_wait(&bar_z[u], (uint32_t)((c / SL_ZR) & 1));
        if (c >= 2) umma::mbar_wait(&bar_free[s], (uint32_t)(((c >> 1) - 1) & 1));
        uint8_t* st = stages + s * SL_STAGE;
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const int o = (warp + 8 * i) * SL_SBO + kq_a * SL_LBO + r_in * 16;
          const float4 v = va[u][i];
          float4 h, l;
          sl_split(v.x, h.x, l.x);
          sl_split(v.y, h.y, l.y);
          sl_split(v.z, h.z, l.z);
          sl_split(v.w, h.w, l.w);
          *reinterpret_cast<float4*>(st + o) = h;
          *reinterpret_cast<float4*>(st + SL_A_PART + o) = l;
        }
        if (c + SL_ZR < nchunks) load_a(c + SL_ZR, va[u]);
        if (c_b < SL_BN) {
          const float* slot = zring + u * (SL_KC * SL_G);
          const int jleft = L - c * SL_KC;                   // valid rows in this chunk (>= 1)
#pragma unroll
          for (int i = 0; i < 2; i++) {
            const int kq = (warp >> 2) + 2 * i;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c_b < SL_G) {                                // columns 100..111 pad N; rows beyond L are masked
              const float* q = slot + (4 * kq) * SL_G + c_b;
              if (4 * kq < jleft) v.x = q[0];
              if (4 * kq + 1 < jleft) v.y = q[SL_G];
              if (4 * kq + 2 < jleft) v.z = q[2 * SL_G];
              if (4 * kq + 3 < jleft) v.w = q[3 * SL_G];
            }
            const int o = (c_b >> 3) * SL_SBO + kq * SL_LBO + (c_b & 7) * 16;
            float4 h, l;
            sl_split(v.x, h.x, l.x);
            sl_split(v.y, h.y, l.y);
            sl_split(v.z, h.z, l.z);
            sl_split(v.w, h.w, l.w);
            *reinterpret_cast<float4*>(st + 2 * SL_A_PART + o) = h;
            *reinterpret_cast<float4*>(st + 2 * SL_A_PART + SL_B_PART + o) = l;
          }
        }
        umma::warp_arrive_full(&bar_full[s]);
      }
    }
  }

  // ---- epilogue (as in spmm_tc.cu, for the tile's rows r0 .. r0 + nrows - 1) ----
  const float4* x1 = reinterpret_cast<const float4*>(p.x + ((i64)o1 * p.N + off + r0) * SL_G);
  const float4* x2 = reinterpret_cast<const float4*>(p.x + ((i64)o2 * p.N + off + r0) * SL_G);
  const int total = nrows * (SL_G / 4);
  constexpr int EB = 5;
  float4 a1[EB], a2[EB];
#pragma unroll
  for (int u = 0; u < EB; u++) {
    const int i = tid + u * SL_CONV;
    if (i < total) {
      a1[u] = __ldg(x1 + i);
      a2[u] = __ldg(x2 + i);
    }
  }
  {
    const int last = nchunks - 1;
    umma::mbar_wait(&bar_free[last & 1], (uint32_t)((last >> 1) & 1));
  }
  umma::tc_fence_after_sync();
  float* tile_f = reinterpret_cast<float*>(stages);          // the operand ring is free: every MMA has completed
  {
    const int r = (warp & 3) * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int cb_begin = (warp < 4) ? 0 : 64, cb_end = (warp < 4) ? 64 : SL_BN;
#pragma unroll 1
    for (int cb = cb_begin; cb < cb_end; cb += 16) {
      if (cb >= SL_G) break;
      float v[16], w[16];
      umma::tmem_ld16x2(taddr + cb, taddr + SL_CORR + cb, v, w);
#pragma unroll
      for (int q4 = 0; q4 < 16; q4 += 4) {
        if (cb + q4 < SL_G)
          *reinterpret_cast<float4*>(tile_f + r * SL_G + cb + q4) =
              make_float4(v[q4] + w[q4], v[q4 + 1] + w[q4 + 1], v[q4 + 2] + w[q4 + 2], v[q4 + 3] + w[q4 + 3]);
      }
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  {
    float4* tile = reinterpret_cast<float4*>(stages);
#pragma unroll 1
    for (int base = tid; base < total; base += EB * SL_CONV) {
      if (base != tid) {
#pragma unroll
        for (int u = 0; u < EB; u++) {
          const int i = base + u * SL_CONV;
          if (i < total) {
            a1[u] = __ldg(x1 + i);
            a2[u] = __ldg(x2 + i);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < EB; u++) {
        const int i = base + u * SL_CONV;
        if (i < total) {
          const int r = i / (SL_G / 4);
          const float e1 = dsm[0][r], e2 = dsm[1][r];
          const float4 t = tile[i];
          tile[i] = make_float4(t.x + e1 * a1[u].x + e2 * a2[u].x, t.y + e1 * a1[u].y + e2 * a2[u].y,
                                t.z + e1 * a1[u].z + e2 * a2[u].z, t.w + e1 * a1[u].w + e2 * a2[u].w);
        }
      }
    }
  }
  umma::fence_proxy_async_smem();
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (tid == 0) {
    float* yb = p.y + ((i64)m * p.N + off + r0) * SL_G;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(yb), "r"(umma::smem_u32(stages)), "r"((uint32_t)nrows * SL_G * 4) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  umma::tc_fence_before_sync();
  __syncthreads();
}

// tensor-core aggregate for G == 100 and any dialogue length (16-byte aligned x / y)
int adj_spmm_tc_long(int B, int N, int Lmax, const int* dia_off, const i64* blk_off, const float* adj_blk,
                     const float* adj_diag, const float* x, float* y, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    MMDFN_CUDA(cudaFuncSetAttribute(adj_spmm_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM));
    configured = true;
  }
  SpmmLongArgs a{B, N, dia_off, blk_off, adj_blk, adj_diag, x, y};
  adj_spmm_tc_long_kernel<<<dim3(B * 3, ceil_div(Lmax, SL_ROWS)), SL_THREADS, SL_SMEM, st>>>(a);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmdfn
