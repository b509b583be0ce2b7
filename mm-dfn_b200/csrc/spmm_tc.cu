// k6 on the tensor cores: hi = A_hat z for G = 100 and blocks of up to 128 utterances, one CTA per (dialogue, modality)
// block, tcgen05.mma kind::tf32 with the 3-term split (fp32-level accuracy), accumulators in TMEM.
//
//   D (128 x 112) = A (128 x K, rows = block rows r, K = j)  *  B (112 x K, rows = feature column c, K = j)
//
// A = the L x L block of A_hat (rows are K-contiguous: 16-byte pieces go straight from global memory through registers
// into the UMMA K-major layout).  B = z^T: z's rows are the contraction index, so the block of z is first copied raw
// into shared memory with TMA bulk copies (one mbarrier per 16-row chunk) and then read back *transposed* with
// conflict-free 32-bit loads (lanes = consecutive columns) to form K-major pieces -- no conflicted scalar stores.
// Warp-specialised like umma_gemm.cu: 8 converter warps fill a 2-stage operand ring (K chunks of 16), warp 8 issues
// hi*hi into TMEM columns 0-111 and lo*hi, hi*lo into columns 128-239; tcgen05.commit releases stages.
// Epilogue: the cross-modal diagonal terms d[r] * z_n[r, :] need the other two modalities' z rows; their first batch
// of 128-bit loads is issued BEFORE the wait for the last MMA, the accumulators go TMEM -> registers (thread = row)
// -> a row-major tile in the free raw-z region, a coalesced pass adds the cross terms in shared memory, and ONE bulk
// (TMA) store writes the block's L x 100 contiguous output rows.  (A cluster-of-3 variant that read the peers' raw z
// through DSMEM was measured 2.3x slower: DSMEM delivers ~20 B/clk/SM, L2 ~64.)  112.6 KB of shared memory and 256
// TMEM columns per CTA: two CTAs per SM, so one block's loads and epilogue overlap the other's MMAs.
#include "umma.cuh"
#include "internal.cuh"

namespace mmdfn {

constexpr int ST_G = 100, ST_BN = 112, ST_KC = 16, ST_LMAX = 128, ST_NCH = ST_LMAX / ST_KC;
constexpr int ST_LBO = 128, ST_SBO = 512;                  // 4 core matrices (K = 16) per 8-row group, unpadded
constexpr int ST_A_PART = 16 * ST_SBO;                     // 128 rows:  8192 B
constexpr int ST_B_PART = (ST_BN / 8) * ST_SBO;            // 112 rows:  7168 B
constexpr int ST_STAGE = 2 * (ST_A_PART + ST_B_PART);      // hi + lo of both operands: 30720 B
constexpr int ST_RAWZ = ST_LMAX * ST_G * 4;                // 51200 B (raw z block, later the output tile)
constexpr int ST_SMEM = ST_RAWZ + 2 * ST_STAGE;            // 112640 B -> two CTAs per SM
constexpr int ST_CONV = 256, ST_THREADS = ST_CONV + 32;
constexpr int ST_CORR = 128, ST_TMEM = 256;                // two CTAs x 256 columns = the SM's 512
constexpr int ST_ADEPTH = 4;                               // chunks of A held in registers ahead of their conversion

struct SpmmTcArgs {
  int B, N;
  const int* dia_off;
  const i64* blk_off;
  const float* adj_blk;
  const float* adj_diag;
  const float* x;
  float* y;
  long long* dbg;
};

__device__ __forceinline__ int st_pair_of(int m, int n) { return m + n - 1; }

// x = hi + lo, hi = x rounded to tf32 (add half an ulp of the 10-bit mantissa to the magnitude, clear the low 13 bits:
// round-to-nearest, ties away -- the same value cvt.rna.tf32.f32 yields for finite x, in two integer instructions)
__device__ __forceinline__ void st_split(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

__global__ void __launch_bounds__(ST_THREADS, 2) adj_spmm_tc_kernel(SpmmTcArgs p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_free[2];
  __shared__ __align__(8) uint64_t bar_full[2];
  __shared__ __align__(8) uint64_t bar_z[ST_NCH];             // one per 16-row chunk of the raw z copy (TMA bulk, single use)
  __shared__ uint32_t tmem_base_s;
  __shared__ float dsm[2][ST_LMAX];                           // cross-modal diagonal entries of this block's rows
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / 3, m = blockIdx.x % 3;
  const int off = p.dia_off[b], L = p.dia_off[b + 1] - off;
  const float* A = p.adj_blk + p.blk_off[b] + (i64)m * L * L;
  const float* Z = p.x + ((i64)m * p.N + off) * ST_G;
  const int nchunks = (L + ST_KC - 1) / ST_KC;
  const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
  int dbg_n = 0;
#define ST_STAMP() do { if (dbg_on && dbg_n < 60) p.dbg[dbg_n++] = clock64(); } while (0)
  ST_STAMP();                                                 // [0] entry

  float* rawz = reinterpret_cast<float*>(smem);
  // raw copy of the z block: one bulk (TMA) copy per 16-row chunk, each completing its own mbarrier, so chunk 0 can be
  // converted as soon as its 6.4 KB landed.  The block's rows are one contiguous, 16-byte aligned run.
  auto issue_z = [&](int c) {
    const int j0 = c * ST_KC;
    const uint32_t bytes = (uint32_t)(min(L, j0 + ST_KC) - j0) * ST_G * 4;
    const uint32_t bar = umma::smem_u32(&bar_z[c]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(umma::smem_u32(rawz + j0 * ST_G)), "l"(Z + (i64)j0 * ST_G), "r"(bytes), "r"(bar) : "memory");
  };
  if (warp == 8) umma::tmem_alloc(&tmem_base_s, ST_TMEM);
  if (tid == 0) {
    for (int s = 0; s < 2; s++) {
      umma::mbar_init(&bar_free[s], 1);
      umma::mbar_init(&bar_full[s], ST_CONV / 32);
    }
    for (int c = 0; c < ST_NCH; c++) umma::mbar_init(&bar_z[c], 1);
    umma::fence_barrier_init();
    if (nchunks > 0) issue_z(0);                              // chunk 0 now, the rest by the (idle) issuer thread below
  }
  uint8_t* stages = smem + ST_RAWZ;
  constexpr uint32_t IDESC = umma::idesc_tf32(128, ST_BN);

  // ---- operand traffic is put in flight before the set-up barrier ----
  // A pieces of this thread: row group warp + 8 i (i = 0, 1), row lane & 7, k-quad lane >> 3
  const int r_in = lane & 7, kq_a = lane >> 3;
  const bool a_vec = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  auto load_a = [&](int c, float4 (&va)[2]) {
    const int k = c * ST_KC + 4 * kq_a;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int row = (warp + 8 * i) * 8 + r_in;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < L && k < L) {
        const float* q = A + (i64)row * L + k;
        if (a_vec) {                                          // L % 4 == 0 -> k + 3 < L
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
  // cross-modal partners of modality m; their diagonal entries go to shared memory now (read by the epilogue)
  const int o1 = (m == 0) ? 1 : 0, o2 = (m == 2) ? 1 : 2;
  if (tid < 2 * ST_LMAX) {
    const int r = tid & (ST_LMAX - 1), which = tid >> 7, o = which ? o2 : o1;
    dsm[which][r] = (r < L) ? __ldg(p.adj_diag + (i64)st_pair_of(min(m, o), max(m, o)) * p.N + off + r) : 0.f;
  }
  float4 va[ST_ADEPTH][2];
  if (warp < 8) {
    // rows L .. 16*nchunks-1 of the raw tile are read (as zeros) by the last chunk's transposed loads
    for (int i = L * ST_G + tid; i < nchunks * ST_KC * ST_G; i += ST_CONV) rawz[i] = 0.f;
#pragma unroll
    for (int c = 0; c < ST_ADEPTH; c++)
      if (c < nchunks) load_a(c, va[c]);
  }
  ST_STAMP();                                                 // [1] loads issued
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  ST_STAMP();                                                 // [2] set-up done

  if (warp == 8) {
    // ===== MMA issuer =====
    if (lane == 0)
      for (int c = 1; c < nchunks; c++) issue_z(c);           // each issue costs ~250 cycles: off the converters' path
    __syncwarp();
    {
      const uint64_t d0 = umma::smem_desc(umma::smem_u32(stages), ST_LBO, ST_SBO);
      const uint32_t dhi = (uint32_t)(d0 >> 32), dlo = (uint32_t)d0;
      for (int c = 0; c < nchunks; c++) {
        const int s = c & 1;
        umma::mbar_wait(&bar_full[s], (uint32_t)((c >> 1) & 1));
        umma::tc_fence_after_sync();
        const uint32_t o = dlo + (uint32_t)s * (ST_STAGE >> 4);
        const int kleft = L - c * ST_KC;
        const int ksteps = kleft >= ST_KC ? ST_KC / 8 : (kleft + 7) / 8;
        for (int j = 0; j < ksteps; j++) {
          const uint32_t oj = o + (uint32_t)j * ((2 * ST_LBO) >> 4);
          umma::kstep3_elect(tmem, tmem + ST_CORR, dhi, oj, oj + (ST_A_PART >> 4), oj + ((2 * ST_A_PART) >> 4),
                             oj + ((2 * ST_A_PART + ST_B_PART) >> 4), IDESC, (c > 0 || j > 0) ? 1u : 0u);
        }
        umma::mma_commit_elect(&bar_free[s]);
      }
    }
    __syncwarp();
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tmem_dealloc(tmem, ST_TMEM);
    return;
  }

  // ===== converters (warps 0-7) =====
  // B pieces of this thread: feature column c_b = lane + 32 (warp & 3), k-quads (warp >> 2) + 2 i
  const int c_b = lane + 32 * (warp & 3);
#pragma unroll
  for (int c = 0; c < ST_NCH; c++) {
    if (c < nchunks) {
      const int s = c & 1;
      umma::mbar_wait(&bar_z[c], 0u);                         // chunk c's rows of z landed (async-proxy writes visible)
      ST_STAMP();                                             // z chunk landed
      if (c >= 2) umma::mbar_wait(&bar_free[s], (uint32_t)(((c >> 1) - 1) & 1));
      ST_STAMP();                                             // stage free
      uint8_t* st = stages + s * ST_STAGE;
      // ---- A: registers -> hi/lo
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const int o = (warp + 8 * i) * ST_SBO + kq_a * ST_LBO + r_in * 16;
        const float4 v = va[c % ST_ADEPTH][i];
        float4 h, l;
        st_split(v.x, h.x, l.x);
        st_split(v.y, h.y, l.y);
        st_split(v.z, h.z, l.z);
        st_split(v.w, h.w, l.w);
        *reinterpret_cast<float4*>(st + o) = h;
        *reinterpret_cast<float4*>(st + ST_A_PART + o) = l;
      }
      if (c + ST_ADEPTH < nchunks) load_a(c + ST_ADEPTH, va[c % ST_ADEPTH]);
      // ---- B: transposed read of the raw z rows (conflict-free: lanes = consecutive columns)
      if (c_b < ST_BN) {
        const int j0 = c * ST_KC;
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const int kq = (warp >> 2) + 2 * i;
          const int j = j0 + 4 * kq;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c_b < ST_G) {                                   // rows beyond L were zero-filled, columns 100..111 pad N
            const float* q = rawz + j * ST_G + c_b;
            v = make_float4(q[0], q[ST_G], q[2 * ST_G], q[3 * ST_G]);
          }
          const int o = (c_b >> 3) * ST_SBO + kq * ST_LBO + (c_b & 7) * 16;
          float4 h, l;
          st_split(v.x, h.x, l.x);
          st_split(v.y, h.y, l.y);
          st_split(v.z, h.z, l.z);
          st_split(v.w, h.w, l.w);
          *reinterpret_cast<float4*>(st + 2 * ST_A_PART + o) = h;
          *reinterpret_cast<float4*>(st + 2 * ST_A_PART + ST_B_PART + o) = l;
        }
      }
      umma::warp_arrive_full(&bar_full[s]);
      ST_STAMP();                                             // converted
    }
  }
  // ---- epilogue ----
  // cross-modal rows: float4 i of the block (row i / 25) for i = tid + u * 256; the first EB of them are requested
  // before the wait for the last MMA so that their latency hides behind it
  const float4* x1 = reinterpret_cast<const float4*>(p.x + ((i64)o1 * p.N + off) * ST_G);
  const float4* x2 = reinterpret_cast<const float4*>(p.x + ((i64)o2 * p.N + off) * ST_G);
  const int total = L * (ST_G / 4);
  constexpr int EB = 5;
  float4 a1[EB], a2[EB];
#pragma unroll
  for (int u = 0; u < EB; u++) {
    const int i = tid + u * ST_CONV;
    if (i < total) {
      a1[u] = __ldg(x1 + i);
      a2[u] = __ldg(x2 + i);
    }
  }
  if (nchunks > 0) {
    const int last = nchunks - 1;
    umma::mbar_wait(&bar_free[last & 1], (uint32_t)((last >> 1) & 1));
  }
  umma::tc_fence_after_sync();
  ST_STAMP();                                                 // all MMAs done

  // TMEM -> the (now free) raw z region as a row-major [r][100] tile; thread = block row; warps 0-3 take columns
  // 0..63, warps 4-7 columns 64..99.  Every converter passed the last full barrier before the final commit could
  // fire, so nobody still reads rawz.
  {
    const int r = (warp & 3) * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int cb_begin = (warp < 4) ? 0 : 64, cb_end = (warp < 4) ? 64 : ST_BN;
#pragma unroll 1
    for (int cb = cb_begin; cb < cb_end; cb += 16) {
      if (cb >= ST_G) break;
      float v[16], w[16];
      umma::tmem_ld16x2(taddr + cb, taddr + ST_CORR + cb, v, w);
#pragma unroll
      for (int q4 = 0; q4 < 16; q4 += 4) {
        if (cb + q4 < ST_G)                                   // G = 100: the last chunk stops after one quad
          *reinterpret_cast<float4*>(rawz + r * ST_G + cb + q4) =
              make_float4(v[q4] + w[q4], v[q4 + 1] + w[q4 + 1], v[q4 + 2] + w[q4 + 2], v[q4 + 3] + w[q4 + 3]);
      }
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  ST_STAMP();                                                 // tile in shared memory
  {
    float4* tile = reinterpret_cast<float4*>(rawz);
#pragma unroll 1
    for (int base = tid; base < total; base += EB * ST_CONV) {
      if (base != tid) {                                      // later batches: loads issued here
#pragma unroll
        for (int u = 0; u < EB; u++) {
          const int i = base + u * ST_CONV;
          if (i < total) {
            a1[u] = __ldg(x1 + i);
            a2[u] = __ldg(x2 + i);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < EB; u++) {
        const int i = base + u * ST_CONV;
        if (i < total) {
          const int r = i / (ST_G / 4);
          const float e1 = dsm[0][r], e2 = dsm[1][r];
          const float4 t = tile[i];
          tile[i] = make_float4(t.x + e1 * a1[u].x + e2 * a2[u].x, t.y + e1 * a1[u].y + e2 * a2[u].y,
                                t.z + e1 * a1[u].z + e2 * a2[u].z, t.w + e1 * a1[u].w + e2 * a2[u].w);
        }
      }
    }
  }
  umma::fence_proxy_async_smem();                             // tile writes -> visible to the bulk-copy engine
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (tid == 0 && L > 0) {
    // the block's output rows are one contiguous, 16-byte aligned run of L * 400 bytes
    float* yb = p.y + ((i64)m * p.N + off) * ST_G;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(yb), "r"(umma::smem_u32(rawz)), "r"((uint32_t)L * ST_G * 4) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  ST_STAMP();                                                 // stored
  if (dbg_on) p.dbg[63] = dbg_n;
#undef ST_STAMP
  umma::tc_fence_before_sync();
  __syncthreads();
}

static long long* g_spmm_dbg = nullptr;

// tensor-core aggregate for G == 100, Lmax <= 128, 16-byte aligned x / y
int adj_spmm_tc(int B, int N, const int* dia_off, const i64* blk_off, const float* adj_blk, const float* adj_diag,
                const float* x, float* y, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    MMDFN_CUDA(cudaFuncSetAttribute(adj_spmm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
    configured = true;
  }
  SpmmTcArgs a{B, N, dia_off, blk_off, adj_blk, adj_diag, x, y, g_spmm_dbg};
  adj_spmm_tc_kernel<<<B * 3, ST_THREADS, ST_SMEM, st>>>(a);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmdfn

extern "C" int mmdfn_adj_spmm_set_debug(long long* device_buf) {
  mmdfn::g_spmm_dbg = device_buf;
  return 0;
}
