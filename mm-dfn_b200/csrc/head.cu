// k9: classifier head (dropout -> ReLU -> Linear(900,C) -> log_softmax; code/model.py:1328-1337)
// and FocalLoss (code/loss.py:14-34), forward + backward; plus the counter-based dropout
// mask generator and the fused flat-buffer Adam(+L2) step used by the data-parallel trainer
// (optim.Adam(lr, weight_decay=l2), code/run_train_erc.py:512).
#include "internal.cuh"
#include "../../include/mmdfn_b200.h"

namespace mmdfn {

constexpr int HF = 300;     // per-modality feature width [x200 | g100]
constexpr int MAXC = 16;

// in-place row-wise log_softmax over C (thread per row)
__global__ void log_softmax_kernel(int N, int C, const float* logits, float* out) {   // in place allowed
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* x = logits + (i64)n * C;
  float mx = x[0];
  for (int c = 1; c < C; c++) mx = fmaxf(mx, x[c]);
  float s = 0.f;
  for (int c = 0; c < C; c++) s += expf(x[c] - mx);
  const float lse = mx + logf(s);
  for (int c = 0; c < C; c++) out[(i64)n * C + c] = x[c] - lse;
}

// dlogits = dlp - exp(lp) * sum_c dlp
__global__ void log_softmax_bwd_kernel(int N, int C, const float* __restrict__ lp, const float* __restrict__ dlp,
                                       float* __restrict__ dlogits) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int c = 0; c < C; c++) s += dlp[(i64)n * C + c];
  for (int c = 0; c < C; c++) dlogits[(i64)n * C + c] = dlp[(i64)n * C + c] - expf(lp[(i64)n * C + c]) * s;
}

// ---- fused head (round 2): one launch for the forward, two for the backward ---------------------------------------
// The head is 3N x 300 inputs against a (C, 900) weight with C <= 16: ~11 MFLOP over 11.5 MB -- an HBM-bound row
// reduction, not a GEMM.  Forward: one warp per utterance walks its three 300-wide segments with 128-bit loads, applies
// keep * scale and ReLU, stores R (saved for the backward), accumulates the C dot products against the weight held in
// shared memory, reduces them over the warp and writes log_softmax.  Backward kernel 1 (warp per utterance): dlogits
// (stored for kernel 2) and dF = (dlogits Wc) * keep * scale * [R > 0].  Backward kernel 2: dWc / dbc as column-parallel
// sums over row chunks (a thread owns one of the 900 (+1 bias) columns), combined with atomics into zeroed targets.
// weight (C, 900) -> shared memory: 128-bit loads, all of a thread's loads in flight before the first store (a plain
// load-store loop paid one global-memory latency per iteration: 21 of them, ~6 us of the kernel's 24)
__device__ __forceinline__ void head_load_w(float* wsm, const float* __restrict__ Wc, int C) {
  const int n4 = C * 225;                                     // float4 pieces
  if ((reinterpret_cast<uintptr_t>(Wc) & 15) == 0) {
    for (int i0 = threadIdx.x; i0 < n4; i0 += 8 * blockDim.x) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (i0 + u * blockDim.x < n4) v[u] = __ldg(reinterpret_cast<const float4*>(Wc) + i0 + u * blockDim.x);
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (i0 + u * blockDim.x < n4) reinterpret_cast<float4*>(wsm)[i0 + u * blockDim.x] = v[u];
    }
  } else {
    for (int i = threadIdx.x; i < C * 900; i += blockDim.x) wsm[i] = Wc[i];
  }
}

template <int CT>
__global__ void __launch_bounds__(256) head_fused_fwd_kernel(int N, int C, const float* __restrict__ F,
                                                              const unsigned char* __restrict__ mask, float scale, int relu,
                                                              const float* __restrict__ Wc, const float* __restrict__ bc,
                                                              float* __restrict__ R, float* __restrict__ lp) {
  extern __shared__ __align__(16) float wsm[];                // (C, 900)
  head_load_w(wsm, Wc, C);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    float acc[CT];
#pragma unroll
    for (int c = 0; c < CT; c++) acc[c] = 0.f;
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const i64 rb = ((i64)m * N + n) * HF;
#pragma unroll
      for (int it = 0; it < 3; it++) {
        const int g = lane + 32 * it;                          // float4 group of the 300-wide segment (75 groups)
        if (g < 75) {
          float4 v = ldg_stream4(F + rb + 4 * g);
          if (mask) {
            const uint32_t mk = *reinterpret_cast<const uint32_t*>(mask + (i64)n * 900 + m * HF + 4 * g);
            v.x = (mk & 0xFFu) ? v.x * scale : 0.f;
            v.y = (mk & 0xFF00u) ? v.y * scale : 0.f;
            v.z = (mk & 0xFF0000u) ? v.z * scale : 0.f;
            v.w = (mk & 0xFF000000u) ? v.w * scale : 0.f;
          }
          if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          *reinterpret_cast<float4*>(R + rb + 4 * g) = v;
#pragma unroll
          for (int c = 0; c < CT; c++) {
            if (c < C) {
              const float4 w = *reinterpret_cast<const float4*>(wsm + c * 900 + m * HF + 4 * g);
              acc[c] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[c]))));
            }
          }
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CT; c++) {
      acc[c] = warp_sum(acc[c]) + (c < C ? bc[c] : 0.f);
      if (c < C) mx = fmaxf(mx, acc[c]);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CT; c++)
      if (c < C) s += expf(acc[c] - mx);
    const float lse = mx + logf(s);
#pragma unroll
    for (int c = 0; c < CT; c++)
      if (c == lane && c < C) lp[(i64)n * C + c] = acc[c] - lse;
  }
}

template <int CT>
__global__ void __launch_bounds__(256) head_fused_bwd_kernel(int N, int C, const unsigned char* __restrict__ mask, float scale,
                                                              int relu, const float* __restrict__ Wc, const float* __restrict__ R,
                                                              const float* __restrict__ lp, const float* __restrict__ dlp,
                                                              float* __restrict__ dF, float* __restrict__ dlogits) {
  extern __shared__ __align__(16) float wsm[];
  head_load_w(wsm, Wc, C);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    float dl[CT];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CT; c++) {
      dl[c] = c < C ? dlp[(i64)n * C + c] : 0.f;
      sum += dl[c];
    }
#pragma unroll
    for (int c = 0; c < CT; c++) {
      if (c < C) dl[c] -= expf(lp[(i64)n * C + c]) * sum;
      if (c == lane && c < C) dlogits[(i64)n * C + c] = dl[c];
    }
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const i64 rb = ((i64)m * N + n) * HF;
#pragma unroll
      for (int it = 0; it < 3; it++) {
        const int g = lane + 32 * it;
        if (g < 75) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int c = 0; c < CT; c++) {
            if (c < C) {
              const float4 w = *reinterpret_cast<const float4*>(wsm + c * 900 + m * HF + 4 * g);
              a.x = fmaf(dl[c], w.x, a.x); a.y = fmaf(dl[c], w.y, a.y); a.z = fmaf(dl[c], w.z, a.z); a.w = fmaf(dl[c], w.w, a.w);
            }
          }
          if (mask) {
            const uint32_t mk = *reinterpret_cast<const uint32_t*>(mask + (i64)n * 900 + m * HF + 4 * g);
            a.x = (mk & 0xFFu) ? a.x * scale : 0.f;
            a.y = (mk & 0xFF00u) ? a.y * scale : 0.f;
            a.z = (mk & 0xFF0000u) ? a.z * scale : 0.f;
            a.w = (mk & 0xFF000000u) ? a.w * scale : 0.f;
          }
          if (relu) {
            const float4 r = ldg_stream4(R + rb + 4 * g);
            if (!(r.x > 0.f)) a.x = 0.f;
            if (!(r.y > 0.f)) a.y = 0.f;
            if (!(r.z > 0.f)) a.z = 0.f;
            if (!(r.w > 0.f)) a.w = 0.f;
          }
          *reinterpret_cast<float4*>(dF + rb + 4 * g) = a;
        }
      }
    }
  }
}

// dWc[c][j] += sum_n dlogits[n][c] R[n][j]  (j < 900; column 900 = the bias gradient), rows chunked over blockIdx.y
constexpr int HW_ROWS = 64;
template <int CT>
__global__ void __launch_bounds__(128) head_wgrad_kernel(int N, int C, const float* __restrict__ R, const float* __restrict__ dlogits,
                                                          float* __restrict__ dWc, float* __restrict__ dbc) {
  __shared__ float dls[HW_ROWS * CT];
  const int n0 = blockIdx.y * HW_ROWS, rows = min(HW_ROWS, N - n0);
  for (int i = threadIdx.x; i < rows * C; i += blockDim.x) dls[(i / C) * CT + (i % C)] = dlogits[(i64)n0 * C + i];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > 900) return;
  float acc[CT];
#pragma unroll
  for (int c = 0; c < CT; c++) acc[c] = 0.f;
  if (j < 900) {
    const int m = j / HF, k = j - m * HF;
    const float* rp = R + ((i64)m * N + n0) * HF + k;
#pragma unroll 4
    for (int r = 0; r < rows; r++) {
      const float v = rp[(i64)r * HF];
#pragma unroll
      for (int c = 0; c < CT; c++) acc[c] = fmaf(dls[r * CT + c], v, acc[c]);
    }
#pragma unroll
    for (int c = 0; c < CT; c++)
      if (c < C) atomicAdd(dWc + c * 900 + j, acc[c]);
  } else {
    for (int r = 0; r < rows; r++)
#pragma unroll
      for (int c = 0; c < CT; c++) acc[c] += dls[r * CT + c];
#pragma unroll
    for (int c = 0; c < CT; c++)
      if (c < C) atomicAdd(dbc + c, acc[c]);
  }
}

// loss = (mean|sum)_n -(1-pt)^gamma * alpha[y] * lp[n,y],  pt = exp(lp[n,y]) (no gradient through pt)
__device__ __forceinline__ float focal_term(int n, int C, const float* __restrict__ lp, const long long* __restrict__ target,
                                            const float* __restrict__ alpha, float gamma, float norm) {
  const long long y = target[n];
  const float l = lp[(i64)n * C + y];
  const float a = alpha ? alpha[y] : 1.f;
  const float w = gamma == 0.f ? 1.f : powf(1.f - expf(l), gamma);
  return -w * a * l * norm;
}

__device__ __forceinline__ float focal_block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
  return warp_sum(v);                                      // valid in warp 0
}

__global__ void focal_fwd_kernel(int N, int C, const float* __restrict__ lp, const long long* __restrict__ target,
                                 const float* __restrict__ alpha, float gamma, float norm, float* __restrict__ loss) {
  __shared__ float red[32];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const float v = focal_block_sum(n < N ? focal_term(n, C, lp, target, alpha, gamma, norm) : 0.f, red);
  if (threadIdx.x == 0) atomicAdd(loss, v);
}

// Up to FOCAL_ONE_MAX rows: ONE block that WRITES the loss -- no zero-fill before it (in a replayed CUDA graph the 4-byte
// memset node ahead of the atomics version cost 15 us of dependency latency on the step's critical chain, far more than
// the kernel) and a fixed summation order.
constexpr int FOCAL_ONE_MAX = 32768;
__global__ void __launch_bounds__(1024) focal_fwd_one_kernel(int N, int C, const float* __restrict__ lp,
                                                             const long long* __restrict__ target,
                                                             const float* __restrict__ alpha, float gamma, float norm,
                                                             float* __restrict__ loss) {
  __shared__ float red[32];
  float v = 0.f;
#pragma unroll 4
  for (int n = threadIdx.x; n < N; n += 1024) v += focal_term(n, C, lp, target, alpha, gamma, norm);
  v = focal_block_sum(v, red);
  if (threadIdx.x == 0) *loss = v;
}

__global__ void focal_bwd_kernel(int N, int C, const float* __restrict__ lp, const long long* __restrict__ target,
                                 const float* __restrict__ alpha, float gamma, float norm,
                                 const float* __restrict__ dloss, float* __restrict__ dlp) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (i64)N * C) return;
  const int n = (int)(idx / C), c = (int)(idx - (i64)n * C);
  const long long y = target[n];
  float g = 0.f;
  if (c == y) {
    const float l = lp[idx];
    const float a = alpha ? alpha[y] : 1.f;
    const float w = gamma == 0.f ? 1.f : powf(1.f - expf(l), gamma);
    g = -w * a * norm * dloss[0];
  }
  dlp[idx] = g;
}

// counter-based keep mask: keep with probability 1-p; one 64-bit mix per element (splitmix64)
// keep byte of element `idx`: one splitmix64 mix of (seed, counter base + idx)
__device__ __forceinline__ unsigned int keep_byte(unsigned long long seed, unsigned long long ctr, float p) {
  unsigned long long z = seed * 0x9E3779B97F4A7C15ull + ctr * 0xD1342543DE82EF95ull;
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);   // [0,1)
  return u >= p ? 1u : 0u;
}
// a thread produces 8 consecutive keep bytes and stores them as one 64-bit word (the byte-per-thread version spent its
// time in 1-byte stores: 30 us for the 12 MB of masks of a bench step); the value of every element is unchanged
__device__ __forceinline__ void keep_bytes8(i64 n, float p, unsigned long long seed, unsigned long long base,
                                            unsigned char* __restrict__ mask) {
  const i64 i0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i0 >= n) return;
  if (i0 + 8 <= n && (reinterpret_cast<uintptr_t>(mask) & 7) == 0) {
    unsigned long long w = 0;
#pragma unroll
    for (int e = 0; e < 8; e++) w |= (unsigned long long)keep_byte(seed, base + (unsigned long long)(i0 + e), p) << (8 * e);
    *reinterpret_cast<unsigned long long*>(mask + i0) = w;
  } else {
    for (i64 i = i0; i < n && i < i0 + 8; i++) mask[i] = (unsigned char)keep_byte(seed, base + (unsigned long long)i, p);
  }
}

__global__ void dropout_mask_kernel(i64 n, float p, unsigned long long seed, unsigned long long offset,
                                    unsigned char* __restrict__ mask) {
  keep_bytes8(n, p, seed, offset, mask);
}

// Device-resident step state for CUDA-graph replays (nothing that changes from step to step may be a kernel
// argument there): state[0] = number of optimizer steps taken, state[1] = two floats {1 - beta1^t, sqrt(1 - beta2^t)}
// for the step that is being taken.
__global__ void step_advance_kernel(unsigned long long* state, float beta1, float beta2) {
  const unsigned long long t = state[0] + 1ull;
  state[0] = t;
  float* bc = reinterpret_cast<float*>(state + 1);
  bc[0] = 1.0f - powf(beta1, (float)t);
  bc[1] = sqrtf(1.0f - powf(beta2, (float)t));
}

// keep mask whose counter base advances with the device-resident step index: offset + state[0] * per_step
__global__ void dropout_mask_dev_kernel(i64 n, float p, unsigned long long seed, const unsigned long long* __restrict__ state,
                                        unsigned long long per_step, unsigned long long offset,
                                        unsigned char* __restrict__ mask) {
  keep_bytes8(n, p, seed, offset + state[0] * per_step, mask);
}

__global__ void adam_dev_kernel(i64 n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, float lr, float beta1, float beta2, float eps, float wd,
                                const unsigned long long* __restrict__ state, float gscale) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float* bc = reinterpret_cast<const float*>(state + 1);
  const float bc1 = bc[0], bc2_sqrt = bc[1];
  const float w = p[idx];
  const float gr = fmaf(wd, w, g[idx] * gscale);
  const float mm = beta1 * m[idx] + (1.f - beta1) * gr;
  const float vv = beta2 * v[idx] + (1.f - beta2) * gr * gr;
  m[idx] = mm;
  v[idx] = vv;
  const float denom = sqrtf(vv) / bc2_sqrt + eps;
  p[idx] = w - (lr / bc1) * (mm / denom);
}

// Adam with L2 folded into the gradient (torch.optim.Adam weight_decay semantics), grads pre-scaled by gscale
__global__ void adam_kernel(i64 n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, float lr, float beta1, float beta2, float eps, float wd,
                            float bc1, float bc2_sqrt, float gscale) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float w = p[idx];
  const float gr = fmaf(wd, w, g[idx] * gscale);
  const float mm = beta1 * m[idx] + (1.f - beta1) * gr;
  const float vv = beta2 * v[idx] + (1.f - beta2) * gr * gr;
  m[idx] = mm;
  v[idx] = vv;
  const float denom = sqrtf(vv) / bc2_sqrt + eps;
  p[idx] = w - (lr / bc1) * (mm / denom);
}

}  // namespace mmdfn

using namespace mmdfn;

// one-time opt-in to > 48 KB of dynamic shared memory (C = 14..16 classes); returns the SM count
static int head_init(int* sms_out) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    MMDFN_CUDA(cudaGetDevice(&dev));
    MMDFN_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    MMDFN_CUDA(cudaFuncSetAttribute(head_fused_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXC * 900 * 4));
    MMDFN_CUDA(cudaFuncSetAttribute(head_fused_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXC * 900 * 4));
    MMDFN_CUDA(cudaFuncSetAttribute(head_fused_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXC * 900 * 4));
    MMDFN_CUDA(cudaFuncSetAttribute(head_fused_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXC * 900 * 4));
    sms = n;
  }
  *sms_out = sms;
  return 0;
}

extern "C" int mmdfn_head_fwd(int N, int C, const float* F, const unsigned char* mask, float mask_scale, int relu,
                              const float* Wc, const float* bc, float* R, float* log_prob, void* stream) {
  if (!F || !Wc || !bc || !R || !log_prob) return MMDFN_ENULL;
  if (C <= 0 || C > MAXC || N < 0) return MMDFN_EINVAL;
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 0;
  MMDFN_TRY(head_init(&sms));
  if (((uintptr_t)F | (uintptr_t)R) & 15 || (mask && ((uintptr_t)mask & 3))) return MMDFN_EINVAL;
  const int grid = min(4 * sms, ceil_div(N, 8));             // one utterance per warp while the CTAs are co-resident
  const size_t smem = (size_t)C * 900 * sizeof(float);
  if (C <= 8) head_fused_fwd_kernel<8><<<grid, 256, smem, st>>>(N, C, F, mask, mask_scale, relu, Wc, bc, R, log_prob);
  else head_fused_fwd_kernel<16><<<grid, 256, smem, st>>>(N, C, F, mask, mask_scale, relu, Wc, bc, R, log_prob);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// dlogits_ws: (N, C) scratch.  dF (3N,300), dWc (C,900), dbc (C) are overwritten.
extern "C" int mmdfn_head_bwd(int N, int C, const unsigned char* mask, float mask_scale, int relu, const float* Wc,
                              const float* R, const float* log_prob, const float* dlog_prob, float* dF, float* dWc,
                              float* dbc, int grads_zeroed, float* dlogits_ws, void* stream) {
  if (!Wc || !R || !log_prob || !dlog_prob || !dF || !dWc || !dbc || !dlogits_ws) return MMDFN_ENULL;
  if (C <= 0 || C > MAXC || N < 0) return MMDFN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    if (grads_zeroed) return 0;
    MMDFN_TRY(fill_zero(dWc, (size_t)C * 900 * sizeof(float), st));
    return fill_zero(dbc, (size_t)C * sizeof(float), st);
  }
  if (!grads_zeroed) {
    MMDFN_TRY(fill_zero(dWc, (size_t)C * 900 * sizeof(float), st));
    MMDFN_TRY(fill_zero(dbc, (size_t)C * sizeof(float), st));
  }
  if (((uintptr_t)dF | (uintptr_t)R) & 15 || (mask && ((uintptr_t)mask & 3))) return MMDFN_EINVAL;
  int sms = 0;
  MMDFN_TRY(head_init(&sms));
  const int grid = min(4 * sms, ceil_div(N, 8));
  const size_t smem = (size_t)C * 900 * sizeof(float);
  const dim3 wg(ceil_div(901, 128), ceil_div(N, HW_ROWS));
  if (C <= 8) {
    head_fused_bwd_kernel<8><<<grid, 256, smem, st>>>(N, C, mask, mask_scale, relu, Wc, R, log_prob, dlog_prob, dF, dlogits_ws);
    MMDFN_LAUNCH_CHECK();
    head_wgrad_kernel<8><<<wg, 128, 0, st>>>(N, C, R, dlogits_ws, dWc, dbc);
  } else {
    head_fused_bwd_kernel<16><<<grid, 256, smem, st>>>(N, C, mask, mask_scale, relu, Wc, R, log_prob, dlog_prob, dF, dlogits_ws);
    MMDFN_LAUNCH_CHECK();
    head_wgrad_kernel<16><<<wg, 128, 0, st>>>(N, C, R, dlogits_ws, dWc, dbc);
  }
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_focal_loss_fwd(int N, int C, const float* log_prob, const long long* target, const float* alpha,
                                    float gamma, int size_average, float* loss, void* stream) {
  if (!log_prob || !target || !loss) return MMDFN_ENULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 0) return fill_zero(loss, sizeof(float), st);
  const float norm = size_average ? 1.0f / (float)N : 1.0f;
  if (N <= FOCAL_ONE_MAX) {
    focal_fwd_one_kernel<<<1, 1024, 0, st>>>(N, C, log_prob, target, alpha, gamma, norm, loss);
  } else {
    MMDFN_TRY(fill_zero(loss, sizeof(float), st));
    focal_fwd_kernel<<<ceil_div(N, 256), 256, 0, st>>>(N, C, log_prob, target, alpha, gamma, norm, loss);
  }
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_focal_loss_bwd(int N, int C, const float* log_prob, const long long* target, const float* alpha,
                                    float gamma, int size_average, const float* dloss, float* dlog_prob,
                                    void* stream) {
  if (!log_prob || !target || !dloss || !dlog_prob) return MMDFN_ENULL;
  if (N <= 0) return 0;
  const float norm = size_average ? 1.0f / (float)N : 1.0f;
  focal_bwd_kernel<<<(unsigned)ceil_div64((i64)N * C, 256), 256, 0, (cudaStream_t)stream>>>(N, C, log_prob, target, alpha, gamma,
                                                                                    norm, dloss, dlog_prob);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_dropout_mask(long long n, float p, unsigned long long seed, unsigned long long offset,
                                  unsigned char* mask, void* stream) {
  if (!mask) return MMDFN_ENULL;
  if (n <= 0) return 0;
  dropout_mask_kernel<<<(unsigned)ceil_div64(n, 2048), 256, 0, (cudaStream_t)stream>>>(n, p, seed, offset, mask);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_adam_step(long long n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                               float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                               float grad_scale, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq) return MMDFN_ENULL;
  if (n <= 0) return 0;
  if (step <= 0) return MMDFN_EINVAL;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = sqrtf(1.0f - powf(beta2, (float)step));
  adam_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(n, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2,
                                                                      eps, weight_decay, bc1, bc2, grad_scale);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// ---- CUDA-graph friendly variants: the step index lives on the device (state: 2 x uint64, see step_advance_kernel) ----
extern "C" int mmdfn_step_advance(unsigned long long* state, float beta1, float beta2, void* stream) {
  if (!state) return MMDFN_ENULL;
  step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, beta1, beta2);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_dropout_mask_dev(long long n, float p, unsigned long long seed, const unsigned long long* state,
                                      unsigned long long per_step, unsigned long long offset, unsigned char* mask,
                                      void* stream) {
  if (!mask || !state) return MMDFN_ENULL;
  if (n <= 0) return 0;
  dropout_mask_dev_kernel<<<(unsigned)ceil_div64(n, 2048), 256, 0, (cudaStream_t)stream>>>(n, p, seed, state, per_step, offset, mask);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_adam_step_dev(long long n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                   float lr, float beta1, float beta2, float eps, float weight_decay,
                                   const unsigned long long* state, float grad_scale, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !state) return MMDFN_ENULL;
  if (n <= 0) return 0;
  adam_dev_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(n, param, grad, exp_avg, exp_avg_sq, lr, beta1,
                                                                          beta2, eps, weight_decay, state, grad_scale);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// stand-alone log_softmax over the class axis (the relation path's 'gated' head: dropout -> smax_fc -> log_softmax,
// code/model.py:1238-1239); dlogits = dlp - exp(lp) * sum_c dlp
extern "C" int mmdfn_log_softmax_fwd(int N, int C, const float* logits, float* log_prob, void* stream) {
  if (!logits || !log_prob) return MMDFN_ENULL;
  if (C <= 0 || N < 0) return MMDFN_EINVAL;
  if (N == 0) return 0;
  log_softmax_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(N, C, logits, log_prob);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_log_softmax_bwd(int N, int C, const float* log_prob, const float* dlog_prob, float* dlogits,
                                     void* stream) {
  if (!log_prob || !dlog_prob || !dlogits) return MMDFN_ENULL;
  if (C <= 0 || N < 0) return MMDFN_EINVAL;
  if (N == 0) return 0;
  log_softmax_bwd_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(N, C, log_prob, dlog_prob, dlogits);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

// ---- device-side evaluation metrics (code/run_train_erc.py:202-203,228-235: argmax -> .cpu() per batch, then sklearn) ----
// pred[n] = argmax_c log_prob[n, c] (first maximum, like torch.argmax); conf[target * C + pred] += 1.  The confusion
// matrix of a whole epoch accumulates on the device; accuracy / weighted F1 follow from its C*C counts on the host.
namespace mmdfn {
__global__ void confusion_kernel(int N, int C, const float* __restrict__ lp, const long long* __restrict__ target,
                                 long long* __restrict__ pred, unsigned long long* __restrict__ conf) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* row = lp + (i64)n * C;
  int best = 0;
  float bv = row[0];
  for (int c = 1; c < C; c++) {
    const float v = row[c];
    if (v > bv) { bv = v; best = c; }
  }
  if (pred) pred[n] = best;
  const long long t = target[n];
  if (conf && t >= 0 && t < C) atomicAdd(conf + t * C + best, 1ULL);
}
}  // namespace mmdfn

extern "C" int mmdfn_confusion_accumulate(int N, int C, const float* log_prob, const long long* target,
                                          long long* pred, unsigned long long* conf, void* stream) {
  if (!log_prob || !target) return MMDFN_ENULL;
  if (C <= 0 || N < 0) return MMDFN_EINVAL;
  if (N == 0) return 0;
  mmdfn::confusion_kernel<<<mmdfn::ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(N, C, log_prob, target, pred, conf);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
