// a13 (star row): Memory Fusion Network block, forward and backward -- replaces MFN.forward (code/model_fusion.py:62-120,
// constructor :14-60) as used by DialogueGNNModel with att_type 'mfn' (code/model.py:1263-1285, 1303-1325).
//   x (T, n, 900) = [l | a | v]  ->  out (T, n, 400) = [h_l | h_a | h_v | mem]
// Stage decomposition (derived and checked against autograd in oracle/mfn_manual.py; R = T n rows):
//   S1  pre_m = x_m W_ih^T + b_ih + b_hh                      three GEMMs
//   S2  three LSTM recurrences, ONE persistent launch         mfn_lstm_fwd_kernel (W_hh row per thread, h in smem)
//   S3  cStar_t = [c_{t-1} | c_t]                              gather
//   S4  A1 = drop(relu(cStar W11^T + b11)); att = softmax(A1 W12^T + b12); attended = att * cStar     GEMMs + warp-per-row softmax
//   S5  A2 = drop(relu(attended W21^T + b21)); cHat = tanh(A2 W22^T + b22)
//   S6  U_k = attended Wgk1[:, :600]^T + bgk1  (k = 1, 2: gamma_k_fc1 split into its attended / memory columns)
//   S7  memory recurrence, ONE persistent launch               mfn_mem_fwd_kernel
//       q_k = drop(relu(U_k[t] + mem Wgk1[:, 600:]^T)); gamma_k = sigmoid(q_k Wgk2^T + bgk2); mem = gamma1 mem + gamma2 cHat[t]
// The four Dropout(0.2) layers of the block take optional keep masks (train mode).  All dense products go through the
// library's GEMM (tcgen05 3xTF32 above its size threshold); weight gradients are TN GEMMs over all R rows.
#include "internal.cuh"
#include <math.h>

namespace mmdfn {
namespace {

constexpr int MH = 100;          // hidden / memory width
constexpr int MG = 400;          // LSTM gate rows
constexpr int MD = 300;          // per-modality input width
constexpr int MW = 600;          // attention window width
constexpr int MNB = 4;           // sequences per CTA in the recurrences
constexpr int MTH = 400;         // threads per CTA in the recurrences

__device__ __forceinline__ float dot100(const float (&w)[MH], const float* __restrict__ v) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int k = 0; k < MH; k += 4) {
    const float4 x = *reinterpret_cast<const float4*>(v + k);
    a0 = fmaf(w[k], x.x, a0);
    a1 = fmaf(w[k + 1], x.y, a1);
    a2 = fmaf(w[k + 2], x.z, a2);
    a3 = fmaf(w[k + 3], x.w, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

struct LstmArgs {
  int T, n;
  const float* pre;            // (3, R, 400)
  const float* w_hh[3];        // (400, 100)
  float* gates;                // (3, R, 400) activated i f g o
  float* c_all;                // ((T+1) n, 300), rows 0..n-1 zero
  float* out;                  // (R, 400): h_m -> columns 100 m ..
};

// grid (ceil(n / MNB), 3); thread j = gate row j (its W_hh row in registers), thread (s, u) = pointwise item
__global__ void __launch_bounds__(MTH, 1) mfn_lstm_fwd_kernel(LstmArgs p) {
  __shared__ __align__(16) float hs[MNB][MH];
  __shared__ float gp[MNB][MG];
  const int tid = threadIdx.x, m = blockIdx.y, seq0 = blockIdx.x * MNB;
  const i64 R = (i64)p.T * p.n;
  float w[MH];
#pragma unroll
  for (int k = 0; k < MH; k++) w[k] = p.w_hh[m][tid * MH + k];
  const int s_me = tid / MH, u_me = tid - MH * s_me;
  const bool v_me = seq0 + s_me < p.n;
  hs[s_me][u_me] = 0.f;
  float c = 0.f;
  const float* pre = p.pre + (i64)m * R * MG;
  float* gates = p.gates + (i64)m * R * MG;
  __syncthreads();
  for (int t = 0; t < p.T; t++) {
#pragma unroll
    for (int s = 0; s < MNB; s++) {
      float a = 0.f;
      if (seq0 + s < p.n) a = pre[((i64)t * p.n + seq0 + s) * MG + tid] + dot100(w, hs[s]);
      gp[s][tid] = a;
    }
    __syncthreads();
    if (v_me) {
      const float i_ = sigmoidf_(gp[s_me][u_me]), f_ = sigmoidf_(gp[s_me][MH + u_me]), g_ = tanhf(gp[s_me][2 * MH + u_me]),
                  o_ = sigmoidf_(gp[s_me][3 * MH + u_me]);
      c = f_ * c + i_ * g_;
      const float h = o_ * tanhf(c);
      const i64 row = (i64)t * p.n + seq0 + s_me;
      float* g = gates + row * MG;
      g[u_me] = i_; g[MH + u_me] = f_; g[2 * MH + u_me] = g_; g[3 * MH + u_me] = o_;
      p.c_all[(row + p.n) * MD + MH * m + u_me] = c;
      p.out[row * MG + MH * m + u_me] = h;
      hs[s_me][u_me] = h;
    }
    __syncthreads();
  }
}

struct LstmBwdArgs {
  int T, n;
  const float* w_hh[3];
  const float* gates;          // (3, R, 400)
  const float* c_all;
  const float* dout;           // (R, 400)
  const float* dc_ext;         // (R, 300): gradient reaching c_t through the attention window
  float* dpre;                 // (3, R, 400)
};

// thread (q, k): partial sum over gate rows 100 q .. 100 q + 99 of (dpre W_hh)[k]; thread (s, u) = pointwise item
__global__ void __launch_bounds__(MTH, 1) mfn_lstm_bwd_kernel(LstmBwdArgs p) {
  __shared__ __align__(16) float dp[MNB][MG];
  __shared__ float part[4][MNB][MH];
  const int tid = threadIdx.x, m = blockIdx.y, seq0 = blockIdx.x * MNB;
  const i64 R = (i64)p.T * p.n;
  const int q = tid / MH, k = tid - MH * q;
  float w[MH];
#pragma unroll
  for (int j = 0; j < MH; j++) w[j] = p.w_hh[m][(MH * q + j) * MH + k];
  const int s_me = q, u_me = k;
  const bool v_me = seq0 + s_me < p.n;
#pragma unroll
  for (int a = 0; a < 4; a++) part[a][s_me][u_me] = 0.f;
  float cc = 0.f;
  const float* gates = p.gates + (i64)m * R * MG;
  float* dpre = p.dpre + (i64)m * R * MG;
  __syncthreads();
  for (int t = p.T - 1; t >= 0; t--) {
    const i64 row = (i64)t * p.n + seq0 + s_me;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    if (v_me) {
      const float ch = (part[0][s_me][u_me] + part[1][s_me][u_me]) + (part[2][s_me][u_me] + part[3][s_me][u_me]);
      const float* g = gates + row * MG;
      const float i_ = g[u_me], f_ = g[MH + u_me], g_ = g[2 * MH + u_me], o_ = g[3 * MH + u_me];
      const float c_t = p.c_all[(row + p.n) * MD + MH * m + u_me], c_prev = p.c_all[row * MD + MH * m + u_me];
      const float tc = tanhf(c_t);
      const float dh = p.dout[row * MG + MH * m + u_me] + ch;
      const float dc = p.dc_ext[row * MD + MH * m + u_me] + cc + dh * o_ * (1.f - tc * tc);
      d0 = dc * g_ * i_ * (1.f - i_);
      d1 = dc * c_prev * f_ * (1.f - f_);
      d2 = dc * i_ * (1.f - g_ * g_);
      d3 = dh * tc * o_ * (1.f - o_);
      cc = dc * f_;
      float* d = dpre + row * MG;
      d[u_me] = d0; d[MH + u_me] = d1; d[2 * MH + u_me] = d2; d[3 * MH + u_me] = d3;
    }
    __syncthreads();                       // every thread has read `part`
    dp[s_me][u_me] = d0; dp[s_me][MH + u_me] = d1; dp[s_me][2 * MH + u_me] = d2; dp[s_me][3 * MH + u_me] = d3;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < MNB; s++) part[q][s][k] = dot100(w, &dp[s][MH * q]);
    __syncthreads();
  }
}

struct MemArgs {
  int T, n;
  const float* w1m[2];         // gamma_k_fc1.weight (100, 700): memory columns 600..699
  const float* w2[2];          // gamma_k_fc2.weight (100, 100)
  const float* b2[2];
  const float* U;              // (2, R, 100)
  const float* chat;           // (R, 100)
  const unsigned char* mq[2];  // optional keep masks of q_k (R, 100)
  float scale;
  float* qd;                   // (2, R, 100) masked relu outputs
  float* gam;                  // (2, R, 100)
  float* mem_prev;             // (R, 100)
  float* out;                  // (R, 400): mem -> columns 300..399
};

// thread (r, j): r = 0, 1 -> fc1 memory half of gamma_{r+1}; r = 2, 3 -> fc2 of gamma_{r-1}; thread (s, u) = memory item
__global__ void __launch_bounds__(MTH, 1) mfn_mem_fwd_kernel(MemArgs p) {
  __shared__ __align__(16) float mem_s[MNB][MH];
  __shared__ __align__(16) float qd_s[2][MNB][MH];
  __shared__ float gam_s[2][MNB][MH];
  const int tid = threadIdx.x, seq0 = blockIdx.x * MNB;
  const i64 R = (i64)p.T * p.n;
  const int r = tid / MH, j = tid - MH * r, kk = r & 1;
  float w[MH];
#pragma unroll
  for (int c = 0; c < MH; c++) w[c] = r < 2 ? p.w1m[kk][j * 700 + 600 + c] : p.w2[kk][j * MH + c];
  const float bias = r < 2 ? 0.f : p.b2[kk][j];
  const int s_me = r, u_me = j;
  const bool v_me = seq0 + s_me < p.n;
  mem_s[s_me][u_me] = 0.f;
  __syncthreads();
  for (int t = 0; t < p.T; t++) {
    if (r < 2) {
#pragma unroll
      for (int s = 0; s < MNB; s++) {
        float qv = 0.f;
        if (seq0 + s < p.n) {
          const i64 row = (i64)t * p.n + seq0 + s;
          qv = fmaxf(p.U[((i64)kk * R + row) * MH + j] + dot100(w, mem_s[s]), 0.f);
          if (p.mq[kk]) qv = p.mq[kk][row * MH + j] ? qv * p.scale : 0.f;
          p.qd[((i64)kk * R + row) * MH + j] = qv;
          if (kk == 0) p.mem_prev[row * MH + j] = mem_s[s][j];
        }
        qd_s[kk][s][j] = qv;
      }
    }
    __syncthreads();
    if (r >= 2) {
#pragma unroll
      for (int s = 0; s < MNB; s++) {
        float gv = 0.f;
        if (seq0 + s < p.n) {
          gv = sigmoidf_(bias + dot100(w, qd_s[kk][s]));
          p.gam[((i64)kk * R + (i64)t * p.n + seq0 + s) * MH + j] = gv;
        }
        gam_s[kk][s][j] = gv;
      }
    }
    __syncthreads();
    if (v_me) {
      const i64 row = (i64)t * p.n + seq0 + s_me;
      const float mv = gam_s[0][s_me][u_me] * mem_s[s_me][u_me] + gam_s[1][s_me][u_me] * p.chat[row * MH + u_me];
      mem_s[s_me][u_me] = mv;
      p.out[row * MG + 3 * MH + u_me] = mv;
    }
    __syncthreads();
  }
}

struct MemBwdArgs {
  int T, n;
  const float* w1m[2];
  const float* w2[2];
  const float* chat;
  const float* qd;
  const float* gam;
  const float* mem_prev;
  const float* dout;           // (R, 400): columns 300..399
  int has_mask;
  float scale;
  float* dchat;                // (R, 100)
  float* dz;                   // (2, R, 100): d / d(gamma pre-activation)
  float* dU;                   // (2, R, 100)
};

// thread (r, j): r = 0, 1 -> column j of gamma_{r+1}_fc2 (dq = dz W2); r = 2, 3 -> memory column j of gamma_{r-1}_fc1 (carry += dq W1m)
__global__ void __launch_bounds__(MTH, 1) mfn_mem_bwd_kernel(MemBwdArgs p) {
  __shared__ __align__(16) float dz_s[2][MNB][MH];
  __shared__ __align__(16) float dq_s[2][MNB][MH];
  __shared__ float part[2][MNB][MH];
  const int tid = threadIdx.x, seq0 = blockIdx.x * MNB;
  const i64 R = (i64)p.T * p.n;
  const int r = tid / MH, j = tid - MH * r, kk = r & 1;
  float w[MH];
#pragma unroll
  for (int c = 0; c < MH; c++) w[c] = r < 2 ? p.w2[kk][c * MH + j] : p.w1m[kk][c * 700 + 600 + j];
  const int s_me = r, u_me = j;
  const bool v_me = seq0 + s_me < p.n;
  part[0][s_me][u_me] = 0.f;
  part[1][s_me][u_me] = 0.f;
  float carry = 0.f;
  const float ind_scale = p.has_mask ? p.scale : 1.f;
  __syncthreads();
  for (int t = p.T - 1; t >= 0; t--) {
    float z1 = 0.f, z2 = 0.f;
    if (v_me) {
      const i64 row = (i64)t * p.n + seq0 + s_me;
      const float dmem = p.dout[row * MG + 3 * MH + u_me] + carry + part[0][s_me][u_me] + part[1][s_me][u_me];
      const float g1 = p.gam[row * MH + u_me], g2 = p.gam[(R + row) * MH + u_me];
      const float ch = p.chat[row * MH + u_me];
      p.dchat[row * MH + u_me] = dmem * g2;
      carry = dmem * g1;
      z1 = dmem * p.mem_prev[row * MH + u_me] * g1 * (1.f - g1);
      z2 = dmem * ch * g2 * (1.f - g2);
      p.dz[row * MH + u_me] = z1;
      p.dz[(R + row) * MH + u_me] = z2;
    }
    __syncthreads();                       // every thread has read `part`
    dz_s[0][s_me][u_me] = z1;
    dz_s[1][s_me][u_me] = z2;
    __syncthreads();
    if (r < 2) {
#pragma unroll
      for (int s = 0; s < MNB; s++) {
        float dq = 0.f;
        if (seq0 + s < p.n) {
          const i64 row = (i64)t * p.n + seq0 + s;
          dq = p.qd[((i64)kk * R + row) * MH + j] != 0.f ? dot100(w, dz_s[kk][s]) * ind_scale : 0.f;
          p.dU[((i64)kk * R + row) * MH + j] = dq;
        }
        dq_s[kk][s][j] = dq;
      }
    }
    __syncthreads();
    if (r >= 2) {
#pragma unroll
      for (int s = 0; s < MNB; s++) part[kk][s][j] = dot100(w, dq_s[kk][s]);
    }
    __syncthreads();
  }
}

// ---- pointwise / row kernels ------------------------------------------------------------------------------------------
__global__ void mfn_bias_sum_kernel(const float* a0, const float* b0, const float* a1, const float* b1, const float* a2, const float* b2, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * MG) return;
  const int m = i / MG, c = i - MG * m;
  const float* a = m == 0 ? a0 : m == 1 ? a1 : a2;
  const float* b = m == 0 ? b0 : m == 1 ? b1 : b2;
  out[i] = a[c] + b[c];
}

// cstar[(t, s)] = [c_all[t, s] | c_all[t + 1, s]]
__global__ void mfn_cstar_kernel(i64 R, int n, const float* __restrict__ c_all, float* __restrict__ cstar) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * MW) return;
  const i64 row = idx / MW;
  const int c = (int)(idx - row * MW);
  cstar[idx] = c < MD ? c_all[row * MD + c] : c_all[(row + n) * MD + c - MD];
}

// dc_ext[t] = dcstar[t][:, 300:] + dcstar[t + 1][:, :300]      (c_{-1} = 0 is a constant)
__global__ void mfn_dcext_kernel(i64 R, int n, const float* __restrict__ dcstar, float* __restrict__ dc_ext) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * MD) return;
  const i64 row = idx / MD;
  const int c = (int)(idx - row * MD);
  float v = dcstar[row * MW + MD + c];
  if (row + n < R) v += dcstar[(row + n) * MW + c];
  dc_ext[idx] = v;
}

__global__ void mfn_mask_kernel(i64 n, float* x, const unsigned char* __restrict__ m, float scale) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = m[i] ? x[i] * scale : 0.f;
}
__global__ void mfn_tanh_kernel(i64 n, float* x) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = tanhf(x[i]);
}
// dx = dy * ind_scale * [y != 0]   (through dropout(relu(.)): y is the masked, scaled output); in place on dy allowed
__global__ void mfn_relu_bwd_kernel(i64 n, const float* dy, const float* __restrict__ y, float ind_scale, float* dx) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = y[i] != 0.f ? dy[i] * ind_scale : 0.f;
}
__global__ void mfn_tanh_bwd_kernel(i64 n, const float* dy, const float* __restrict__ y, float* dx) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = dy[i] * (1.f - y[i] * y[i]);
}

// att (in: logits, out: softmax) and attended = att * cstar; one warp per row of 600
__global__ void mfn_softmax_mul_kernel(i64 R, float* att, const float* __restrict__ cstar, float* __restrict__ attended) {
  const i64 row = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  float* a = att + row * MW;
  float mx = -INFINITY;
  for (int c = lane; c < MW; c += 32) mx = fmaxf(mx, a[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < MW; c += 32) sum += expf(a[c] - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int c = lane; c < MW; c += 32) {
    const float v = expf(a[c] - mx) * inv;
    a[c] = v;
    attended[row * MW + c] = v * cstar[row * MW + c];
  }
}

// dS = att * (datt - sum(datt * att)), datt = dattended * cstar; dcstar = dattended * att
__global__ void mfn_softmax_mul_bwd_kernel(i64 R, const float* __restrict__ dattended, const float* __restrict__ cstar,
                                           const float* __restrict__ att, float* __restrict__ dS, float* __restrict__ dcstar) {
  const i64 row = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const i64 o = row * MW;
  float dotv = 0.f;
  for (int c = lane; c < MW; c += 32) dotv += dattended[o + c] * cstar[o + c] * att[o + c];
  dotv = warp_sum(dotv);
  for (int c = lane; c < MW; c += 32) {
    const float da = dattended[o + c], a = att[o + c];
    dS[o + c] = a * (da * cstar[o + c] - dotv);
    dcstar[o + c] = da * a;
  }
}

// x[t, b, 300 j + c] = F[(perm[j] N + off_b + t) 300 + c] (t < L_b) or 0: the stacked per-modality node features
// (3N, 300) as the padded, time-major (T, B, 900) window the reference builds (code/model.py:1264-1276, 1304-1316)
__global__ void mfn_pack_kernel(int T, int B, int N, const int* __restrict__ dia_off, int p0, int p1, int p2,
                                const float* __restrict__ F, float* __restrict__ x) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (i64)T * B * 900) return;
  const int c9 = (int)(idx % 900);
  const i64 tb = idx / 900;
  const int b = (int)(tb % B), t = (int)(tb / B);
  const int j = c9 / MD, c = c9 - MD * j;
  const int pm = j == 0 ? p0 : j == 1 ? p1 : p2;
  const int off = dia_off[b], L = dia_off[b + 1] - off;
  x[idx] = t < L ? F[((i64)pm * N + off + t) * MD + c] : 0.f;
}
__global__ void mfn_pack_bwd_kernel(int T, int B, int N, const int* __restrict__ dia_off, int p0, int p1, int p2,
                                    const float* __restrict__ dx, float* __restrict__ dF) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;          // one element of dF (3N, 300)
  if (idx >= (i64)3 * N * MD) return;
  const int c = (int)(idx % MD);
  const i64 r = idx / MD;
  const int pm = (int)(r / N), node = (int)(r - (i64)pm * N);
  const int j = p0 == pm ? 0 : p1 == pm ? 1 : 2;
  int lo = 0, hi = B;                                                    // dialogue of the node: dia_off[b] <= node < dia_off[b + 1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (dia_off[mid] <= node) lo = mid; else hi = mid;
  }
  const int t = node - dia_off[lo];
  dF[idx] = dx[((i64)t * B + lo) * 900 + MD * j + c];
}
// feat[off_b + t, :] = out[t, b, :] (t < L_b), width 400; the backward zero-fills the padding rows
__global__ void mfn_unpad_kernel(int T, int B, const int* __restrict__ dia_off, const float* __restrict__ out, float* __restrict__ feat) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (i64)T * B * MG) return;
  const int c = (int)(idx % MG);
  const i64 tb = idx / MG;
  const int b = (int)(tb % B), t = (int)(tb / B);
  const int off = dia_off[b], L = dia_off[b + 1] - off;
  if (t < L) feat[((i64)off + t) * MG + c] = out[idx];
}
__global__ void mfn_unpad_bwd_kernel(int T, int B, const int* __restrict__ dia_off, const float* __restrict__ dfeat, float* __restrict__ dout) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (i64)T * B * MG) return;
  const int c = (int)(idx % MG);
  const i64 tb = idx / MG;
  const int b = (int)(tb % B), t = (int)(tb / B);
  const int off = dia_off[b], L = dia_off[b + 1] - off;
  dout[idx] = t < L ? dfeat[((i64)off + t) * MG + c] : 0.f;
}
// y = relu(x) * keep * scale (dropout then ReLU, code/model.py:1290-1291 / 1327-1328)
__global__ void mfn_relu_mask_kernel(i64 n, const float* __restrict__ x, const unsigned char* __restrict__ m, float scale, float* __restrict__ y) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = fmaxf(x[i], 0.f);
  y[i] = m ? (m[i] ? v * scale : 0.f) : v;
}

inline unsigned nb(i64 n) { return (unsigned)ceil_div64(n, 256); }

// parameter slots (state_dict order of MFN, code/model_fusion.py:39-60)
enum {
  P_WIH = 0, P_WHH = 1, P_BIH = 2, P_BHH = 3,      // + 4 m for lstm_l / lstm_a / lstm_v
  P_A11W = 12, P_A11B, P_A12W, P_A12B, P_A21W, P_A21B, P_A22W, P_A22B, P_G11W, P_G11B, P_G12W, P_G12B, P_G21W, P_G21B, P_G22W, P_G22B,
  P_COUNT = 28
};

// forward workspace (floats); everything the backward needs except x and out
struct MfnWs {
  i64 gates, c_all, cstar, a1, att, attended, a2, chat, qd, gam, mem_prev, pre, u, bsum, total;
};
MfnWs mfn_ws(i64 T, i64 n) {
  const i64 R = T * n;
  MfnWs w;
  i64 o = 0;
  w.gates = o; o += 3 * R * MG;
  w.c_all = o; o += (T + 1) * n * MD;
  w.cstar = o; o += R * MW;
  w.a1 = o; o += R * MH;
  w.att = o; o += R * MW;
  w.attended = o; o += R * MW;
  w.a2 = o; o += R * MH;
  w.chat = o; o += R * MH;
  w.qd = o; o += 2 * R * MH;
  w.gam = o; o += 2 * R * MH;
  w.mem_prev = o; o += R * MH;
  w.pre = o; o += 3 * R * MG;          // scratch of the forward
  w.u = o; o += 2 * R * MH;            // scratch of the forward
  w.bsum = o; o += 3 * MG + 16;
  w.total = o;
  return w;
}

}  // namespace
}  // namespace mmdfn

using namespace mmdfn;

extern "C" long long mmdfn_mfn_ws_floats(int T, int n) { return (T < 0 || n < 0) ? 0 : mfn_ws(T, n).total; }

/* backward scratch: dchat dz(2) dU(2) (R x 100 each) | dattended dS dcstar (R x 600) | dA (R x 100) | dc_ext (R x 300) | dpre (3 R x 400) */
extern "C" long long mmdfn_mfn_bwd_ws_floats(int T, int n) {
  const i64 R = (i64)(T < 0 ? 0 : T) * (n < 0 ? 0 : n);
  return R * (5 * MH + 3 * MW + MH + MD + 3 * MG);
}

extern "C" int mmdfn_mfn_fwd(int T, int n, const float* x, const float* const* P, const unsigned char* const* masks,
                             float mask_scale, float* out, float* ws, void* stream) {
  if (!x || !P || !out || !ws) return MMDFN_ENULL;
  if (T < 0 || n < 0) return MMDFN_EINVAL;
  if (T == 0 || n == 0) return 0;
  for (int i = 0; i < P_COUNT; i++)
    if (!P[i]) return MMDFN_ENULL;
  cudaStream_t st = (cudaStream_t)stream;
  const i64 R = (i64)T * n;
  const MfnWs w = mfn_ws(T, n);
  float* pre = ws + w.pre;
  float* bsum = ws + w.bsum;
  const unsigned char* m_a1 = masks ? masks[0] : nullptr;
  const unsigned char* m_a2 = masks ? masks[1] : nullptr;
  const unsigned char* m_q1 = masks ? masks[2] : nullptr;
  const unsigned char* m_q2 = masks ? masks[3] : nullptr;
  // S1
  mfn_bias_sum_kernel<<<ceil_div(3 * MG, 256), 256, 0, st>>>(P[P_BIH], P[P_BHH], P[4 + P_BIH], P[4 + P_BHH], P[8 + P_BIH], P[8 + P_BHH], bsum);
  MMDFN_LAUNCH_CHECK();
  for (int m = 0; m < 3; m++)
    MMDFN_TRY(gemm(false, true, (int)R, MG, MD, 1.f, x + MD * m, 3 * MD, P[4 * m + P_WIH], MD, 0.f, pre + (i64)m * R * MG, MG, bsum + MG * m, 0, st));
  // S2
  MMDFN_TRY(fill_zero(ws + w.c_all, (size_t)n * MD * sizeof(float), st));
  {
    LstmArgs a;
    a.T = T; a.n = n; a.pre = pre; a.gates = ws + w.gates; a.c_all = ws + w.c_all; a.out = out;
    for (int m = 0; m < 3; m++) a.w_hh[m] = P[4 * m + P_WHH];
    mfn_lstm_fwd_kernel<<<dim3(ceil_div(n, MNB), 3), MTH, 0, st>>>(a);
    MMDFN_LAUNCH_CHECK();
  }
  // S3 .. S6
  float* cstar = ws + w.cstar;
  float* a1 = ws + w.a1;
  float* att = ws + w.att;
  float* attended = ws + w.attended;
  float* a2 = ws + w.a2;
  float* chat = ws + w.chat;
  float* u = ws + w.u;
  mfn_cstar_kernel<<<nb(R * MW), 256, 0, st>>>(R, n, ws + w.c_all, cstar);
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(gemm(false, true, (int)R, MH, MW, 1.f, cstar, MW, P[P_A11W], MW, 0.f, a1, MH, P[P_A11B], 1, st));
  if (m_a1) { mfn_mask_kernel<<<nb(R * MH), 256, 0, st>>>(R * MH, a1, m_a1, mask_scale); MMDFN_LAUNCH_CHECK(); }
  MMDFN_TRY(gemm(false, true, (int)R, MW, MH, 1.f, a1, MH, P[P_A12W], MH, 0.f, att, MW, P[P_A12B], 0, st));
  mfn_softmax_mul_kernel<<<(unsigned)ceil_div64(R, 8), 256, 0, st>>>(R, att, cstar, attended);
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(gemm(false, true, (int)R, MH, MW, 1.f, attended, MW, P[P_A21W], MW, 0.f, a2, MH, P[P_A21B], 1, st));
  if (m_a2) { mfn_mask_kernel<<<nb(R * MH), 256, 0, st>>>(R * MH, a2, m_a2, mask_scale); MMDFN_LAUNCH_CHECK(); }
  MMDFN_TRY(gemm(false, true, (int)R, MH, MH, 1.f, a2, MH, P[P_A22W], MH, 0.f, chat, MH, P[P_A22B], 0, st));
  mfn_tanh_kernel<<<nb(R * MH), 256, 0, st>>>(R * MH, chat);
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(gemm(false, true, (int)R, MH, MW, 1.f, attended, MW, P[P_G11W], MW + MH, 0.f, u, MH, P[P_G11B], 0, st));
  MMDFN_TRY(gemm(false, true, (int)R, MH, MW, 1.f, attended, MW, P[P_G21W], MW + MH, 0.f, u + R * MH, MH, P[P_G21B], 0, st));
  // S7
  {
    MemArgs a;
    a.T = T; a.n = n;
    a.w1m[0] = P[P_G11W]; a.w1m[1] = P[P_G21W];
    a.w2[0] = P[P_G12W]; a.w2[1] = P[P_G22W];
    a.b2[0] = P[P_G12B]; a.b2[1] = P[P_G22B];
    a.U = u; a.chat = chat; a.mq[0] = m_q1; a.mq[1] = m_q2; a.scale = mask_scale;
    a.qd = ws + w.qd; a.gam = ws + w.gam; a.mem_prev = ws + w.mem_prev; a.out = out;
    mfn_mem_fwd_kernel<<<ceil_div(n, MNB), MTH, 0, st>>>(a);
    MMDFN_LAUNCH_CHECK();
  }
  return 0;
}

/* dP[i] receive "=" for i < 28 (out_fc1 / out_fc2 are never used by the forward: no gradient); bias_ih and bias_hh of an
 * LSTM receive the same gradient.  masks as in the forward (their presence selects the dropout scale of the indicators). */
extern "C" int mmdfn_mfn_bwd(int T, int n, const float* x, const float* const* P, const unsigned char* const* masks,
                             float mask_scale, const float* out, const float* ws_fwd, const float* dout, float* dx,
                             float* const* dP, float* ws, void* stream) {
  if (!x || !P || !out || !ws_fwd || !dout || !dx || !dP || !ws) return MMDFN_ENULL;
  if (T < 0 || n < 0) return MMDFN_EINVAL;
  if (T == 0 || n == 0) return 0;
  for (int i = 0; i < P_COUNT; i++)
    if (!P[i] || !dP[i]) return MMDFN_ENULL;
  cudaStream_t st = (cudaStream_t)stream;
  const i64 R = (i64)T * n;
  const MfnWs w = mfn_ws(T, n);
  const float* gates = ws_fwd + w.gates;
  const float* c_all = ws_fwd + w.c_all;
  const float* cstar = ws_fwd + w.cstar;
  const float* a1 = ws_fwd + w.a1;
  const float* att = ws_fwd + w.att;
  const float* attended = ws_fwd + w.attended;
  const float* a2 = ws_fwd + w.a2;
  const float* chat = ws_fwd + w.chat;
  const float* qd = ws_fwd + w.qd;
  const float* gam = ws_fwd + w.gam;
  const float* mem_prev = ws_fwd + w.mem_prev;
  float* dchat = ws;
  float* dz = dchat + R * MH;
  float* dU = dz + 2 * R * MH;
  float* dattended = dU + 2 * R * MH;
  float* dS = dattended + R * MW;
  float* dcstar = dS + R * MW;
  float* dA = dcstar + R * MW;
  float* dc_ext = dA + R * MH;
  float* dpre = dc_ext + R * MD;
  const bool has_mask = masks != nullptr;
  const float ind = has_mask ? mask_scale : 1.f;
  // ---- S7 backward
  {
    MemBwdArgs a;
    a.T = T; a.n = n;
    a.w1m[0] = P[P_G11W]; a.w1m[1] = P[P_G21W];
    a.w2[0] = P[P_G12W]; a.w2[1] = P[P_G22W];
    a.chat = chat; a.qd = qd; a.gam = gam; a.mem_prev = mem_prev; a.dout = dout;
    a.has_mask = has_mask ? 1 : 0; a.scale = mask_scale;
    a.dchat = dchat; a.dz = dz; a.dU = dU;
    mfn_mem_bwd_kernel<<<ceil_div(n, MNB), MTH, 0, st>>>(a);
    MMDFN_LAUNCH_CHECK();
  }
  for (int k = 0; k < 2; k++) {
    const float* dzk = dz + (i64)k * R * MH;
    const float* dUk = dU + (i64)k * R * MH;
    const float* qk = qd + (i64)k * R * MH;
    const int pw1 = k == 0 ? P_G11W : P_G21W, pw2 = k == 0 ? P_G12W : P_G22W;
    MMDFN_TRY(gemm(true, false, MH, MH, (int)R, 1.f, dzk, MH, qk, MH, 0.f, dP[pw2], MH, nullptr, 0, st));              // fc2.weight = dz^T q
    MMDFN_TRY(colsum((int)R, MH, dzk, MH, 0.f, dP[pw2 + 1], st));
    MMDFN_TRY(gemm(true, false, MH, MW, (int)R, 1.f, dUk, MH, attended, MW, 0.f, dP[pw1], MW + MH, nullptr, 0, st));   // fc1.weight[:, :600]
    MMDFN_TRY(gemm(true, false, MH, MH, (int)R, 1.f, dUk, MH, mem_prev, MH, 0.f, dP[pw1] + MW, MW + MH, nullptr, 0, st));   // fc1.weight[:, 600:]
    MMDFN_TRY(colsum((int)R, MH, dUk, MH, 0.f, dP[pw1 + 1], st));
    MMDFN_TRY(gemm(false, false, (int)R, MW, MH, 1.f, dUk, MH, P[pw1], MW + MH, k == 0 ? 0.f : 1.f, dattended, MW, nullptr, 0, st));
  }
  // ---- S5 backward
  mfn_tanh_bwd_kernel<<<nb(R * MH), 256, 0, st>>>(R * MH, dchat, chat, dchat);                 // dP2 in place
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(gemm(true, false, MH, MH, (int)R, 1.f, dchat, MH, a2, MH, 0.f, dP[P_A22W], MH, nullptr, 0, st));
  MMDFN_TRY(colsum((int)R, MH, dchat, MH, 0.f, dP[P_A22B], st));
  MMDFN_TRY(gemm(false, false, (int)R, MH, MH, 1.f, dchat, MH, P[P_A22W], MH, 0.f, dA, MH, nullptr, 0, st));
  mfn_relu_bwd_kernel<<<nb(R * MH), 256, 0, st>>>(R * MH, dA, a2, ind, dA);
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(gemm(true, false, MH, MW, (int)R, 1.f, dA, MH, attended, MW, 0.f, dP[P_A21W], MW, nullptr, 0, st));
  MMDFN_TRY(colsum((int)R, MH, dA, MH, 0.f, dP[P_A21B], st));
  MMDFN_TRY(gemm(false, false, (int)R, MW, MH, 1.f, dA, MH, P[P_A21W], MW, 1.f, dattended, MW, nullptr, 0, st));
  // ---- S4 backward
  mfn_softmax_mul_bwd_kernel<<<(unsigned)ceil_div64(R, 8), 256, 0, st>>>(R, dattended, cstar, att, dS, dcstar);
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(gemm(true, false, MW, MH, (int)R, 1.f, dS, MW, a1, MH, 0.f, dP[P_A12W], MH, nullptr, 0, st));
  MMDFN_TRY(colsum((int)R, MW, dS, MW, 0.f, dP[P_A12B], st));
  MMDFN_TRY(gemm(false, false, (int)R, MH, MW, 1.f, dS, MW, P[P_A12W], MH, 0.f, dA, MH, nullptr, 0, st));
  mfn_relu_bwd_kernel<<<nb(R * MH), 256, 0, st>>>(R * MH, dA, a1, ind, dA);
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(gemm(true, false, MH, MW, (int)R, 1.f, dA, MH, cstar, MW, 0.f, dP[P_A11W], MW, nullptr, 0, st));
  MMDFN_TRY(colsum((int)R, MH, dA, MH, 0.f, dP[P_A11B], st));
  MMDFN_TRY(gemm(false, false, (int)R, MW, MH, 1.f, dA, MH, P[P_A11W], MW, 1.f, dcstar, MW, nullptr, 0, st));
  // ---- S3 backward
  mfn_dcext_kernel<<<nb(R * MD), 256, 0, st>>>(R, n, dcstar, dc_ext);
  MMDFN_LAUNCH_CHECK();
  // ---- S2 + S1 backward
  {
    LstmBwdArgs a;
    a.T = T; a.n = n; a.gates = gates; a.c_all = c_all; a.dout = dout; a.dc_ext = dc_ext; a.dpre = dpre;
    for (int m = 0; m < 3; m++) a.w_hh[m] = P[4 * m + P_WHH];
    mfn_lstm_bwd_kernel<<<dim3(ceil_div(n, MNB), 3), MTH, 0, st>>>(a);
    MMDFN_LAUNCH_CHECK();
  }
  for (int m = 0; m < 3; m++) {
    const float* dp = dpre + (i64)m * R * MG;
    // weight_hh = dpre[1:]^T h[:-1]  (h_{-1} = 0); with T == 1 there is no contribution
    if (T > 1) {
      MMDFN_TRY(gemm(true, false, MG, MH, (int)(R - n), 1.f, dp + (i64)n * MG, MG, out + MH * m, MG, 0.f, dP[4 * m + P_WHH], MH, nullptr, 0, st));
    } else {
      MMDFN_TRY(fill_zero(dP[4 * m + P_WHH], (size_t)MG * MH * sizeof(float), st));
    }
    MMDFN_TRY(gemm(true, false, MG, MD, (int)R, 1.f, dp, MG, x + MD * m, 3 * MD, 0.f, dP[4 * m + P_WIH], MD, nullptr, 0, st));
    MMDFN_TRY(colsum((int)R, MG, dp, MG, 0.f, dP[4 * m + P_BIH], st));
    MMDFN_CUDA(cudaMemcpyAsync(dP[4 * m + P_BHH], dP[4 * m + P_BIH], MG * sizeof(float), cudaMemcpyDeviceToDevice, st));
    MMDFN_TRY(gemm(false, false, (int)R, MD, MG, 1.f, dp, MG, P[4 * m + P_WIH], MD, 0.f, dx + MD * m, 3 * MD, nullptr, 0, st));
  }
  return 0;
}

/* ---- glue of the 'mfn' head (code/model.py:1263-1291, 1303-1330): pad / un-pad and dropout + ReLU ---- */
extern "C" int mmdfn_mfn_pack_fwd(int T, int B, int N, const int* dia_off, int p0, int p1, int p2, const float* F, float* x, void* stream) {
  if (!dia_off || !F || !x) return MMDFN_ENULL;
  if (T <= 0 || B <= 0) return 0;
  mfn_pack_kernel<<<nb((i64)T * B * 900), 256, 0, (cudaStream_t)stream>>>(T, B, N, dia_off, p0, p1, p2, F, x);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
extern "C" int mmdfn_mfn_pack_bwd(int T, int B, int N, const int* dia_off, int p0, int p1, int p2, const float* dx, float* dF, void* stream) {
  if (!dia_off || !dx || !dF) return MMDFN_ENULL;
  if (N <= 0) return 0;
  mfn_pack_bwd_kernel<<<nb((i64)3 * N * MD), 256, 0, (cudaStream_t)stream>>>(T, B, N, dia_off, p0, p1, p2, dx, dF);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
extern "C" int mmdfn_mfn_unpad_fwd(int T, int B, const int* dia_off, const float* out, float* feat, void* stream) {
  if (!dia_off || !out || !feat) return MMDFN_ENULL;
  if (T <= 0 || B <= 0) return 0;
  mfn_unpad_kernel<<<nb((i64)T * B * MG), 256, 0, (cudaStream_t)stream>>>(T, B, dia_off, out, feat);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
extern "C" int mmdfn_mfn_unpad_bwd(int T, int B, const int* dia_off, const float* dfeat, float* dout, void* stream) {
  if (!dia_off || !dfeat || !dout) return MMDFN_ENULL;
  if (T <= 0 || B <= 0) return 0;
  mfn_unpad_bwd_kernel<<<nb((i64)T * B * MG), 256, 0, (cudaStream_t)stream>>>(T, B, dia_off, dfeat, dout);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
extern "C" int mmdfn_relu_mask_fwd(long long n, const float* x, const unsigned char* mask, float scale, float* y, void* stream) {
  if (!x || !y) return MMDFN_ENULL;
  if (n <= 0) return 0;
  mfn_relu_mask_kernel<<<nb(n), 256, 0, (cudaStream_t)stream>>>(n, x, mask, scale, y);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
/* dx = y != 0 ? dy * ind_scale : 0   (ind_scale = the dropout scale when a mask was applied, else 1) */
extern "C" int mmdfn_relu_mask_bwd(long long n, const float* dy, const float* y, float ind_scale, float* dx, void* stream) {
  if (!dy || !y || !dx) return MMDFN_ENULL;
  if (n <= 0) return 0;
  mfn_relu_bwd_kernel<<<nb(n), 256, 0, (cudaStream_t)stream>>>(n, dy, y, ind_scale, dx);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
