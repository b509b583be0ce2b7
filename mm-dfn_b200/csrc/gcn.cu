// k6/k7/k8: GCNII_lyc stack -- fcs[0]+ReLU, then K x [1-step LSTM gate -> GraphConvolution
// (message aggregate on the block-compact adjacency, [hi|h0]W, theta/alpha mixes) -> ReLU
// -> dropout -> +q], forward and backward.  Replaces code/model_GCN.py:444-488 (stack),
// :176-189 (GraphConvolution.forward), :432-434,466 (nn.LSTM single step).
#include "internal.cuh"
#include "../../include/mmdfn_b200.h"
#include <math.h>

namespace mmdfn {

constexpr int GG = 100;   // graph hidden size (graph_h, code/run_train_erc.py:391)
constexpr int GX = 200;   // graph input width
constexpr int GF = 300;   // output row: [x (200) | z_K (100)]

// per-layer saved activations, in floats per node row
constexpr int SV_GATES = 0;      // 400: i f g o (activated)
constexpr int SV_C = 400;        // 100
constexpr int SV_H = 500;        // 100
constexpr int SV_HI = 600;       // 100
constexpr int SV_RD = 700;       // 100: dropout(relu(u))
constexpr int SV_Z = 800;        // 100: z_{l+1}
constexpr int SV_ROW = 900;

__global__ void lstm_fwd_kernel(i64 n, const float* __restrict__ pre, const float* __restrict__ c_prev,
                                float* __restrict__ gates, float* __restrict__ c, float* __restrict__ h) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * GG) return;
  const i64 r = idx / GG;
  const int u = (int)(idx - r * GG);
  const float* p = pre + r * 4 * GG;
  const float i_ = sigmoidf_(p[u]), f_ = sigmoidf_(p[GG + u]), g_ = tanhf(p[2 * GG + u]), o_ = sigmoidf_(p[3 * GG + u]);
  const float cn = f_ * c_prev[idx] + i_ * g_;
  float* g = gates + r * 4 * GG;
  g[u] = i_; g[GG + u] = f_; g[2 * GG + u] = g_; g[3 * GG + u] = o_;
  c[idx] = cn;
  h[idx] = o_ * tanhf(cn);
}

// dc_in may alias dc_out (each element is read, then written, by the same thread)
__global__ void lstm_bwd_kernel(i64 n, const float* __restrict__ dh, const float* dc_in,
                                const float* __restrict__ gates, const float* __restrict__ c,
                                const float* __restrict__ c_prev, float* __restrict__ dgates, float* dc_out) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * GG) return;
  const i64 r = idx / GG;
  const int u = (int)(idx - r * GG);
  const float* g = gates + r * 4 * GG;
  const float i_ = g[u], f_ = g[GG + u], g_ = g[2 * GG + u], o_ = g[3 * GG + u];
  const float tc = tanhf(c[idx]);
  const float dh_ = dh[idx];
  const float dc = (dc_in ? dc_in[idx] : 0.f) + dh_ * o_ * (1.0f - tc * tc);
  float* d = dgates + r * 4 * GG;
  d[u] = dc * g_ * i_ * (1.0f - i_);
  d[GG + u] = dc * c_prev[idx] * f_ * (1.0f - f_);
  d[2 * GG + u] = dc * i_ * (1.0f - g_ * g_);
  d[3 * GG + u] = dh_ * tc * o_ * (1.0f - o_);
  dc_out[idx] = dc * f_;
}

// u = theta*u1 + (1-theta)*((1-alpha)*hi + alpha*h0); rd = dropout(relu(u)); z = rd (+ q)
__global__ void gcn_epi_fwd_kernel(i64 n, const float* __restrict__ u1, const float* __restrict__ hi,
                                   const float* __restrict__ h0, const float* __restrict__ q,
                                   const unsigned char* __restrict__ mask, float scale, float theta, float omt,
                                   float alpha, float oma, float* __restrict__ rd, float* __restrict__ z) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float r = oma * hi[idx] + alpha * h0[idx];
  const float u = theta * u1[idx] + omt * r;
  float v = fmaxf(u, 0.f);
  if (mask) v = mask[idx] ? v * scale : 0.f;
  rd[idx] = v;
  z[idx] = q ? v + q[idx] : v;
}

// du = dz * [rd > 0] * scale ; du1 = theta*du ; dhi = (1-theta)(1-alpha) du ; dh0 += (1-theta) alpha du
__global__ void gcn_epi_bwd_kernel(i64 n, const float* __restrict__ dz, const float* __restrict__ rd, float scale,
                                   float theta, float omt, float alpha, float oma, float* __restrict__ du1,
                                   float* __restrict__ dhi, float* __restrict__ dh0) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float du = rd[idx] > 0.f ? dz[idx] * scale : 0.f;
  du1[idx] = theta * du;
  dhi[idx] = omt * oma * du;
  dh0[idx] += omt * alpha * du;
}

// dst[r, 0:cols] (ld dld) = src[r, 0:cols] (ld sld) [* mask*scale]
__global__ void copy2d_mask_kernel(i64 rows, int cols, const float* __restrict__ src, i64 sld,
                                   const unsigned char* __restrict__ mask, float scale, float* __restrict__ dst,
                                   i64 dld) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const i64 r = idx / cols;
  const int c = (int)(idx - r * cols);
  float v = src[r * sld + c];
  if (mask) v = mask[idx] ? v * scale : 0.f;
  dst[r * dld + c] = v;
}

// dpre = (dh0 + dz0*mask*scale) * [h0 > 0]
__global__ void h0_bwd_kernel(i64 n, const float* __restrict__ dh0, const float* __restrict__ dz0,
                              const unsigned char* __restrict__ mask, float scale, const float* __restrict__ h0,
                              float* __restrict__ dpre) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float g = dz0[idx];
  if (mask) g = mask[idx] ? g * scale : 0.f;
  g += dh0[idx];
  dpre[idx] = h0[idx] > 0.f ? g : 0.f;
}

// dX[r,c] = (dF[r,c] + dxd[r,c]) * mask*scale
__global__ void x_bwd_kernel(i64 rows, const float* __restrict__ dF, const float* __restrict__ dxd,
                             const unsigned char* __restrict__ mask, float scale, float* __restrict__ dX) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * GX) return;
  const i64 r = idx / GX;
  const int c = (int)(idx - r * GX);
  float v = dF[r * GF + c] + dxd[idx];
  if (mask) v = mask[idx] ? v * scale : 0.f;
  dX[idx] = v;
}

// y += x
__global__ void axpy_kernel(i64 n, const float* __restrict__ x, float* __restrict__ y) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) y[idx] += x[idx];
}

static inline unsigned nblk(i64 n) { return (unsigned)ceil_div64(n, 256); }

}  // namespace mmdfn

using namespace mmdfn;

extern "C" long long mmdfn_gcn_stack_ws_floats(int n3, int K) {
  // h0 (100) | z0 (100) | zeros (100) | K x SV_ROW | scratch: pre (400) + u1 (100)
  return (i64)n3 * (300 + (i64)K * SV_ROW + 500);
}

extern "C" int mmdfn_gcn_stack_fwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off,
                                   const float* adj_blk, const float* adj_diag, const float* X, int K,
                                   int reason_flag, double lamda, double alpha, const float* W0, const float* b0,
                                   const float* const* convW, const float* w_ih, const float* w_hh,
                                   const float* b_ih, const float* b_hh, const unsigned char* mask_x,
                                   const unsigned char* mask_h0, const unsigned char* mask_layers, float mask_scale,
                                   float* F, float* ws, void* stream) {
  if (!dia_off || !blk_off || !adj_blk || !adj_diag || !X || !W0 || !b0 || !F || !ws) return MMDFN_ENULL;
  if (K > 0 && !convW) return MMDFN_ENULL;
  if (reason_flag && K > 0 && (!w_ih || !w_hh || !b_ih || !b_hh)) return MMDFN_ENULL;
  if (K < 0 || N < 0) return MMDFN_EINVAL;
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const i64 n3 = (i64)3 * N;
  float* h0 = ws;
  float* z0 = h0 + n3 * GG;
  float* zeros = z0 + n3 * GG;
  float* layers = zeros + n3 * GG;
  float* pre = layers + (i64)K * n3 * SV_ROW;
  float* u1 = pre + n3 * 4 * GG;
  // x_d = dropout(X) stored straight into F[:, 0:200]                           (model_GCN.py:453,483)
  copy2d_mask_kernel<<<nblk(n3 * GX), 256, 0, st>>>(n3, GX, X, GX, mask_x, mask_scale, F, GF);
  MMDFN_LAUNCH_CHECK();
  // h0 = relu(x_d W0^T + b0)                                                     (:454)
  MMDFN_TRY(gemm(false, true, (int)n3, GG, GX, 1.f, F, GF, W0, GX, 0.f, h0, GG, b0, 1, st));
  copy2d_mask_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3, GG, h0, GG, mask_h0, mask_scale, z0, GG);   // (:456)
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(fill_zero(zeros, (size_t)n3 * GG * sizeof(float), st));
  const float* zl = z0;
  const float* hl = zeros;
  const float* cl = zeros;
  for (int l = 0; l < K; l++) {
    float* sv = layers + (i64)l * n3 * SV_ROW;
    float* gates = sv + n3 * SV_GATES;
    float* c = sv + n3 * SV_C;
    float* h = sv + n3 * SV_H;
    float* hi = sv + n3 * SV_HI;
    float* rd = sv + n3 * SV_RD;
    float* z = sv + n3 * SV_Z;
    const float* agg_in = zl;
    if (reason_flag) {
      MMDFN_TRY(gemm(false, true, (int)n3, 4 * GG, GG, 1.f, zl, GG, w_ih, GG, 0.f, pre, 4 * GG, b_ih, 0, st));
      // layer 0 starts from h = 0: its recurrent product is zero, only b_hh is added (a K = 0 call of the same entry point)
      MMDFN_TRY(gemm(false, true, (int)n3, 4 * GG, l == 0 ? 0 : GG, 1.f, hl, GG, w_hh, GG, 1.f, pre, 4 * GG, b_hh, 0, st));
      lstm_fwd_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3, pre, cl, gates, c, h);
      MMDFN_LAUNCH_CHECK();
      agg_in = h;
    }
    MMDFN_TRY(adj_spmm(B, N, Lmax, dia_off, (const i64*)blk_off, adj_blk, adj_diag, agg_in, GG, hi, st));
    MMDFN_TRY(gemm(false, false, (int)n3, GG, GG, 1.f, hi, GG, convW[l], GG, 0.f, u1, GG, nullptr, 0, st));
    MMDFN_TRY(gemm(false, false, (int)n3, GG, GG, 1.f, h0, GG, convW[l] + GG * GG, GG, 1.f, u1, GG, nullptr, 0, st));
    const double theta_d = log(lamda / (double)(l + 1) + 1.0);      // python float math (model_GCN.py:177)
    const float theta = (float)theta_d, omt = (float)(1.0 - theta_d);
    const unsigned char* mk = mask_layers ? mask_layers + (i64)l * n3 * GG : nullptr;
    gcn_epi_fwd_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3 * GG, u1, hi, h0, reason_flag ? zl : nullptr, mk, mask_scale,
                                                      theta, omt, (float)alpha, (float)(1.0 - alpha), rd, z);
    MMDFN_LAUNCH_CHECK();
    zl = z;
    if (reason_flag) { hl = h; cl = c; }
  }
  copy2d_mask_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3, GG, zl, GG, nullptr, 1.f, F + GX, GF);       // (:482-483)
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" long long mmdfn_gcn_stack_bwd_ws_floats(int n3) {
  // dz (100) dh0 (100) dhc (100) dcc (100) du1 (100) dhi (100) dh (100) dgates (400) dxd (200)
  return (i64)n3 * 1300;
}

// dW pointers receive "=" (not "+="); d_adj_blk/d_adj_diag (nullable pair) receive the true
// gradient w.r.t. the stored adjacency entries, summed over layers.
extern "C" int mmdfn_gcn_stack_bwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off,
                                   const float* adj_blk, const float* adj_diag, int K, int reason_flag, double lamda,
                                   double alpha, const float* W0, const float* const* convW, const float* w_ih,
                                   const float* w_hh, const unsigned char* mask_x, const unsigned char* mask_h0,
                                   const unsigned char* mask_layers, float mask_scale, const float* F,
                                   const float* ws_fwd, const float* dF, float* dX, float* d_adj_blk,
                                   float* d_adj_diag, float* dW0, float* db0, float* const* dconvW, float* dw_ih,
                                   float* dw_hh, float* db_ih, float* db_hh, int grads_zeroed, float* ws,
                                   void* stream) {
  if (!dia_off || !blk_off || !adj_blk || !adj_diag || !W0 || !F || !ws_fwd || !dF || !dX || !dW0 || !db0 || !ws)
    return MMDFN_ENULL;
  if (K > 0 && (!convW || !dconvW)) return MMDFN_ENULL;
  if (reason_flag && K > 0 && (!w_ih || !w_hh || !dw_ih || !dw_hh || !db_ih || !db_hh)) return MMDFN_ENULL;
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const i64 n3 = (i64)3 * N;
  const float* h0 = ws_fwd;
  const float* z0 = h0 + n3 * GG;
  const float* zeros = z0 + n3 * GG;
  const float* layers = zeros + n3 * GG;
  float* dz = ws;
  float* dh0 = dz + n3 * GG;
  float* dhc = dh0 + n3 * GG;
  float* dcc = dhc + n3 * GG;
  float* du1 = dcc + n3 * GG;
  float* dhi = du1 + n3 * GG;
  float* dh = dhi + n3 * GG;
  float* dgates = dh + n3 * GG;
  float* dxd = dgates + n3 * 4 * GG;
  const float scale = mask_layers ? mask_scale : 1.f;
  // dz_K = dF[:, 200:300]
  copy2d_mask_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3, GG, dF + GX, GF, nullptr, 1.f, dz, GG);
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(fill_zero(dh0, (size_t)n3 * GG * sizeof(float), st));
  const float gb = grads_zeroed ? 1.f : 0.f;     // caller pre-zeroed every gradient buffer: accumulate, no zero-init launches
  if (reason_flag && K == 0 && !grads_zeroed) {
    MMDFN_TRY(fill_zero(dw_ih, 4 * GG * GG * sizeof(float), st));
    MMDFN_TRY(fill_zero(dw_hh, 4 * GG * GG * sizeof(float), st));
    MMDFN_TRY(fill_zero(db_ih, 4 * GG * sizeof(float), st));
  }
  bool have_carry = false;     // dhc / dcc valid (gradient flowing into h_{l+1}, c_{l+1} from layer l+1)
  bool first_rnn = true;
  for (int l = K - 1; l >= 0; l--) {
    const float* sv = layers + (i64)l * n3 * SV_ROW;
    const float* gates = sv + n3 * SV_GATES;
    const float* c = sv + n3 * SV_C;
    const float* h = sv + n3 * SV_H;
    const float* hi = sv + n3 * SV_HI;
    const float* rd = sv + n3 * SV_RD;
    const float* zprev = l > 0 ? layers + (i64)(l - 1) * n3 * SV_ROW + n3 * SV_Z : z0;
    const float* hprev = l > 0 ? layers + (i64)(l - 1) * n3 * SV_ROW + n3 * SV_H : zeros;
    const float* cprev = l > 0 ? layers + (i64)(l - 1) * n3 * SV_ROW + n3 * SV_C : zeros;
    const double theta_d = log(lamda / (double)(l + 1) + 1.0);
    const float theta = (float)theta_d, omt = (float)(1.0 - theta_d);
    gcn_epi_bwd_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3 * GG, dz, rd, scale, theta, omt, (float)alpha,
                                                      (float)(1.0 - alpha), du1, dhi, dh0);
    MMDFN_LAUNCH_CHECK();
    // dhi += du1 Wtop^T ; dh0 += du1 Wbot^T ; dW = [hi|h0]^T du1
    MMDFN_TRY(gemm(false, true, (int)n3, GG, GG, 1.f, du1, GG, convW[l], GG, 1.f, dhi, GG, nullptr, 0, st));
    MMDFN_TRY(gemm(false, true, (int)n3, GG, GG, 1.f, du1, GG, convW[l] + GG * GG, GG, 1.f, dh0, GG, nullptr, 0, st));
    MMDFN_TRY(gemm(true, false, GG, GG, (int)n3, 1.f, hi, GG, du1, GG, gb, dconvW[l], GG, nullptr, 0, st));
    MMDFN_TRY(gemm(true, false, GG, GG, (int)n3, 1.f, h0, GG, du1, GG, gb, dconvW[l] + GG * GG, GG, nullptr, 0, st));
    const float* agg_in = reason_flag ? h : zprev;
    if (d_adj_blk)
      MMDFN_TRY(adj_grad_accum(B, N, Lmax, dia_off, (const i64*)blk_off, dhi, agg_in, GG, d_adj_blk, d_adj_diag,
                               l != K - 1, st));
    if (!reason_flag) {
      // z_l -> (aggregate) only: dz_{l} = A_hat dhi
      MMDFN_TRY(adj_spmm(B, N, Lmax, dia_off, (const i64*)blk_off, adj_blk, adj_diag, dhi, GG, dz, st));
      continue;
    }
    MMDFN_TRY(adj_spmm(B, N, Lmax, dia_off, (const i64*)blk_off, adj_blk, adj_diag, dhi, GG, dh, st));
    if (have_carry) {
      // dh += dhc  (gradient into h_{l+1} from layer l+1's recurrent GEMM)
      axpy_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3 * GG, dhc, dh);
      MMDFN_LAUNCH_CHECK();
    }
    lstm_bwd_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3, dh, have_carry ? dcc : nullptr, gates, c, cprev, dgates, dcc);
    MMDFN_LAUNCH_CHECK();
    // dz_l = dz_{l+1} (residual +q) + dgates W_ih ; dhc = dgates W_hh
    MMDFN_TRY(gemm(false, false, (int)n3, GG, 4 * GG, 1.f, dgates, 4 * GG, w_ih, GG, 1.f, dz, GG, nullptr, 0, st));
    if (l > 0) MMDFN_TRY(gemm(false, false, (int)n3, GG, 4 * GG, 1.f, dgates, 4 * GG, w_hh, GG, 0.f, dhc, GG, nullptr, 0, st));
    const float beta = (first_rnn && !grads_zeroed) ? 0.f : 1.f;
    MMDFN_TRY(gemm(true, false, 4 * GG, GG, (int)n3, 1.f, dgates, 4 * GG, zprev, GG, beta, dw_ih, GG, nullptr, 0, st));
    if (l > 0) {
      MMDFN_TRY(gemm(true, false, 4 * GG, GG, (int)n3, 1.f, dgates, 4 * GG, hprev, GG, beta, dw_hh, GG, nullptr, 0, st));
    } else if (beta == 0.f) {
      MMDFN_TRY(fill_zero(dw_hh, (size_t)4 * GG * GG * sizeof(float), st));      // h_{-1} = 0: no contribution from layer 0
    }
    MMDFN_TRY(colsum((int)n3, 4 * GG, dgates, 4 * GG, beta, db_ih, st));
    first_rnn = false;
    have_carry = true;
  }
  if (reason_flag && db_hh && K > 0) {
    MMDFN_CUDA(cudaMemcpyAsync(db_hh, db_ih, 4 * GG * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  // through z0 = dropout(h0), h0 = relu(x_d W0^T + b0)
  float* dpre = du1;
  h0_bwd_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3 * GG, dh0, dz, mask_h0, mask_scale, h0, dpre);
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(gemm(true, false, GG, GX, (int)n3, 1.f, dpre, GG, F, GF, gb, dW0, GX, nullptr, 0, st));
  MMDFN_TRY(colsum((int)n3, GG, dpre, GG, gb, db0, st));
  MMDFN_TRY(gemm(false, false, (int)n3, GX, GG, 1.f, dpre, GG, W0, GX, 0.f, dxd, GX, nullptr, 0, st));
  x_bwd_kernel<<<nblk(n3 * GX), 256, 0, st>>>(n3, dF, dxd, mask_x, mask_scale, dX);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
