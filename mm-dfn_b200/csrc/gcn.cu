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

// per-layer saved activations, in floats per node row (the fused layer kernel of gcn_layer.cu keeps hi / u on chip)
constexpr int SV_GATES = 0;      // 400: i f g o (activated)
constexpr int SV_C = 400;        // 100
constexpr int SV_H = 500;        // 100
constexpr int SV_Z = 600;        // 100: z_{l+1}
constexpr int SV_FLAGS = 700;    // 25 floats = 100 bytes: [relu and keep] flags of the layer's output
constexpr int SV_ROW = 728;      // rounded so that every region of every layer stays 16-byte aligned for any row count

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

// du[r, 0:100] (ld ldu) = flags ? dz[r, 0:100] (ld ldz) * scale : 0     (through dropout and ReLU of the layer output)
__global__ void gcn_du_kernel(i64 rows, const float* __restrict__ dz, i64 ldz, const unsigned char* __restrict__ flags,
                              float scale, float* __restrict__ du, i64 ldu) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;          // one float4 per thread
  if (idx >= rows * (GG / 4)) return;
  const i64 r = idx / (GG / 4);
  const int c4 = (int)(idx - r * (GG / 4));
  const float4 g = *(reinterpret_cast<const float4*>(dz + r * ldz) + c4);
  const uint32_t f = *(reinterpret_cast<const uint32_t*>(flags + r * GG) + c4);
  *(reinterpret_cast<float4*>(du + r * ldu) + c4) =
      make_float4((f & 0xFFu) ? g.x * scale : 0.f, (f & 0xFF00u) ? g.y * scale : 0.f, (f & 0xFF0000u) ? g.z * scale : 0.f,
                  (f & 0xFF000000u) ? g.w * scale : 0.f);
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

// the same copy, four columns per thread (cols, both leading dimensions multiples of 4; 16-byte aligned pointers)
__global__ void copy2d_mask4_kernel(i64 rows, int cols4, const float* __restrict__ src, i64 sld,
                                    const unsigned char* __restrict__ mask, float scale, float* __restrict__ dst,
                                    i64 dld) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols4) return;
  const i64 r = idx / cols4;
  const int c = 4 * (int)(idx - r * cols4);
  float4 v = *reinterpret_cast<const float4*>(src + r * sld + c);
  if (mask) {
    const uint32_t k = *reinterpret_cast<const uint32_t*>(mask + 4 * idx);
    v.x = (k & 0xFFu) ? v.x * scale : 0.f;
    v.y = (k & 0xFF00u) ? v.y * scale : 0.f;
    v.z = (k & 0xFF0000u) ? v.z * scale : 0.f;
    v.w = (k & 0xFF000000u) ? v.w * scale : 0.f;
  }
  *reinterpret_cast<float4*>(dst + r * dld + c) = v;
}

static int copy2d_mask(i64 rows, int cols, const float* src, i64 sld, const unsigned char* mask, float scale, float* dst,
                       i64 dld, cudaStream_t st) {
  const bool vec = (cols & 3) == 0 && (sld & 3) == 0 && (dld & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(mask) & 3) == 0;
  if (vec) copy2d_mask4_kernel<<<(unsigned)ceil_div64(rows * (cols / 4), 256), 256, 0, st>>>(rows, cols / 4, src, sld, mask, scale, dst, dld);
  else copy2d_mask_kernel<<<(unsigned)ceil_div64(rows * cols, 256), 256, 0, st>>>(rows, cols, src, sld, mask, scale, dst, dld);
  MMDFN_LAUNCH_CHECK();
  return 0;
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

static inline unsigned nblk(i64 n) { return (unsigned)ceil_div64(n, 256); }

}  // namespace mmdfn

using namespace mmdfn;

// workspace layout of the forward (floats; every region starts 16-byte aligned):
//   h0 (n3 x 100) | z0 (n3 x 100) | zeros (n3 x 100) | R_all (n3 x 100 K) | K x [layer: n3 x SV_ROW] | pre (n3 x 400)
//   | Mtop_all (100 x 100 K) | Mbot_all (100 x 100 K) | img_f (K images) | img_b (K images)
struct StackWs {
  i64 h0, z0, zeros, r_all, layers, pre, mtop, mbot, img_f, img_b, total;
};
static StackWs stack_ws(i64 n3, int K) {
  StackWs w;
  i64 o = 0;
  w.h0 = o; o += n3 * GG;
  w.z0 = o; o += n3 * GG;
  w.zeros = o; o += n3 * GG;
  w.r_all = o; o += n3 * GG * K;
  w.layers = o; o += n3 * SV_ROW * K;
  w.pre = o; o += n3 * 4 * GG;
  w.mtop = o; o += (i64)GG * GG * K;
  w.mbot = o; o += (i64)GG * GG * K;
  w.img_f = o; o += gcn_layer_img_floats() * K;
  w.img_b = o; o += gcn_layer_img_floats() * K;
  w.total = o;
  return w;
}

extern "C" long long mmdfn_gcn_stack_ws_floats(int n3, int K) { return stack_ws(n3, K < 0 ? 0 : K).total; }

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
  const StackWs w = stack_ws(n3, K);
  float* h0 = ws + w.h0;
  float* z0 = ws + w.z0;
  float* zeros = ws + w.zeros;
  float* r_all = ws + w.r_all;
  float* layers = ws + w.layers;
  float* pre = ws + w.pre;
  const i64 ldk = (i64)GG * K;
  // x_d = dropout(X) stored straight into F[:, 0:200]                           (model_GCN.py:453,483)
  MMDFN_TRY(copy2d_mask(n3, GX, X, GX, mask_x, mask_scale, F, GF, st));
  // h0 = relu(x_d W0^T + b0)                                                     (:454)
  MMDFN_TRY(gemm(false, true, (int)n3, GG, GX, 1.f, F, GF, W0, GX, 0.f, h0, GG, b0, 1, st));
  MMDFN_TRY(copy2d_mask(n3, GG, h0, GG, mask_h0, mask_scale, z0, GG, st));   // (:456)
  if (K == 0) {
    MMDFN_TRY(copy2d_mask(n3, GG, z0, GG, nullptr, 1.f, F + GX, GF, st));    // (:482-483)
    return 0;
  }
  if (reason_flag) MMDFN_TRY(fill_zero(zeros, (size_t)n3 * GG * sizeof(float), st));
  // folded layer weights (theta / alpha mixes inside the operands) and their pre-split tensor-core images, then the
  // layer-invariant half of every layer in ONE product: R_all = h0 [Mbot_1 | .. | Mbot_K]
  MMDFN_TRY(gcn_layer_prep(K, convW, lamda, alpha, ws + w.mtop, ws + w.mbot, ws + w.img_f, ws + w.img_b, st));
  MMDFN_TRY(gemm(false, false, (int)n3, (int)ldk, GG, 1.f, h0, GG, ws + w.mbot, ldk, 0.f, r_all, ldk, nullptr, 0, st));
  const float* zl = z0;
  const float* hl = zeros;
  const float* cl = zeros;
  for (int l = 0; l < K; l++) {
    float* sv = layers + (i64)l * n3 * SV_ROW;
    float* gates = sv + n3 * SV_GATES;
    float* c = sv + n3 * SV_C;
    float* h = sv + n3 * SV_H;
    float* z = sv + n3 * SV_Z;
    unsigned char* flags = reinterpret_cast<unsigned char*>(sv + n3 * SV_FLAGS);
    const float* agg_in = zl;
    if (reason_flag) {
      MMDFN_TRY(gemm(false, true, (int)n3, 4 * GG, GG, 1.f, zl, GG, w_ih, GG, 0.f, pre, 4 * GG, b_ih, 0, st));
      // layer 0 starts from h = 0: its recurrent product is zero, only b_hh is added (a K = 0 call of the same entry point)
      MMDFN_TRY(gemm(false, true, (int)n3, 4 * GG, l == 0 ? 0 : GG, 1.f, hl, GG, w_hh, GG, 1.f, pre, 4 * GG, b_hh, 0, st));
      lstm_fwd_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3, pre, cl, gates, c, h);
      MMDFN_LAUNCH_CHECK();
      agg_in = h;
    }
    // one launch: aggregate -> x Mtop -> + R -> ReLU -> dropout -> (+ q); the last layer writes F[:, 200:300] directly
    const unsigned char* mk = mask_layers ? mask_layers + (i64)l * n3 * GG : nullptr;
    const bool last = (l == K - 1);
    MMDFN_TRY(gcn_layer_fwd(B, N, Lmax, dia_off, (const i64*)blk_off, adj_blk, adj_diag, agg_in,
                            ws + w.img_f + (i64)l * gcn_layer_img_floats(), r_all + (i64)l * GG, ldk,
                            reason_flag ? zl : nullptr, mk, mask_scale, flags, last ? F + GX : z, last ? GF : GG, st));
    zl = z;
    if (reason_flag) { hl = h; cl = c; }
  }
  return 0;
}

// backward workspace: dz dh0 dhc dcc dh dhi (n3 x 100 each) | dgates (n3 x 400) | dxd (n3 x 200) | DU_all (n3 x 100 K)
// | T_all (n3 x 100 K) | dMtop_all (100 x 100 K) | dMbot_all (100 x 100 K)
extern "C" long long mmdfn_gcn_stack_bwd_ws_floats(int n3, int K) {
  if (K < 0) K = 0;
  return (i64)n3 * (1200 + 2 * (i64)GG * K) + 2 * (i64)GG * GG * K;
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
  const StackWs w = stack_ws(n3, K);
  const float* h0 = ws_fwd + w.h0;
  const float* z0 = ws_fwd + w.z0;
  const float* zeros = ws_fwd + w.zeros;
  const float* layers = ws_fwd + w.layers;
  const float* mtop_all = ws_fwd + w.mtop;
  const float* mbot_all = ws_fwd + w.mbot;
  const float* img_b = ws_fwd + w.img_b;
  const i64 ldk = (i64)GG * K;
  float* dz = ws;
  float* dh0 = dz + n3 * GG;
  float* dhc = dh0 + n3 * GG;
  float* dcc = dhc + n3 * GG;
  float* dh = dcc + n3 * GG;
  float* dhi = dh + n3 * GG;
  float* dgates = dhi + n3 * GG;
  float* dxd = dgates + n3 * 4 * GG;
  float* du_all = dxd + n3 * GX;
  float* t_all = du_all + n3 * ldk;
  float* dmtop_all = t_all + n3 * ldk;
  float* dmbot_all = dmtop_all + (i64)GG * ldk;
  const float scale = mask_layers ? mask_scale : 1.f;
  // dz_K = dF[:, 200:300]
  MMDFN_TRY(copy2d_mask(n3, GG, dF + GX, GF, nullptr, 1.f, dz, GG, st));
  if (K == 0) MMDFN_TRY(fill_zero(dh0, (size_t)n3 * GG * sizeof(float), st));
  else MMDFN_TRY(fill_zero(dmtop_all, (size_t)2 * GG * ldk * sizeof(float), st));      // dMtop_all | dMbot_all: split-K targets
  const float gb = grads_zeroed ? 1.f : 0.f;     // caller pre-zeroed every gradient buffer: accumulate, no zero-init launches
  bool have_carry = false;     // dhc / dcc valid (gradient flowing into h_{l+1}, c_{l+1} from layer l+1)
  bool first_rnn = true;
  for (int l = K - 1; l >= 0; l--) {
    const float* sv = layers + (i64)l * n3 * SV_ROW;
    const float* gates = sv + n3 * SV_GATES;
    const float* c = sv + n3 * SV_C;
    const float* h = sv + n3 * SV_H;
    const unsigned char* flags = reinterpret_cast<const unsigned char*>(sv + n3 * SV_FLAGS);
    const float* zprev = l > 0 ? layers + (i64)(l - 1) * n3 * SV_ROW + n3 * SV_Z : z0;
    const float* hprev = l > 0 ? layers + (i64)(l - 1) * n3 * SV_ROW + n3 * SV_H : zeros;
    const float* cprev = l > 0 ? layers + (i64)(l - 1) * n3 * SV_ROW + n3 * SV_C : zeros;
    const float* agg_in = reason_flag ? h : zprev;
    float* du = du_all + (i64)l * GG;             // column block l of DU_all (row stride 100 K)
    float* tl = t_all + (i64)l * GG;
    // du = dz * [relu and keep] * scale
    gcn_du_kernel<<<nblk(n3 * (GG / 4)), 256, 0, st>>>(n3, dz, GG, flags, scale, du, ldk);
    MMDFN_LAUNCH_CHECK();
    // one launch: t = A_hat du (kept: dMtop = agg_in^T t) and d agg_in = t Mtop^T (+ the gradient carried into h from
    // layer l+1's recurrent product)
    MMDFN_TRY(gcn_layer_bwd(B, N, Lmax, dia_off, (const i64*)blk_off, adj_blk, adj_diag, du, ldk,
                            img_b + (i64)l * gcn_layer_img_floats(), tl, ldk,
                            (reason_flag && have_carry) ? dhc : nullptr, reason_flag ? dh : dz, st));
    MMDFN_TRY(gemm(true, false, GG, GG, (int)n3, 1.f, agg_in, GG, tl, ldk, 1.f, dmtop_all + (i64)l * GG, ldk, nullptr, 0, st));
    if (d_adj_blk) {
      // dhi = du Mtop^T, dA_hat += sym(dhi agg_in^T)
      MMDFN_TRY(gemm(false, true, (int)n3, GG, GG, 1.f, du, ldk, mtop_all + (i64)l * GG, ldk, 0.f, dhi, GG, nullptr, 0, st));
      MMDFN_TRY(adj_grad_accum(B, N, Lmax, dia_off, (const i64*)blk_off, dhi, agg_in, GG, d_adj_blk, d_adj_diag,
                               l != K - 1, st));
    }
    if (!reason_flag) continue;                  // dz now holds dz_l = A_hat-path gradient only (no residual, no gate)
    lstm_bwd_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3, dh, have_carry ? dcc : nullptr, gates, c, cprev, dgates, dcc);
    MMDFN_LAUNCH_CHECK();
    // dz_l = dz_{l+1} (residual +q) + dgates W_ih ; dhc = dgates W_hh
    const float beta = (first_rnn && !grads_zeroed) ? 0.f : 1.f;
    if (l > 0) {
      // input and recurrent halves share dgates: one launch each for the two data gradients and the two weight gradients
      MMDFN_TRY(gemm_npair(false, (int)n3, GG, 4 * GG, dgates, 4 * GG, w_ih, w_hh, GG, 1.f, dz, 0.f, dhc, GG, st));
      MMDFN_TRY(gemm_npair(true, 4 * GG, GG, (int)n3, dgates, 4 * GG, zprev, hprev, GG, beta, dw_ih, beta, dw_hh, GG, st));
    } else {
      MMDFN_TRY(gemm(false, false, (int)n3, GG, 4 * GG, 1.f, dgates, 4 * GG, w_ih, GG, 1.f, dz, GG, nullptr, 0, st));
      MMDFN_TRY(gemm(true, false, 4 * GG, GG, (int)n3, 1.f, dgates, 4 * GG, zprev, GG, beta, dw_ih, GG, nullptr, 0, st));
      if (beta == 0.f) MMDFN_TRY(fill_zero(dw_hh, (size_t)4 * GG * GG * sizeof(float), st));      // h_{-1} = 0: no contribution from layer 0
    }
    MMDFN_TRY(colsum((int)n3, 4 * GG, dgates, 4 * GG, beta, db_ih, st));
    first_rnn = false;
    have_carry = true;
  }
  if (K > 0) {
    // the h0 half of every layer at once: dh0 = DU_all Mbot_all^T ; dMbot_all = h0^T DU_all ; then dW_l = theta_l [dMtop_l ; dMbot_l]
    MMDFN_TRY(gemm(false, true, (int)n3, GG, (int)ldk, 1.f, du_all, ldk, mbot_all, ldk, 0.f, dh0, GG, nullptr, 0, st));
    MMDFN_TRY(gemm(true, false, GG, (int)ldk, (int)n3, 1.f, h0, GG, du_all, ldk, 1.f, dmbot_all, ldk, nullptr, 0, st));
    MMDFN_TRY(gcn_layer_unfold(K, dconvW, lamda, dmtop_all, dmbot_all, grads_zeroed, st));
  }
  if (reason_flag && db_hh && K > 0) {
    MMDFN_CUDA(cudaMemcpyAsync(db_hh, db_ih, 4 * GG * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  // through z0 = dropout(h0), h0 = relu(x_d W0^T + b0)
  float* dpre = dhi;
  h0_bwd_kernel<<<nblk(n3 * GG), 256, 0, st>>>(n3 * GG, dh0, dz, mask_h0, mask_scale, h0, dpre);
  MMDFN_LAUNCH_CHECK();
  MMDFN_TRY(gemm(true, false, GG, GX, (int)n3, 1.f, dpre, GG, F, GF, gb, dW0, GX, nullptr, 0, st));
  MMDFN_TRY(colsum((int)n3, GG, dpre, GG, gb, db0, st));
  MMDFN_TRY(gemm(false, false, (int)n3, GX, GG, 1.f, dpre, GG, W0, GX, 0.f, dxd, GX, nullptr, 0, st));
  x_bwd_kernel<<<nblk(n3 * GX), 256, 0, st>>>(n3, dF, dxd, mask_x, mask_scale, dX);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
