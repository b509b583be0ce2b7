// k3/k4: speaker-party partition (integer, bit-exact), the fused scatter + speaker-weight
// combine + ragged pack that writes the stacked graph input [a; v; l], and its backward.
// Replaces the Python double loops with nonzero() syncs at code/model.py:1070-1090,
// 1101-1121, 1134-1154 and simple_batch_graphify (code/model.py:553-565).
#include "internal.cuh"
#include "../../include/mmdfn_b200.h"

namespace mmdfn {

// one warp per (dialogue b, speaker p): ranks are assigned 32 time steps at a time with a ballot + prefix popcount,
// which keeps the ascending order of torch.nonzero (stable partition) without the 100 dependent loads per thread of a
// sequential scan (that version was a 55 us single-CTA kernel at the head of the party encoder's critical path).
__global__ void spk_partition_kernel(int T, int B, int S, const float* __restrict__ qmask, int* __restrict__ pos,
                                     int* __restrict__ cnt, int* __restrict__ rowmap) {
  const int id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (id >= B * S) return;                      // warp-uniform
  const int b = id / S, p = id - b * S;
  const i64 nseq = (i64)3 * B * S;
  int k = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const i64 e = ((i64)t * B + b) * S + p;
    const bool on = (t < T) && (qmask[e] != 0.0f);
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (t < T) {
      if (on) {
        const int r = k + __popc(m & ((1u << lane) - 1u));
        pos[e] = r;
        if (rowmap)
          for (int mm = 0; mm < 3; mm++) rowmap[(i64)r * nseq + ((i64)mm * B + b) * S + p] = (mm * T + t) * B + b;
      } else {
        pos[e] = -1;
      }
    }
    k += __popc(m);
  }
  if (lane == 0) cnt[id] = k;
  if (rowmap)
    for (int kk = k + lane; kk < T; kk += 32)
      for (int mm = 0; mm < 3; mm++) rowmap[(i64)kk * nseq + ((i64)mm * B + b) * S + p] = -1;
}

// sel[t,b] = last speaker p with qmask != 0 (assignment order of code/model.py:1084-1088), -1 if none
__global__ void spk_select_kernel(int T, int B, int S, const float* __restrict__ qmask, int* __restrict__ sel) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= T * B) return;
  int s = -1;
  for (int p = 0; p < S; p++)
    if (qmask[(i64)id * S + p] != 0.0f) s = p;
  sel[id] = s;
}

struct PackArgs {
  int T, B, S, N;
  const int* dia_off;
  const int* sel;
  const int* pos;
  const float* base[3];   // (T,B,200) each
  const float* Q;         // (T, 3*B*S, 200) party encoder output, nullable
  float w[3];
  float* X;               // (3N, 200)
};

// warp per (modality, dialogue, t)
__global__ void party_pack_fwd_kernel(PackArgs p) {
  const int lane = threadIdx.x;
  const int t = blockIdx.x * blockDim.y + threadIdx.y;
  const int b = blockIdx.y, m = blockIdx.z;
  const int off = p.dia_off[b], L = p.dia_off[b + 1] - off;
  if (t >= L) return;
  const i64 tb = (i64)t * p.B + b;
  const float4* base = reinterpret_cast<const float4*>(p.base[m] + tb * 200);
  const int s = p.Q ? p.sel[tb] : -1;
  const float4* q = nullptr;
  if (s >= 0) {
    const int k = p.pos[tb * p.S + s];
    const i64 nseq = (i64)3 * p.B * p.S;
    q = reinterpret_cast<const float4*>(p.Q + ((i64)k * nseq + ((i64)m * p.B + b) * p.S + s) * 200);
  }
  float4* out = reinterpret_cast<float4*>(p.X + ((i64)m * p.N + off + t) * 200);
  const float w = p.w[m];
  for (int c = lane; c < 50; c += 32) {
    float4 v = base[c];
    if (q) {
      const float4 u = q[c];
      v.x = fmaf(w, u.x, v.x); v.y = fmaf(w, u.y, v.y); v.z = fmaf(w, u.z, v.z); v.w = fmaf(w, u.w, v.w);
    }
    out[c] = v;
  }
}

struct PackBwdArgs {
  int T, B, S, N;
  const int* dia_off;
  const int* sel;
  const int* pos;
  const float* dX;        // (3N, 200)
  float w[3];
  float* dbase[3];        // (T,B,200): written for every (t,b): dX row for t < L, 0 on padding
  float* dQ;              // (T, 3*B*S, 200), pre-zeroed; nullable
};

__global__ void party_pack_bwd_kernel(PackBwdArgs p) {
  const int lane = threadIdx.x;
  const int t = blockIdx.x * blockDim.y + threadIdx.y;
  const int b = blockIdx.y, m = blockIdx.z;
  if (t >= p.T) return;
  const int off = p.dia_off[b], L = p.dia_off[b + 1] - off;
  const i64 tb = (i64)t * p.B + b;
  float4* db = reinterpret_cast<float4*>(p.dbase[m] + tb * 200);
  if (t >= L) {
    for (int c = lane; c < 50; c += 32) db[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float4* g = reinterpret_cast<const float4*>(p.dX + ((i64)m * p.N + off + t) * 200);
  const int s = p.dQ ? p.sel[tb] : -1;
  float4* dq = nullptr;
  if (s >= 0) {
    const int k = p.pos[tb * p.S + s];
    const i64 nseq = (i64)3 * p.B * p.S;
    dq = reinterpret_cast<float4*>(p.dQ + ((i64)k * nseq + ((i64)m * p.B + b) * p.S + s) * 200);
  }
  const float w = p.w[m];
  for (int c = lane; c < 50; c += 32) {
    const float4 v = g[c];
    db[c] = v;
    if (dq) dq[c] = make_float4(w * v.x, w * v.y, w * v.z, w * v.w);
  }
}

// padded (T,B,D) view of one modality: rows t < L_b come from the packed graph input, the padded tail from `pad_src`
// (relation graph type: MaskedEdgeAttention soft-maxes over all T rows of the padded encoder output, code/model.py:449)
__global__ void unpack_pad_kernel(int T, int B, int D, const int* __restrict__ dia_off, const float* __restrict__ packed,
                                  const float* __restrict__ pad_src, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.y + threadIdx.y;
  const int b = blockIdx.y;
  if (t >= T) return;
  const int off = dia_off[b], L = dia_off[b + 1] - off;
  const float* src = t < L ? packed + (i64)(off + t) * D : pad_src + ((i64)t * B + b) * D;
  float* dst = out + ((i64)t * B + b) * D;
  for (int c = threadIdx.x; c < D; c += 32) dst[c] = src[c];
}

// adjoint: d_packed[off+t] = dM[t,b] (t < L) ; d_pad[t,b] = t < L ? 0 : dM[t,b]
__global__ void unpack_pad_bwd_kernel(int T, int B, int D, const int* __restrict__ dia_off, const float* __restrict__ dM,
                                      float* __restrict__ d_packed, float* __restrict__ d_pad) {
  const int t = blockIdx.x * blockDim.y + threadIdx.y;
  const int b = blockIdx.y;
  if (t >= T) return;
  const int off = dia_off[b], L = dia_off[b + 1] - off;
  const float* g = dM + ((i64)t * B + b) * D;
  float* dp = d_pad + ((i64)t * B + b) * D;
  if (t < L) {
    float* dk = d_packed + (i64)(off + t) * D;
    for (int c = threadIdx.x; c < D; c += 32) { dk[c] = g[c]; dp[c] = 0.f; }
  } else {
    for (int c = threadIdx.x; c < D; c += 32) dp[c] = g[c];
  }
}

}  // namespace mmdfn

using namespace mmdfn;

extern "C" int mmdfn_spk_partition(int T, int B, int S, const float* qmask, int* pos, int* cnt, int* sel,
                                   int* rowmap, void* stream) {
  if (!qmask || !pos || !cnt || !sel) return MMDFN_ENULL;
  if (T < 0 || B < 0 || S <= 0) return MMDFN_EINVAL;
  if ((i64)3 * T * B > 2000000000LL) return MMDFN_ERANGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return 0;
  spk_partition_kernel<<<ceil_div(B * S, 4), 128, 0, st>>>(T, B, S, qmask, pos, cnt, rowmap);
  MMDFN_LAUNCH_CHECK();
  if (T > 0) {
    spk_select_kernel<<<ceil_div(T * B, 256), 256, 0, st>>>(T, B, S, qmask, sel);
    MMDFN_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int mmdfn_party_pack_fwd(int T, int B, int S, int N, const int* dia_off, const int* sel, const int* pos,
                                    const float* base_a, const float* base_v, const float* base_l, const float* Q,
                                    float wa, float wv, float wl, float* X, void* stream) {
  if (!dia_off || !base_a || !base_v || !base_l || !X) return MMDFN_ENULL;
  if (Q && (!sel || !pos)) return MMDFN_ENULL;
  if (T <= 0 || B <= 0 || N <= 0) return 0;
  PackArgs a{T, B, S, N, dia_off, sel, pos, {base_a, base_v, base_l}, Q, {wa, wv, wl}, X};
  party_pack_fwd_kernel<<<dim3(ceil_div(T, 8), B, 3), dim3(32, 8), 0, (cudaStream_t)stream>>>(a);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_party_pack_bwd(int T, int B, int S, int N, const int* dia_off, const int* sel, const int* pos,
                                    const float* dX, float wa, float wv, float wl, float* dbase_a, float* dbase_v,
                                    float* dbase_l, float* dQ, void* stream) {
  if (!dia_off || !dX || !dbase_a || !dbase_v || !dbase_l) return MMDFN_ENULL;
  if (dQ && (!sel || !pos)) return MMDFN_ENULL;
  if (T <= 0 || B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dQ) MMDFN_TRY(fill_zero(dQ, (size_t)T * 3 * B * S * 200 * sizeof(float), st));
  PackBwdArgs a{T, B, S, N, dia_off, sel, pos, dX, {wa, wv, wl}, {dbase_a, dbase_v, dbase_l}, dQ};
  party_pack_bwd_kernel<<<dim3(ceil_div(T, 8), B, 3), dim3(32, 8), 0, st>>>(a);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_unpack_pad_fwd(int T, int B, int D, const int* dia_off, const float* packed, const float* pad_src,
                                    float* out, void* stream) {
  if (!dia_off || !packed || !pad_src || !out) return MMDFN_ENULL;
  if (T <= 0 || B <= 0) return 0;
  unpack_pad_kernel<<<dim3(ceil_div(T, 8), B), dim3(32, 8), 0, (cudaStream_t)stream>>>(T, B, D, dia_off, packed, pad_src, out);
  MMDFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmdfn_unpack_pad_bwd(int T, int B, int D, const int* dia_off, const float* dM, float* d_packed,
                                    float* d_pad, void* stream) {
  if (!dia_off || !dM || !d_packed || !d_pad) return MMDFN_ENULL;
  if (T <= 0 || B <= 0) return 0;
  unpack_pad_bwd_kernel<<<dim3(ceil_div(T, 8), B), dim3(32, 8), 0, (cudaStream_t)stream>>>(T, B, D, dia_off, dM, d_packed, d_pad);
  MMDFN_LAUNCH_CHECK();
  return 0;
}
