// Shared between the two generations of the fused graph-conv layer kernel (gcn_layer.cu, gcn_layer2.cu).
#pragma once
#include "common.cuh"

namespace mmdfn {

struct GcnLayerArgs {
  int B, N;
  const int* dia_off;
  const i64* blk_off;
  const float* adj_blk;
  const float* adj_diag;
  const float* zin; i64 ldz;          // (3N, 100) rows, row stride ldz floats (multiple of 4)
  const float* wimg;                  // pre-split weight operand of phase B (GLGeo::IMG floats)
  float* t_out; i64 ldt;              // optional: T rows
  // forward epilogue
  const float* r; i64 ldr;            // R rows (h0 Mbot)
  const float* q;                     // optional residual rows (ld 100)
  const unsigned char* mask;          // optional keep mask (3N x 100)
  float scale;
  unsigned char* flags;               // out: relu-and-keep flags (3N x 100)
  // both
  const float* add;                   // backward: optional rows added to the result (ld 100)
  float* out; i64 ldo;
  long long* dbg;
};

// x = hi + lo, hi = x rounded to tf32 (nearest, ties away from zero), lo = the exact fp32 remainder.  The rounding is done
// on the bit pattern (add half an ulp of the 13 dropped bits to the magnitude, clear them): the same values as
// cvt.rna.tf32.f32 for finite inputs in two integer operations -- the conversion instruction expands to five with its
// NaN / infinity handling, and the converter warps of these kernels are issue-bound.
__device__ __forceinline__ void gl_split(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}


// second generation (gcn_layer2.cu): persistent, one CTA per SM, A operands in tensor memory; dialogues of <= 128
// utterances, 16-wide K chunks.  Returns false when the launch does not fit (the caller then uses gcn_layer.cu).
bool gcn_layer2_eligible(int Lmax);
int gcn_layer2_launch(bool fwd, const GcnLayerArgs& a, int Lmax, cudaStream_t st);

}  // namespace mmdfn
