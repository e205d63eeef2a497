// GroupNorm statistics of one (image, group) from the per-slot partial sums the producing convolutions left behind
// (kernels.cuh GnIn): shared by the consumers that evaluate them in their own prologue - the halo convolution's
// transform warps (conv_halo.cu) and the stand-alone apply kernels (norm.cu).
#pragma once
#include "kernels.cuh"

namespace hsidm {

struct GnFoldP {
  const float* part0;   // [N][slots0][C0][2] (sum, sumsq)
  const float* part1;   // [N][slots1][C1][2], source concatenated behind source 0 (C1 may be 0)
  int slots0, slots1, C0, C1;
  int cpg;              // channels per group over C0 + C1 (even)
  float inv_cnt;        // 1 / (cpg * H * W)
  float eps;
};

// (mean, rstd) of group g of image n.  The group's channel pairs x all slots of their source are read as 16-byte loads,
// eight in flight, and added in (channel pair, slot) order in float64: the slots are short fp32 sums, and
// E[x^2] - mean^2 keeps its digits for channels whose mean dominates their spread.  Deterministic.
// S > 1 (a power of two <= 32): S ADJACENT lanes share the group - lane `sub` takes every S-th batch of eight loads and
// the partial sums meet in a shuffle tree (fixed order); all 32 lanes of the warp must call, `valid` masks the loads.
static __device__ __forceinline__ float2 gn_fold_group(const GnFoldP& f, int n, int g, int sub = 0, int S = 1, bool valid = true) {
  double a = 0.0, b = 0.0;
  // the group's channels [g*cpg, (g+1)*cpg) may straddle the concat boundary: one run per source
#pragma unroll 1
  for (int src = 0; src < (valid ? 2 : 0); ++src) {
    const int off = src ? f.C0 : 0, Cs = src ? f.C1 : f.C0;
    const int lo = max(g * f.cpg, off) - off, hi = min((g + 1) * f.cpg, off + Cs) - off;
    if (hi <= lo) continue;
    const int slots = src ? f.slots1 : f.slots0, pairs = (hi - lo) >> 1;
    const float4* base = reinterpret_cast<const float4*>((src ? f.part1 : f.part0) + ((long long)n * slots * Cs + lo) * 2);
    const int tot = pairs * slots, half_c = Cs >> 1;   // float4 units: channel pair q of slot s sits at s*Cs/2 + q
#pragma unroll 1
    for (int k0 = 8 * sub; k0 < tot; k0 += 8 * S) {
      float4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int kk = min(k0 + k, tot - 1), q = kk / slots, sl = kk - q * slots;
        u[k] = __ldg(base + (long long)sl * half_c + q);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k0 + k < tot) a += (double)u[k].x + (double)u[k].z, b += (double)u[k].y + (double)u[k].w;
    }
  }
  for (int o = 1; o < S; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o), b += __shfl_xor_sync(0xffffffffu, b, o);
  const double m = a * (double)f.inv_cnt;
  double var = b * (double)f.inv_cnt - m * m;
  var = var < 0.0 ? 0.0 : var;
  return make_float2((float)m, (float)(1.0 / sqrt(var + (double)f.eps)));
}

}  // namespace hsidm
