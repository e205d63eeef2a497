// GroupNorm (+Swish) over NHWC activations, including the virtual concatenation of two tensors whose groups
// may straddle the concat boundary (unet.py:84, :120, :259; C = 192/384/768 in the `ups` blocks).
//
// Both kernels are HBM-bound: 128-bit accesses (8 channels per thread), 4 independent loads in flight per thread,
// per-channel register accumulation, one shared-memory fold per block.  The reduction is DETERMINISTIC: every
// block writes its per-group partial (sum, sumsq) to a fixed slot, and the last block of an image to finish
// (ticket counter) adds the slots in index order and publishes (mean, rstd).  No floating-point atomics.
//   gn_stats : read x once                 -> stats[n][g] = (mean, rstd)
//   gn_apply : read x once, write y once   -> y = swish?(x * A[n,c] + B[n,c])
#include <algorithm>

#include "gn_fold.cuh"

namespace hsidm {
namespace {

constexpr int kMaxThreads = 256;
constexpr int kUnroll = 4;
constexpr int kMaxGroups = 128;   // groups a block can fold itself (gn_apply_fused)

// Swish.  fp32 mode keeps the exact x*sigmoid(x); bf16 mode uses the single-MUFU identity sigmoid(y) = 0.5 + 0.5*tanh(y/2)
// (tanh.approx error ~2^-11, far below the bf16 rounding of the result) so the pass stays HBM-bound, not MUFU-bound.
template <typename AT>
__device__ __forceinline__ float swish_act(float y);
template <>
__device__ __forceinline__ float swish_act<float>(float y) {
  return swish_f(y);
}
template <>
__device__ __forceinline__ float swish_act<bf16>(float y) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * y));
  return y * fmaf(0.5f, t, 0.5f);
}

// 8 consecutive activations kept in their storage format while in flight (4 registers for bf16, 8 for fp32)
template <typename AT>
struct Raw8;
template <>
struct Raw8<bf16> {
  uint4 r;
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x, v[2 * i + 1] = f.y;
    }
  }
};
template <>
struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  }
};

template <typename AT>
__device__ __forceinline__ const AT* src_ptr(const AT* x0, int C0, const AT* x1, int C1, int64_t pix, int c) {
  return c < C0 ? x0 + pix * C0 + c : x1 + pix * C1 + (c - C0);
}

// grid = (slabs, N); block = lanes*CV threads, CV = C/8 channel-vectors, each thread owns one channel-vector.
template <typename AT>
__global__ void __launch_bounds__(kMaxThreads, 4)
gn_stats_kernel(const AT* __restrict__ x0, int C0, const AT* __restrict__ x1, int C1, int HW, int groups, int CV,
                int lanes, int pix_per_block, float eps, double* __restrict__ partial /*[N][slabs][groups][2]*/,
                unsigned* __restrict__ tickets /*[N], zero on entry and on exit*/, float* __restrict__ stats) {
  extern __shared__ float sm[];  // [lanes][2][C]  then reused as [2][C]
  __shared__ bool is_last;
  const int C = C0 + C1;
  const int n = blockIdx.y, slab = blockIdx.x, slabs = gridDim.x;
  const int tid = threadIdx.x;
  const int cv = tid % CV, lane = tid / CV;
  const int c = cv * 8;
  const int p0 = slab * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f, q[j] = 0.f;
  const int64_t img = (int64_t)n * HW;
  int pix = p0 + lane;
  for (; pix + (kUnroll - 1) * lanes < p1; pix += kUnroll * lanes) {
    float v[kUnroll][8];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) load8(src_ptr<AT>(x0, C0, x1, C1, img + pix + u * lanes, c), v[u]);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[u][j], q[j] = fmaf(v[u][j], v[u][j], q[j]);
  }
  for (; pix < p1; pix += lanes) {
    float v[8];
    load8(src_ptr<AT>(x0, C0, x1, C1, img + pix, c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += v[j], q[j] = fmaf(v[j], v[j], q[j]);
  }
  float* mine = sm + (size_t)lane * 2 * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) mine[c + j] = s[j], mine[C + c + j] = q[j];
  __syncthreads();
  // fold the lanes in index order: thread -> (which, channel)
  for (int i = tid; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += sm[(size_t)l * 2 * C + i];
    sm[i] = acc;  // lane 0's slot doubles as the result (each i is read by this thread only before the write)
  }
  __syncthreads();
  const int cpg = C / groups;
  double* my_partial = partial + ((int64_t)n * slabs + slab) * groups * 2;
  for (int g = tid; g < groups; g += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int j = 0; j < cpg; ++j) a += (double)sm[g * cpg + j], b += (double)sm[C + g * cpg + j];
    my_partial[g * 2] = a;
    my_partial[g * 2 + 1] = b;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(&tickets[n], 1u) == (unsigned)(slabs - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const double cnt = (double)cpg * HW;
  for (int g = tid; g < groups; g += blockDim.x) {
    double a = 0.0, b = 0.0;
    const double* p = partial + (int64_t)n * slabs * groups * 2 + g * 2;
    for (int sl = 0; sl < slabs; ++sl) a += p[(int64_t)sl * groups * 2], b += p[(int64_t)sl * groups * 2 + 1];
    const double mean = a / cnt;
    double var = b / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    stats[((int64_t)n * groups + g) * 2] = (float)mean;
    stats[((int64_t)n * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  if (tid == 0) tickets[n] = 0;  // ready for the next GroupNorm on this stream
}

template <typename AT>
__global__ void __launch_bounds__(kMaxThreads, 4)
gn_apply_kernel(const AT* __restrict__ x0, int C0, const AT* __restrict__ x1, int C1, int HW, int groups, int CV,
                int lanes, int pix_per_block, const float* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, int swish, AT* __restrict__ out, const GnFoldP fold, float eps_unused) {
  __shared__ float2 gs[kMaxGroups];   // (mean, rstd) per group when the block folds the statistics itself
  griddep_wait();
  griddep_launch();
  const int C = C0 + C1;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int cpg = C / groups;
  const int cv = tid % CV, lane = tid / CV;
  const int c = cv * 8;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const int64_t img = (int64_t)n * HW;
  int pix = p0 + lane;
  // first batch of loads goes out before the (dependent) statistics / affine prologue
  Raw8<AT> v[kUnroll];
  bool have = pix + (kUnroll - 1) * lanes < p1;
  if (have) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) v[u].load(src_ptr<AT>(x0, C0, x1, C1, img + pix + u * lanes, c));
  }
  if (fold.part0) {   // statistics from the producers' per-slot partial sums: no statistics / finalize kernel ran
    for (int g = tid; g < groups; g += blockDim.x) gs[g] = gn_fold_group(fold, n, g);
    __syncthreads();
  }
  float A[8], B[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cpg;
    float mean, rstd;
    if (fold.part0) mean = gs[g].x, rstd = gs[g].y;
    else mean = stats[((int64_t)n * groups + g) * 2], rstd = stats[((int64_t)n * groups + g) * 2 + 1];
    A[j] = rstd * __ldg(gamma + c + j);
    B[j] = __ldg(beta + c + j) - mean * A[j];
  }
  while (have) {
    const int cur = pix;
    pix += kUnroll * lanes;
    Raw8<AT> w[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) w[u] = v[u];
    have = pix + (kUnroll - 1) * lanes < p1;
    if (have) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) v[u].load(src_ptr<AT>(x0, C0, x1, C1, img + pix + u * lanes, c));
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float f[8];
      w[u].unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = fmaf(f[j], A[j], B[j]);
        f[j] = swish ? swish_act<AT>(y) : y;
      }
      store8(out + (img + cur + u * lanes) * C + c, f);
    }
  }
  for (; pix < p1; pix += lanes) {
    float v[8];
    load8(src_ptr<AT>(x0, C0, x1, C1, img + pix, c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float y = fmaf(v[j], A[j], B[j]);
      v[j] = swish ? swish_act<AT>(y) : y;
    }
    store8(out + (img + pix) * C + c, v);
  }
}

// GroupNorm-apply for the attention block (bf16): y = x*A[n,c] + B[n,c] written twice - as [N][S][C] (the tensor the
// projections read) and transposed as [N][C][S] (keys contiguous: the K-major B operand of P.X).  One block = 64 pixels
// x 64 channels of one image: 128-byte rows in, 128-byte rows out in both layouts, transposed through shared memory.
__global__ void __launch_bounds__(256)
gn_apply_t_kernel(const bf16* __restrict__ x, int S, int C, const GnFoldP fold, const float* __restrict__ stats, int groups,
                  const float* __restrict__ gamma, const float* __restrict__ beta, bf16* __restrict__ y, bf16* __restrict__ yt) {
  __shared__ bf16 tile[64][66];   // [channel][pixel], padded: both access patterns conflict-free enough
  __shared__ float2 gs[64];       // (mean, rstd) of the groups this block's 64 channels belong to
  griddep_wait();
  griddep_launch();
  const int n = blockIdx.z, c0 = blockIdx.y * 64;
  const int tid = threadIdx.x;
  const int j8 = tid & 7;          // 16-byte chunk (8 channels) of a pixel's 64-channel row
  const int cpg = C / groups, g_lo = c0 / cpg, g_n = (c0 + 63) / cpg - g_lo + 1;
  if (tid < g_n) {
    const int g = g_lo + tid;
    gs[tid] = fold.part0 ? gn_fold_group(fold, n, g) : make_float2(stats[((int64_t)n * groups + g) * 2], stats[((int64_t)n * groups + g) * 2 + 1]);
  }
  __syncthreads();
  float A[8], B[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j8 * 8 + j;
    const float2 st = gs[c / cpg - g_lo];
    A[j] = st.y * __ldg(gamma + c);
    B[j] = __ldg(beta + c) - st.x * A[j];
  }
  // the block walks the 64-pixel tiles blockIdx.x, blockIdx.x + gridDim.x, ... of its channel slice: the statistics fold
  // and the coefficients above are paid once per block, not once per 8 KB tile
  for (int p0 = blockIdx.x * 64; p0 < S; p0 += gridDim.x * 64) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int px = (tid >> 3) + 32 * it;
      const int64_t off = ((int64_t)n * S + p0 + px) * C + c0 + j8 * 8;
      float f[8];
      load8(x + off, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], A[j], B[j]);
      store8(y + off, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) tile[j8 * 8 + j][px] = __float2bfloat16_rn(f[j]);
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int ch = (tid >> 3) + 32 * it;
      uint4 o;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __halves2bfloat162(tile[ch][j8 * 8 + 2 * j], tile[ch][j8 * 8 + 2 * j + 1]);
      *reinterpret_cast<uint4*>(yt + ((int64_t)n * C + c0 + ch) * S + p0 + j8 * 8) = o;
    }
    __syncthreads();   // the tile is rewritten by the next pixel tile
  }
}

}  // namespace

static int gn_apply_impl(const void* x0, int C0, const void* x1, int C1, int N, int HW, int groups, const float* stats,
                         const float* gamma, const float* beta, int swish, void* out, int prec, cudaStream_t stream, const GnFoldP& f);

// Fills the device-side fold descriptor; fails when the block-local fold cannot take this GroupNorm.
static int make_fold(const GnIn& gn, int HW, GnFoldP* f) {
  const int Ct = gn.C[0] + gn.C[1];
  if (gn.groups <= 0 || Ct % gn.groups) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "GroupNorm: %d channels not divisible by %d groups", Ct, gn.groups);
  *f = GnFoldP{nullptr, nullptr, 0, 0, gn.C[0], gn.C[1], Ct / gn.groups, (float)(1.0 / ((double)(Ct / gn.groups) * HW)), gn.eps};
  if (gn.stats) return HSIDM_OK;
  if (!gn.part[0] || (gn.C[1] && !gn.part[1]) || (f->cpg & 1) || gn.C[0] % 2)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "GroupNorm from partial sums needs both sources' slots and an even number of channels per group");
  f->part0 = gn.part[0], f->part1 = gn.part[1], f->slots0 = gn.slots[0], f->slots1 = gn.slots[1];
  return HSIDM_OK;
}

int gn_apply_transposed(const void* x, int N, int S, const GnIn& gn, void* y, void* yt, cudaStream_t stream) {
  const int C = gn.C[0];
  if (S % 64 || C % 64 || gn.C[1]) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "gn_apply_transposed: S=%d and C=%d must be multiples of 64 (one source)", S, C);
  GnFoldP f;
  HSIDM_TRY(make_fold(gn, S, &f));
  char tag[64];
  snprintf(tag, sizeof(tag), "apply+T c%d hw%d n%d", C, S, N);
  ProfScope prof(PROF_GN_APPLY, 3.0 * N * S * C * 2, stream, tag);
  // enough blocks to fill the machine a few times over, as few pixel tiles per image as that allows (variant 16384: one
  // block per 64x64 tile, the round-2 form)
  int gx = S / 64;
  if (!(conv_tc_variant() & 16384))
    while (gx > 1 && gx % 2 == 0 && (int64_t)(gx / 2) * (C / 64) * N >= 8 * 148) gx /= 2;
  HSIDM_CUDA(launch_pdl(gn_apply_t_kernel, dim3(gx, C / 64, N), dim3(256), 0, stream, 1, (const bf16*)x, S, C, f, gn.stats, gn.groups,
                        gn.gamma, gn.beta, (bf16*)y, (bf16*)yt));
  return after_launch("gn_apply_t_kernel");
}

int gn_apply_fused(const void* x0, const void* x1, int N, int HW, const GnIn& gn, void* out, int prec, cudaStream_t stream) {
  GnFoldP f;
  HSIDM_TRY(make_fold(gn, HW, &f));
  if (!gn.stats && gn.groups > kMaxGroups) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "gn_apply_fused: more than %d groups", kMaxGroups);
  return gn_apply_impl(x0, gn.C[0], x1, gn.C[1], N, HW, gn.groups, gn.stats, gn.gamma, gn.beta, gn.swish, out, prec, stream, f);
}

int gn_geometry(int C0, int C1, int N, int HW, GnGeo* g) {
  const int C = C0 + C1;
  if (C0 % 8 || C1 % 8 || C <= 0)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "GroupNorm needs channel counts that are multiples of 8 (got %d+%d)", C0, C1);
  g->CV = C / 8;
  if (g->CV > kMaxThreads) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "GroupNorm over %d channels exceeds the supported 2048", C);
  g->lanes = kMaxThreads / g->CV;
  g->threads = g->lanes * g->CV;
  // many more blocks than resident slots (148 SMs x 4) so the last wave is a small fraction of the run, but at
  // least 2*kUnroll pixels per lane per block
  int slabs = (int)ceil_div(148 * 4 * 10, N);
  const int min_pix = g->lanes * 2 * kUnroll;
  slabs = (int)std::min<int64_t>(slabs, ceil_div(HW, min_pix));
  if (slabs < 1) slabs = 1;
  g->pix_per_block = (int)ceil_div(HW, slabs);
  g->slabs = (int)ceil_div(HW, g->pix_per_block);
  return HSIDM_OK;
}

// gn_apply geometry: about one resident wave (148 SMs x 4 blocks) of long-lived blocks, so the per-block prologue
// (statistics + affine -> A, B) is amortised over hundreds of KB and there is no partial last wave.
static void apply_geometry(const GnGeo& base, int N, int HW, GnGeo* g) {
  *g = base;
  int slabs = std::max(1, (148 * 4) / std::max(1, N));
  const int min_pix = g->lanes * 2 * kUnroll;
  slabs = (int)std::min<int64_t>(slabs, std::max<int64_t>(1, HW / min_pix));
  g->pix_per_block = (int)ceil_div(HW, slabs);
  g->slabs = (int)ceil_div(HW, g->pix_per_block);
}

int64_t gn_scratch_bytes(int C0, int C1, int N, int HW, int groups) {
  GnGeo g;
  if (gn_geometry(C0, C1, N, HW, &g) != HSIDM_OK) return 0;
  return (int64_t)sizeof(double) * 2 * N * g.slabs * groups + (int64_t)sizeof(float) * 2 * N * groups;
}

int gn_stats(const void* x0, int C0, const void* x1, int C1, int N, int HW, int groups, float eps, void* scratch,
             unsigned* tickets, float* stats, int prec, cudaStream_t stream) {
  GnGeo g;
  HSIDM_TRY(gn_geometry(C0, C1, N, HW, &g));
  const int C = C0 + C1;
  if (C % groups) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "GroupNorm: %d channels not divisible by %d groups", C, groups);
  double* partial = static_cast<double*>(scratch);
  dim3 grid(g.slabs, N);
  const size_t smem = sizeof(float) * 2 * C * g.lanes;
  ProfScope prof(PROF_GN_STATS, (double)N * HW * C * (prec == HSIDM_BF16 ? 2 : 4), stream);
  if (prec == HSIDM_BF16)
    gn_stats_kernel<bf16><<<grid, g.threads, smem, stream>>>((const bf16*)x0, C0, (const bf16*)x1, C1, HW, groups, g.CV,
                                                              g.lanes, g.pix_per_block, eps, partial, tickets, stats);
  else
    gn_stats_kernel<float><<<grid, g.threads, smem, stream>>>((const float*)x0, C0, (const float*)x1, C1, HW, groups,
                                                               g.CV, g.lanes, g.pix_per_block, eps, partial, tickets,
                                                               stats);
  return after_launch("gn_stats_kernel");
}

// grid = N; thread -> (slot lane, channel pair), float4 loads = (sum, sumsq) of two channels; fixed-order folds only.
// The per-slot partials are short fp32 sums (32 rows); everything above them is folded in float64, so the statistics do
// not depend on how the producing kernel happened to tile the tensor beyond the last fp32 bit of a 32-row sum, and
// E[x^2] - mean^2 keeps its digits for channels whose mean dominates their spread.
__global__ void __launch_bounds__(1024, 2)   // two blocks per SM: a 176-image batch is one wave, not two
gn_finalize_kernel(const float* __restrict__ part0, int slots0, int C0, const float* __restrict__ part1, int slots1, int C1,
                   int groups, int lanes, double cnt, float eps, float* __restrict__ stats, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ ab) {
  extern __shared__ double dsm[];   // [lanes][C/2][4], then reused as double[2][C]
  griddep_wait();
  griddep_launch();
  const int C = C0 + C1, n = blockIdx.x, tid = threadIdx.x;
  const int pairs = C >> 1;
  const int pr = tid % pairs, lane = tid / pairs;
  if (lane < lanes) {
    const int c = 2 * pr;
    const bool second = c >= C0;
    const float* base = second ? part1 + ((long long)n * slots1 * C1 + (c - C0)) * 2 : part0 + ((long long)n * slots0 * C0 + c) * 2;
    const int slots = second ? slots1 : slots0;
    const long long stride = (long long)(second ? C1 : C0) * 2;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int sl = lane;
    for (; sl + 3 * lanes < slots; sl += 4 * lanes) {   // four loads in flight, folded in slot order
      const float4 u0 = *reinterpret_cast<const float4*>(base + (long long)sl * stride);
      const float4 u1 = *reinterpret_cast<const float4*>(base + (long long)(sl + lanes) * stride);
      const float4 u2 = *reinterpret_cast<const float4*>(base + (long long)(sl + 2 * lanes) * stride);
      const float4 u3 = *reinterpret_cast<const float4*>(base + (long long)(sl + 3 * lanes) * stride);
      a0 += ((double)u0.x + (double)u1.x) + ((double)u2.x + (double)u3.x);
      a1 += ((double)u0.y + (double)u1.y) + ((double)u2.y + (double)u3.y);
      a2 += ((double)u0.z + (double)u1.z) + ((double)u2.z + (double)u3.z);
      a3 += ((double)u0.w + (double)u1.w) + ((double)u2.w + (double)u3.w);
    }
    for (; sl < slots; sl += lanes) {
      const float4 u0 = *reinterpret_cast<const float4*>(base + (long long)sl * stride);
      a0 += (double)u0.x, a1 += (double)u0.y, a2 += (double)u0.z, a3 += (double)u0.w;
    }
    double* mine = dsm + ((long long)lane * pairs + pr) * 4;
    mine[0] = a0, mine[1] = a1, mine[2] = a2, mine[3] = a3;
  }
  __syncthreads();
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  if (tid < pairs)
    for (int l = 0; l < lanes; ++l) {
      const double* u = dsm + ((long long)l * pairs + tid) * 4;
      t0 += u[0], t1 += u[1], t2 += u[2], t3 += u[3];
    }
  __syncthreads();
  double* fs = dsm;   // [2][C]: sums then sums of squares
  if (tid < pairs) {
    fs[2 * tid] = t0, fs[2 * tid + 1] = t2;
    fs[C + 2 * tid] = t1, fs[C + 2 * tid + 1] = t3;
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = tid; g < groups; g += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int j = 0; j < cpg; ++j) a += fs[g * cpg + j], b += fs[C + g * cpg + j];
    const double mean = a / cnt;
    double var = b / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    stats[((long long)n * groups + g) * 2] = (float)mean;
    stats[((long long)n * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  if (ab) {   // per-channel affine, the same expressions gn_apply evaluates
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {
      const int g = c / cpg;
      const float mean = stats[((long long)n * groups + g) * 2], rstd = stats[((long long)n * groups + g) * 2 + 1];
      const float A = rstd * __ldg(gamma + c);
      *reinterpret_cast<float2*>(ab + ((long long)n * C + c) * 2) = make_float2(A, __ldg(beta + c) - mean * A);
    }
  }
}

int gn_finalize(const float* part0, int slots0, int C0, const float* part1, int slots1, int C1, int N, int HW, int groups,
                float eps, float* stats, cudaStream_t stream, const float* gamma, const float* beta, float* ab) {
  const int C = C0 + C1;
  if (C0 % 64 || C1 % 64 || C % groups || C > 2048)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "gn_finalize: channel counts %d+%d must be multiples of 64 and of the group count", C0, C1);
  const int pairs = C / 2;
  const int lanes = std::max(1, std::min(32, 1024 / pairs));
  const int threads = (int)round_up(lanes * pairs, 32);
  const size_t smem = std::max<size_t>(sizeof(double) * 4 * lanes * pairs, sizeof(double) * 2 * C);
  ProfScope prof(PROF_GN_STATS, 8.0 * N * ((double)slots0 * C0 + (double)slots1 * C1), stream, "finalize");
  HSIDM_CUDA(launch_pdl(gn_finalize_kernel, dim3(N), dim3(threads), smem, stream, 1, part0, slots0, C0, part1, slots1, C1, groups, lanes,
                        (double)(C / groups) * HW, eps, stats, gamma, beta, ab));
  return after_launch("gn_finalize_kernel");
}

int gn_apply(const void* x0, int C0, const void* x1, int C1, int N, int HW, int groups, const float* stats,
             const float* gamma, const float* beta, int swish, void* out, int prec, cudaStream_t stream) {
  return gn_apply_impl(x0, C0, x1, C1, N, HW, groups, stats, gamma, beta, swish, out, prec, stream, GnFoldP{});
}

static int gn_apply_impl(const void* x0, int C0, const void* x1, int C1, int N, int HW, int groups, const float* stats,
                         const float* gamma, const float* beta, int swish, void* out, int prec, cudaStream_t stream, const GnFoldP& f) {
  GnGeo g0, g;
  HSIDM_TRY(gn_geometry(C0, C1, N, HW, &g0));
  apply_geometry(g0, N, HW, &g);
  dim3 grid(g.slabs, N);
  char tag[64];
  snprintf(tag, sizeof(tag), "apply c%d+%d hw%d n%d grid%dx%d", C0, C1, HW, N, g.slabs, N);
  ProfScope prof(PROF_GN_APPLY, 2.0 * N * HW * (C0 + C1) * (prec == HSIDM_BF16 ? 2 : 4), stream, tag);
  if (prec == HSIDM_BF16)
    HSIDM_CUDA(launch_pdl(gn_apply_kernel<bf16>, grid, dim3(g.threads), 0, stream, 1, (const bf16*)x0, C0, (const bf16*)x1, C1, HW, groups,
                          g.CV, g.lanes, g.pix_per_block, stats, gamma, beta, swish, (bf16*)out, f, 0.f));
  else
    HSIDM_CUDA(launch_pdl(gn_apply_kernel<float>, grid, dim3(g.threads), 0, stream, 1, (const float*)x0, C0, (const float*)x1, C1, HW,
                          groups, g.CV, g.lanes, g.pix_per_block, stats, gamma, beta, swish, (float*)out, f, 0.f));
  return after_launch("gn_apply_kernel");
}

}  // namespace hsidm
