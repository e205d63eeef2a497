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

#include "kernels.cuh"

namespace hsidm {
namespace {

constexpr int kMaxThreads = 256;
constexpr int kUnroll = 4;

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
                const float* __restrict__ beta, int swish, AT* __restrict__ out) {
  const int C = C0 + C1;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int cpg = C / groups;
  const int cv = tid % CV, lane = tid / CV;
  const int c = cv * 8;
  float A[8], B[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cpg;
    const float mean = stats[((int64_t)n * groups + g) * 2], rstd = stats[((int64_t)n * groups + g) * 2 + 1];
    A[j] = rstd * __ldg(gamma + c + j);
    B[j] = __ldg(beta + c + j) - mean * A[j];
  }
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const int64_t img = (int64_t)n * HW;
  int pix = p0 + lane;
  for (; pix + (kUnroll - 1) * lanes < p1; pix += kUnroll * lanes) {
    float v[kUnroll][8];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) load8(src_ptr<AT>(x0, C0, x1, C1, img + pix + u * lanes, c), v[u]);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = fmaf(v[u][j], A[j], B[j]);
        v[u][j] = swish ? swish_act<AT>(y) : y;
      }
      store8(out + (img + pix + u * lanes) * C + c, v[u]);
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

}  // namespace

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

// grid = N, block = 256: thread -> (slot lane, channel); fixed-order folds only.
__global__ void gn_finalize_kernel(const float* __restrict__ part0, int slots0, int C0, const float* __restrict__ part1,
                                   int slots1, int C1, int groups, double cnt, float eps, float* __restrict__ stats) {
  extern __shared__ float fsm[];   // [lanes][2][C] then [2][C]
  const int C = C0 + C1, n = blockIdx.x, tid = threadIdx.x;
  const int lanes = blockDim.x / 64;                     // channels are walked 64 at a time
  const int cl = tid & 63, lane = tid >> 6;
  for (int cb = 0; cb < C; cb += 64) {
    const int c = cb + cl;
    const bool second = c >= C0;
    const float* base = second ? part1 + ((long long)n * slots1 * C1 + (c - C0)) * 2 : part0 + ((long long)n * slots0 * C0 + c) * 2;
    const int slots = second ? slots1 : slots0, stride = (second ? C1 : C0) * 2;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    int sl = lane;
    for (; sl + lanes < slots; sl += 2 * lanes) {
      const float2 u = *reinterpret_cast<const float2*>(base + (long long)sl * stride);
      const float2 v = *reinterpret_cast<const float2*>(base + (long long)(sl + lanes) * stride);
      a0 += u.x, b0 += u.y, a1 += v.x, b1 += v.y;
    }
    if (sl < slots) {
      const float2 u = *reinterpret_cast<const float2*>(base + (long long)sl * stride);
      a0 += u.x, b0 += u.y;
    }
    fsm[(lane * 2 + 0) * C + c] = a0 + a1;
    fsm[(lane * 2 + 1) * C + c] = b0 + b1;
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += blockDim.x) {
    const int which = i / C, c = i - which * C;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += fsm[(l * 2 + which) * C + c];
    fsm[which * C + c] = acc;   // lane 0's own slot: read above by this thread only, no other thread touches it
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = tid; g < groups; g += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int j = 0; j < cpg; ++j) a += (double)fsm[g * cpg + j], b += (double)fsm[C + g * cpg + j];
    const double mean = a / cnt;
    double var = b / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    stats[((long long)n * groups + g) * 2] = (float)mean;
    stats[((long long)n * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

int gn_finalize(const float* part0, int slots0, int C0, const float* part1, int slots1, int C1, int N, int HW, int groups,
                float eps, float* stats, cudaStream_t stream) {
  const int C = C0 + C1;
  if (C0 % 64 || C1 % 64 || C % groups || C > 4096)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "gn_finalize: channel counts %d+%d must be multiples of 64 and of the group count", C0, C1);
  // as many slot lanes as 48 KB of shared memory allows (16 for C = 64 ... 2 for C = 1536)
  int lanes = 16;
  while (lanes > 1 && sizeof(float) * 2 * C * lanes > 48 * 1024) lanes >>= 1;
  const int threads = 64 * lanes;
  const size_t smem = sizeof(float) * 2 * C * lanes;
  if (smem > 48 * 1024) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "gn_finalize: %d channels need too much shared memory", C);
  ProfScope prof(PROF_GN_STATS, 8.0 * N * ((double)slots0 * C0 + (double)slots1 * C1), stream, "finalize");
  gn_finalize_kernel<<<N, threads, smem, stream>>>(part0, slots0, C0, part1, slots1, C1, groups, (double)(C / groups) * HW, eps,
                                                   stats);
  return after_launch("gn_finalize_kernel");
}

int gn_apply(const void* x0, int C0, const void* x1, int C1, int N, int HW, int groups, const float* stats,
             const float* gamma, const float* beta, int swish, void* out, int prec, cudaStream_t stream) {
  GnGeo g;
  HSIDM_TRY(gn_geometry(C0, C1, N, HW, &g));
  dim3 grid(g.slabs, N);
  ProfScope prof(PROF_GN_APPLY, 2.0 * N * HW * (C0 + C1) * (prec == HSIDM_BF16 ? 2 : 4), stream);
  if (prec == HSIDM_BF16)
    gn_apply_kernel<bf16><<<grid, g.threads, 0, stream>>>((const bf16*)x0, C0, (const bf16*)x1, C1, HW, groups, g.CV,
                                                           g.lanes, g.pix_per_block, stats, gamma, beta, swish, (bf16*)out);
  else
    gn_apply_kernel<float><<<grid, g.threads, 0, stream>>>((const float*)x0, C0, (const float*)x1, C1, HW, groups, g.CV,
                                                            g.lanes, g.pix_per_block, stats, gamma, beta, swish,
                                                            (float*)out);
  return after_launch("gn_apply_kernel");
}

}  // namespace hsidm
