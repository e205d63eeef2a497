// GroupNorm (+Swish) over NHWC activations, including the virtual concatenation of two tensors whose groups
// may straddle the concat boundary (unet.py:84, :120, :259; C = 192/384/768 in the `ups` blocks).
//
// Both kernels are HBM-bound: 128-bit accesses (8 channels per thread), warp-shuffle-free per-channel register
// accumulation, one shared-memory fold per block, fp64 atomics for the few cross-block partials.
//   gn_stats : read x once                 -> (sum, sumsq) per (image, group)
//   gn_apply : read x once, write y once   -> y = swish?(x * A[n,c] + B[n,c])
#include "kernels.cuh"

namespace hsidm {
namespace {

constexpr int kMaxThreads = 256;

template <typename AT>
__device__ __forceinline__ const AT* src_ptr(const AT* x0, int C0, const AT* x1, int C1, int64_t pix, int c) {
  return c < C0 ? x0 + pix * C0 + c : x1 + pix * C1 + (c - C0);
}

// grid = (slabs, N); block = lanes*CV threads, CV = C/8 channel-vectors, each thread owns one channel-vector.
template <typename AT>
__global__ void gn_stats_kernel(const AT* __restrict__ x0, int C0, const AT* __restrict__ x1, int C1, int HW,
                                int groups, int CV, int lanes, int pix_per_block, double* __restrict__ gsum) {
  extern __shared__ float sm[];  // [2][C]
  const int C = C0 + C1;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int cv = tid % CV, lane = tid / CV;
  const int c = cv * 8;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  for (int i = tid; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f, q[j] = 0.f;
  for (int pix = p0 + lane; pix < p1; pix += lanes) {
    float v[8];
    load8(src_ptr<AT>(x0, C0, x1, C1, (int64_t)n * HW + pix, c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += v[j], q[j] = fmaf(v[j], v[j], q[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&sm[c + j], s[j]);
    atomicAdd(&sm[C + c + j], q[j]);
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = tid; g < groups; g += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int j = 0; j < cpg; ++j) a += (double)sm[g * cpg + j], b += (double)sm[C + g * cpg + j];
    atomicAdd(&gsum[((int64_t)n * groups + g) * 2 + 0], a);
    atomicAdd(&gsum[((int64_t)n * groups + g) * 2 + 1], b);
  }
}

template <typename AT>
__global__ void gn_apply_kernel(const AT* __restrict__ x0, int C0, const AT* __restrict__ x1, int C1, int HW,
                                int groups, int CV, int lanes, int pix_per_block, const double* __restrict__ gsum,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int swish,
                                AT* __restrict__ out) {
  extern __shared__ float sm[];  // A[C], B[C]
  const int C = C0 + C1;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int cpg = C / groups;
  const double cnt = (double)cpg * HW;
  for (int c = tid; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double mean = gsum[((int64_t)n * groups + g) * 2] / cnt;
    double var = gsum[((int64_t)n * groups + g) * 2 + 1] / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float a = rstd * gamma[c];
    sm[c] = a;
    sm[C + c] = beta[c] - (float)mean * a;
  }
  __syncthreads();
  const int cv = tid % CV, lane = tid / CV;
  const int c = cv * 8;
  float A[8], B[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) A[j] = sm[c + j], B[j] = sm[C + c + j];
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  for (int pix = p0 + lane; pix < p1; pix += lanes) {
    float v[8];
    const int64_t gp = (int64_t)n * HW + pix;
    load8(src_ptr<AT>(x0, C0, x1, C1, gp, c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = fmaf(v[j], A[j], B[j]);
      v[j] = swish ? swish_f(y) : y;
    }
    store8(out + gp * C + c, v);
  }
}

struct Geo {
  int CV, lanes, threads, pix_per_block, slabs;
};

int geometry(int C0, int C1, int N, int HW, Geo* g) {
  const int C = C0 + C1;
  if (C0 % 8 || C1 % 8 || C <= 0) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "GroupNorm needs channel counts that are multiples of 8 (got %d+%d)", C0, C1);
  g->CV = C / 8;
  if (g->CV > kMaxThreads) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "GroupNorm over %d channels exceeds the supported 2048", C);
  g->lanes = kMaxThreads / g->CV;
  g->threads = g->lanes * g->CV;
  // enough blocks to fill 148 SMs a few times over, but at least 4 pixels per lane per block
  int slabs = (int)ceil_div(148 * 8, N);
  int min_pix = g->lanes * 4;
  slabs = (int)std::min<int64_t>(slabs, ceil_div(HW, min_pix));
  if (slabs < 1) slabs = 1;
  g->pix_per_block = (int)ceil_div(HW, slabs);
  g->slabs = (int)ceil_div(HW, g->pix_per_block);
  return HSIDM_OK;
}

}  // namespace

int gn_stats(const void* x0, int C0, const void* x1, int C1, int N, int HW, int groups, double* gsum, int prec,
             cudaStream_t stream) {
  Geo g;
  HSIDM_TRY(geometry(C0, C1, N, HW, &g));
  const int C = C0 + C1;
  if (C % groups) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "GroupNorm: %d channels not divisible by %d groups", C, groups);
  HSIDM_CUDA(cudaMemsetAsync(gsum, 0, sizeof(double) * 2 * N * groups, stream));
  dim3 grid(g.slabs, N);
  size_t smem = sizeof(float) * 2 * C;
  if (prec == HSIDM_BF16)
    gn_stats_kernel<bf16><<<grid, g.threads, smem, stream>>>((const bf16*)x0, C0, (const bf16*)x1, C1, HW, groups, g.CV,
                                                              g.lanes, g.pix_per_block, gsum);
  else
    gn_stats_kernel<float><<<grid, g.threads, smem, stream>>>((const float*)x0, C0, (const float*)x1, C1, HW, groups,
                                                               g.CV, g.lanes, g.pix_per_block, gsum);
  return after_launch("gn_stats_kernel");
}

int gn_apply(const void* x0, int C0, const void* x1, int C1, int N, int HW, int groups, const double* gsum,
             const float* gamma, const float* beta, float eps, int swish, void* out, int prec, cudaStream_t stream) {
  Geo g;
  HSIDM_TRY(geometry(C0, C1, N, HW, &g));
  const int C = C0 + C1;
  dim3 grid(g.slabs, N);
  size_t smem = sizeof(float) * 2 * C;
  if (prec == HSIDM_BF16)
    gn_apply_kernel<bf16><<<grid, g.threads, smem, stream>>>((const bf16*)x0, C0, (const bf16*)x1, C1, HW, groups, g.CV,
                                                              g.lanes, g.pix_per_block, gsum, gamma, beta, eps, swish,
                                                              (bf16*)out);
  else
    gn_apply_kernel<float><<<grid, g.threads, smem, stream>>>((const float*)x0, C0, (const float*)x1, C1, HW, groups,
                                                               g.CV, g.lanes, g.pix_per_block, gsum, gamma, beta, eps,
                                                               swish, (float*)out);
  return after_launch("gn_apply_kernel");
}

}  // namespace hsidm
