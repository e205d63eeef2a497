// Micro-benchmark: tcgen05.mma.cta_group::2 (CTA pair, M = 256) issue rate with both operands in shared memory.
// Each CTA holds its own 128 A rows and HALF of the B rows (N/2) at the same shared-memory offsets; the leader (cluster
// rank 0) issues, completion is multicast to a barrier in both CTAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I hsi_dmgasr_b200/csrc scripts/mma_rate2.cu -o build_tmp/mma_rate2 -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace hsidm;
using namespace hsidm::tc;

static __device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
static __device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
static __device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
static __device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
static __device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
static __device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2_kernel(long long* out, int iters, int nacc, float* probe) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // A: rows hold (1 + rank) in k-element 0, else 0; B: rows hold 1 in k-element 0 -> D[m][n] = (1 + rank) * (#accumulations)
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  for (int r = threadIdx.x; r < 128; r += blockDim.x) {
    // SW128 K-major: logical 16-byte chunk 0 of row r sits at physical chunk (0 ^ (r & 7))
    reinterpret_cast<__nv_bfloat16*>(smem + r * 128 + ((r & 7) << 4))[0] = __float2bfloat16(1.0f + rank);
    if (r < N / 2) reinterpret_cast<__nv_bfloat16*>(smem + 96 * 1024 + r * 128 + ((r & 7) << 4))[0] = __float2bfloat16(1.0f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) mbar_init(smem_u32(&bar), 1), fence_barrier_init();
  if (warp == 0) tmem_alloc2(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && rank == 0 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(256, N);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t bdesc = umma_desc_sw128(b_base);
      for (int s = 0; s < nacc; ++s) {
        const uint64_t adesc = umma_desc_sw128(a_base);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma2_f16(tmem + s * N, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) ? 1u : 0u);
      }
    }
    umma2_commit_mc(smem_u32(&bar), 3);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  // both CTAs wait for the multicast completion on their own barrier, then read one accumulator value
  mbar_wait(smem_u32(&bar), 0, nullptr, 0);
  tc_fence_after();
  if (warp == 0) {
    uint32_t v[16];
    tmem_ld16(tmem, v);
    tmem_ld_wait();
    if (lane == 0 && blockIdx.x < 2) probe[blockIdx.x] = __uint_as_float(v[0]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tc_fence_after(), tmem_dealloc2(tmem, 512);
}

template <int N>
void run(int nacc, long long* d, float* probe) {
  const int iters = 2000;
  cudaFuncSetAttribute(rate2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) rate2_kernel<N><<<148, 128, 200 * 1024>>>(d, iters, nacc, probe);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  float p[2] = {0, 0};
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(p, probe, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / ((double)iters * nacc * 4);
  // D[0][0] after iters*4 accumulations of (1+rank)*1: only k-slice 0 holds non-zeros -> iters accumulations
  printf("cta_group::2 M=256 N=%3d nacc=%d: %6.1f clk/MMA -> %.0f MAC/clk/SM   D[0][0] leader %.0f peer %.0f (expect %d / %d)  (%s)\n", N, nacc, per,
         256.0 * N * 16 / per / 2, p[0], p[1], iters, 2 * iters, cudaGetErrorString(e));
}

int main() {
  long long* d;
  float* probe;
  cudaMalloc(&d, 8);
  cudaMalloc(&probe, 8);
  run<64>(4, d, probe);
  run<128>(2, d, probe);
  run<256>(1, d, probe);
  return 0;
}
