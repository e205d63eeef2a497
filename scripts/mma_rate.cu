// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (SS mode, SWIZZLE_128B K-major operands in shared memory)
// as a function of N and of the A-operand addressing used by the halo kernel.  One CTA per SM, one issuing thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hsi_dmgasr_b200/csrc scripts/mma_rate.cu -o /tmp/mma_rate -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace hsidm;
using namespace hsidm::tc;

static __device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// mode 0: one divergent thread runs the whole loop (what conv_halo.cu does); 1: the whole warp runs the loop, an elected
// lane issues; 2: as 1, with the next iteration's barrier probed before this iteration's MMAs are issued.
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int mode, int nacc, int ovh) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, dummy, ready;
  __shared__ uint32_t slot;
  __shared__ uint32_t flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    flag = 1;
    mbar_init(smem_u32(&bar), 1), mbar_init(smem_u32(&dummy), 1 << 19), mbar_init(smem_u32(&ready), 1);
    fence_barrier_init();
    mbar_arrive(smem_u32(&ready));   // phase 0 of `ready` is complete: waits on parity 0 return at once
  }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = umma_idesc_bf16(128, N);
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
  constexpr uint64_t hi_a = ((uint64_t)((34 * 128) >> 4) << 32 | (1ull << 46) | (2ull << 61) | (1ull << 16));
  if (warp == 1 && mode == 0) {
    if (lane == 0) {
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const int tap = it % 9;
        if (ovh & 4) mbar_wait(smem_u32(&ready), 0, nullptr, 0);
        if (ovh & 2) tc_fence_after();
        const uint64_t bdesc = umma_desc_sw128(b_base + (N <= 128 ? (it & 3) * (N * 128) : 0));
        for (int s = 0; s < nacc; ++s) {
          const uint32_t a_addr = a_base + (uint32_t)(((tap / 3) * 34 + 8 * s + tap % 3) * 128);
          const uint64_t adesc = hi_a | (uint64_t)((a_addr >> 4) & 0x3FFFu);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem + s * N, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
        }
        if (ovh & 1) umma_commit(smem_u32(&dummy));
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0, nullptr, 0);
      long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  } else if (warp == 1) {
    long long t0 = clock64();
    bool ready_next = (mode == 2 && (ovh & 4)) ? mbar_try_wait(smem_u32(&ready), 0) : true;
    for (int it = 0; it < iters; ++it) {
      const int tap = it % 9;
      if (ovh & 4) {
        if (mode == 2) {
          if (!ready_next) mbar_wait(smem_u32(&ready), 0, nullptr, 0);
          ready_next = mbar_try_wait(smem_u32(&ready), 0);   // probe for the next iteration, consumed after the MMAs
        } else {
          mbar_wait(smem_u32(&ready), 0, nullptr, 0);
        }
      }
      if (ovh & 2) tc_fence_after();
      if (ovh & 8) {
        while (*reinterpret_cast<volatile uint32_t*>(&flag) == 0) {}
      }
      if (ovh & 48) {
        uint32_t x = it;
        const int n = (ovh & 32) ? 48 : 12;
        for (int q = 0; q < n; ++q) x = x * 1664525u + 1013904223u;   // dependent chain, ~4 clk each
        if (x == 0x12345u) flag = x;
      }
      const uint64_t bdesc = umma_desc_sw128(b_base + (N <= 128 ? (it & 3) * (N * 128) : 0));
      if (elect_one()) {
        for (int s = 0; s < nacc; ++s) {
          const uint32_t a_addr = a_base + (uint32_t)(((tap / 3) * 34 + 8 * s + tap % 3) * 128);
          const uint64_t adesc = hi_a | (uint64_t)((a_addr >> 4) & 0x3FFFu);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem + s * N, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
        }
        if (ovh & 1) umma_commit(smem_u32(&dummy));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_fence_after(), tmem_dealloc(tmem, 512);
}

template <int N>
void run(const char* name, int mode, int nacc, long long* d, int ovh = 0) {
  const int iters = 2000;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int grid : {148}) {
    rate_kernel<N><<<grid, 128, 200 * 1024>>>(d, iters, mode, nacc, ovh);
    rate_kernel<N><<<grid, 128, 200 * 1024>>>(d, iters, mode, nacc, ovh);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / ((double)iters * nacc * 4);
    printf("%-12s N=%3d mode=%d nacc=%d ovh=%d: %6.1f clk/MMA  -> %.0f MAC/clk/SM  (%s)\n", name, N, mode, nacc, ovh, per,
           128.0 * N * 16 / per, cudaGetErrorString(e));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  for (int ovh : {0, 4, 8, 16, 32, 1, 2, 3}) {
    run<64>("N64", 1, 4, d, ovh);
    run<128>("N128", 1, 2, d, ovh);
    run<256>("N256", 1, 1, d, ovh);
  }
  return 0;
}
