// Micro-benchmark (round 2): does the weight-stationary form tcgen05.mma.ws with the B operand held in a collector
// buffer (collector::bN::fill / use / lastuse) lift the shared-memory operand-bandwidth cap of the N = 64 / N = 128 tiles?
// One B slice (N x 16 bf16) is shared by the MT sub-tiles of a halo tile, so it only has to be fetched once per MT MMAs:
// operand bytes per MMA drop from 6 KB to 4.5 KB (N = 64, MT = 4) and from 8 KB to 6 KB (N = 128, MT = 2).
// Also checks the .ws accumulator layout against the plain form on pseudo-random operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hsi_dmgasr_b200/csrc scripts/mma_rate3.cu -o build_tmp/mma_rate3 -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace hsidm;
using namespace hsidm::tc;

#define WS_MMA(NAME, QUAL)                                                                                                   \
  static __device__ __forceinline__ void NAME(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {           \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                          \
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::" QUAL " [%0], %1, %2, %3, p;\n\t}" ::"r"(d),            \
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)                                                                       \
                 : "memory");                                                                                                \
  }
WS_MMA(ws_b0_fill, "b0::fill")
WS_MMA(ws_b0_use, "b0::use")
WS_MMA(ws_b0_last, "b0::lastuse")
WS_MMA(ws_b0_discard, "b0::discard")
WS_MMA(ws_b1_fill, "b1::fill")
WS_MMA(ws_b1_use, "b1::use")
WS_MMA(ws_b1_last, "b1::lastuse")
WS_MMA(ws_b2_fill, "b2::fill")
WS_MMA(ws_b2_use, "b2::use")
WS_MMA(ws_b2_last, "b2::lastuse")
WS_MMA(ws_b3_fill, "b3::fill")
WS_MMA(ws_b3_use, "b3::use")
WS_MMA(ws_b3_last, "b3::lastuse")

#define A_MMA(NAME, QUAL)                                                                                                    \
  static __device__ __forceinline__ void NAME(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {           \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                          \
                 "tcgen05.mma.cta_group::1.kind::f16.collector::a::" QUAL " [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),    \
                 "l"(b), "r"(idesc), "r"(acc)                                                                               \
                 : "memory");                                                                                                \
  }
A_MMA(mma_a_fill, "fill")
A_MMA(mma_a_use, "use")
A_MMA(mma_a_last, "lastuse")

// variant: 0 plain, sub-tile outer / k inner (what conv_halo.cu does)   1 plain, k outer / sub-tile inner
//          2 .ws b0 fill/use/lastuse, k outer / sub-tile inner            3 .ws b0..b3 (one buffer per k), sub-tile outer / k inner
//          4 .ws discard everywhere (no reuse: the .ws form's own rate)   5 plain with collector::a over two N tiles (A reuse)
template <int N, int MT>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = umma_idesc_bf16(128, N);
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
  constexpr uint64_t hi_a = ((uint64_t)((34 * 128) >> 4) << 32 | (1ull << 46) | (2ull << 61) | (1ull << 16));
  if (warp == 1 && lane == 0) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int tap = it % 9;
      const uint64_t bdesc = umma_desc_sw128(b_base + (N <= 128 ? (it & 3) * (N * 128) : 0));
      uint64_t adesc[MT];
#pragma unroll
      for (int s = 0; s < MT; ++s) {
        const uint32_t a_addr = a_base + (uint32_t)(((tap / 3) * 34 + 8 * s + tap % 3) * 128);
        adesc[s] = hi_a | (uint64_t)((a_addr >> 4) & 0x3FFFu);
      }
      if (variant == 0) {
#pragma unroll
        for (int s = 0; s < MT; ++s)
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem + s * N, adesc[s] + 2 * k, bdesc + 2 * k, idesc, 1u);
      } else if (variant == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int s = 0; s < MT; ++s) umma_f16(tmem + s * N, adesc[s] + 2 * k, bdesc + 2 * k, idesc, 1u);
      } else if (variant == 2) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int s = 0; s < MT; ++s) {
            if (s == 0) ws_b0_fill(tmem + s * N, adesc[s] + 2 * k, bdesc + 2 * k, idesc, 1u);
            else if (s == MT - 1) ws_b0_last(tmem + s * N, adesc[s] + 2 * k, bdesc + 2 * k, idesc, 1u);
            else ws_b0_use(tmem + s * N, adesc[s] + 2 * k, bdesc + 2 * k, idesc, 1u);
          }
      } else if (variant == 3) {
#pragma unroll
        for (int s = 0; s < MT; ++s) {
          if (s == 0) {
            ws_b0_fill(tmem + s * N, adesc[s] + 0, bdesc + 0, idesc, 1u);
            ws_b1_fill(tmem + s * N, adesc[s] + 2, bdesc + 2, idesc, 1u);
            ws_b2_fill(tmem + s * N, adesc[s] + 4, bdesc + 4, idesc, 1u);
            ws_b3_fill(tmem + s * N, adesc[s] + 6, bdesc + 6, idesc, 1u);
          } else if (s == MT - 1) {
            ws_b0_last(tmem + s * N, adesc[s] + 0, bdesc + 0, idesc, 1u);
            ws_b1_last(tmem + s * N, adesc[s] + 2, bdesc + 2, idesc, 1u);
            ws_b2_last(tmem + s * N, adesc[s] + 4, bdesc + 4, idesc, 1u);
            ws_b3_last(tmem + s * N, adesc[s] + 6, bdesc + 6, idesc, 1u);
          } else {
            ws_b0_use(tmem + s * N, adesc[s] + 0, bdesc + 0, idesc, 1u);
            ws_b1_use(tmem + s * N, adesc[s] + 2, bdesc + 2, idesc, 1u);
            ws_b2_use(tmem + s * N, adesc[s] + 4, bdesc + 4, idesc, 1u);
            ws_b3_use(tmem + s * N, adesc[s] + 6, bdesc + 6, idesc, 1u);
          }
        }
      } else if (variant == 4) {
#pragma unroll
        for (int s = 0; s < MT; ++s)
#pragma unroll
          for (int k = 0; k < 4; ++k) ws_b0_discard(tmem + s * N, adesc[s] + 2 * k, bdesc + 2 * k, idesc, 1u);
      } else {
        // A reuse: the same A slice against two different B tiles (two N tiles of one pixel tile); accumulators s and s + MT
        const uint64_t bdesc2 = umma_desc_sw128(b_base + (N <= 128 ? ((it + 1) & 3) * (N * 128) : 0));
#pragma unroll
        for (int s = 0; s < MT; ++s)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            mma_a_fill(tmem + s * N, adesc[s] + 2 * k, bdesc + 2 * k, idesc, 1u);
            mma_a_last(tmem + ((s + MT) * N) % 512, adesc[s] + 2 * k, bdesc2 + 2 * k, idesc, 1u);
          }
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_fence_after(), tmem_dealloc(tmem, 512);
}

// Numerics: D0 = plain form, D1 = .ws form (b0 fill/use/lastuse over MT sub-tiles), same pseudo-random operands; the 128
// threads of warps 0..3 read both accumulators back with tcgen05.ld (lane = row) and count mismatching words.
template <int N, int MT>
__global__ void __launch_bounds__(128, 1) check_kernel(int* mismatches, float* sample) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 2; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u;
    h ^= h >> 15;
    reinterpret_cast<uint16_t*>(smem)[i] = (uint16_t)((h & 0x807Fu) | 0x3F80u);   // +-[1, 2) in bf16
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = umma_idesc_bf16(128, N);
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
  constexpr uint64_t hi_a = ((uint64_t)((34 * 128) >> 4) << 32 | (1ull << 46) | (2ull << 61) | (1ull << 16));
  if (threadIdx.x == 32) {
    const uint64_t bdesc = umma_desc_sw128(b_base);
    uint64_t adesc[MT];
    for (int s = 0; s < MT; ++s) adesc[s] = hi_a | (uint64_t)(((a_base + (uint32_t)((34 + 8 * s + 1) * 128)) >> 4) & 0x3FFFu);
    for (int s = 0; s < MT; ++s)
      for (int k = 0; k < 4; ++k) umma_f16(tmem + s * N, adesc[s] + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
    for (int k = 0; k < 4; ++k)
      for (int s = 0; s < MT; ++s) {
        const uint32_t d = tmem + 256 + s * N;
        if (s == 0) ws_b0_fill(d, adesc[s] + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
        else if (s == MT - 1) ws_b0_last(d, adesc[s] + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
        else ws_b0_use(d, adesc[s] + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
      }
    umma_commit(smem_u32(&bar));
  }
  __syncthreads();
  mbar_wait(smem_u32(&bar), 0, nullptr, 0);
  tc_fence_after();
  int bad = 0;
  for (int c = 0; c < MT * N; c += 16) {
    uint32_t r0[16], r1[16];
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    tmem_ld16(tmem + lane_off + c, r0);
    tmem_ld16(tmem + lane_off + 256 + c, r1);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) bad += r0[j] != r1[j];
    if (c == 0 && threadIdx.x < 4) sample[threadIdx.x * 2] = __uint_as_float(r0[0]), sample[threadIdx.x * 2 + 1] = __uint_as_float(r1[0]);
  }
  atomicAdd(mismatches, bad);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_fence_after(), tmem_dealloc(tmem, 512);
}

// M = 64 form with the operand roles swapped (A = weights 64 x 16, B = N pixels x 16): rate only.
template <int N>
__global__ void __launch_bounds__(128, 1) rate64_kernel(long long* out, int iters, int a_tmem) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = umma_idesc_bf16(64, N);
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 32 * 1024);
  constexpr uint64_t hi_b = ((uint64_t)((34 * 128) >> 4) << 32 | (1ull << 46) | (2ull << 61) | (1ull << 16));
  if (warp == 1 && lane == 0) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int tap = it % 9;
      const uint64_t adesc = umma_desc_sw128(a_base + (it & 3) * (64 * 128));
      const uint32_t b_addr = b_base + (uint32_t)(((tap / 3) * 34 + tap % 3) * 128);
      const uint64_t bdesc = hi_b | (uint64_t)((b_addr >> 4) & 0x3FFFu);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (a_tmem) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
                       "r"(tmem + 256 + 8 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(1u)
                       : "memory");
        } else {
          umma_f16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
        }
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_fence_after(), tmem_dealloc(tmem, 512);
}

template <int N>
void run64(int a_tmem, long long* d) {
  const int iters = 2000;
  cudaFuncSetAttribute(rate64_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  rate64_kernel<N><<<148, 128, 200 * 1024>>>(d, iters, a_tmem);
  rate64_kernel<N><<<148, 128, 200 * 1024>>>(d, iters, a_tmem);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / ((double)iters * 4);
  printf("M=64 N=%3d A from %s: %6.1f clk/MMA -> %.0f MAC/clk/SM (full-rate floor %d clk)  (%s)\n", N, a_tmem ? "TMEM" : "smem", per,
         64.0 * N * 16 / per, N / 4, cudaGetErrorString(e));
}

template <int N, int MT>
void run(int variant, long long* d) {
  const int iters = 2000;
  cudaFuncSetAttribute(rate_kernel<N, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  rate_kernel<N, MT><<<148, 128, 200 * 1024>>>(d, iters, variant);
  rate_kernel<N, MT><<<148, 128, 200 * 1024>>>(d, iters, variant);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const int mmas = (variant == 5 ? 2 : 1) * MT * 4;
  const double per = (double)h / ((double)iters * mmas);
  static const char* names[] = {"plain s-outer", "plain k-outer", "ws b0 k-outer", "ws b0-3 s-outer", "ws discard", "plain A-collector x2N"};
  printf("N=%3d MT=%d %-22s: %6.1f clk/MMA -> %.0f MAC/clk/SM (floor %d clk)  (%s)\n", N, MT, names[variant], per, 128.0 * N * 16 / per,
         N / 2, cudaGetErrorString(e));
}

template <int N, int MT>
void check() {
  int* d;
  float* s;
  cudaMalloc(&d, 4);
  cudaMalloc(&s, 32);
  cudaMemset(d, 0, 4);
  cudaFuncSetAttribute(check_kernel<N, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  check_kernel<N, MT><<<1, 128, 200 * 1024>>>(d, s);
  cudaError_t e = cudaDeviceSynchronize();
  int h = -1;
  float hs[8];
  cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hs, s, 32, cudaMemcpyDeviceToHost);
  printf("check N=%d MT=%d: %d mismatching accumulator words of %d (%s); row0 plain %.3f ws %.3f, row1 %.3f %.3f\n", N, MT, h, 128 * MT * N,
         cudaGetErrorString(e), hs[0], hs[1], hs[2], hs[3]);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  check<64, 4>();
  check<128, 2>();
  for (int v = 0; v < 6; ++v) run<64, 4>(v, d);
  for (int v = 0; v < 6; ++v) run<128, 2>(v, d);
  for (int v : {0, 4}) run<256, 1>(v, d);
  for (int v : {0, 1, 2}) run<64, 2>(v, d);
  for (int v : {0, 1}) run<16, 4>(v, d);
  for (int v : {0, 1}) run<32, 4>(v, d);
  for (int v : {0, 1}) run<16, 2>(v, d);
  run<192, 2>(0, d);   // three horizontal taps stacked as one N = 192 MMA per A window
  run<128, 3>(0, d);
  run64<256>(0, d);
  run64<128>(0, d);
  run64<64>(0, d);
  run64<256>(1, d);
  run64<128>(1, d);
  return 0;
}
