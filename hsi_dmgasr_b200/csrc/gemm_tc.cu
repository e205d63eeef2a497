// Batched "NT" GEMM on tcgen05 for the self-attention contractions (unet.py:133-140):
//   C[b][m][n] = alpha * sum_k A[b][m][k] * B[b][n][k]        A, B bf16 K-major, fp32 accumulation in TMEM
// Both einsums of the reference and the V projection are expressed in this form (see attention() in unet.cu):
//   scores = Q K^T, V^T = W_v X^T (computed transposed so that the next contraction is K-major), O = P V.
// Same skeleton as conv_tc.cu: TMA producer warp, single-thread MMA issuer, 4 epilogue warps, double-buffered
// TMEM accumulator, persistent over (batch, m-tile, n-tile).
#include "tc_common.cuh"

namespace hsidm {
namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kSmemBudget = 196 * 1024;

struct GemmP {
  int M, N, K, batch;
  int m_tiles, n_tiles;
  int a_batched, b_batched;
  float alpha;
  void* C;
  long long ldc, sC;
  int c_f32;
  int row_softmax;
  int* err;
};

template <int BN>
struct GCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // as in conv_tc.cu: kGroup stages share one full barrier so the MMA issuer waits (and idles the pipe) half as often
  static constexpr int kGroup = 2;
  static constexpr int kGroups = ((kSmemBudget / kStageBytes) > 8 ? 8 : (kSmemBudget / kStageBytes)) / kGroup;
  static constexpr int kStages = kGroups * kGroup;
  static_assert(kGroups >= 2, "pipeline too shallow");
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kEpiBytes = 4 * 4096;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 + 256;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmP p) {
  using C = GCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_epi = smem + C::kStages * C::kStageBytes;
  uint8_t* tail = smem_epi + C::kEpiBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_batch = p.m_tiles * p.n_tiles;
  const int total_tiles = tiles_per_batch * p.batch;
  const int kblocks = (p.K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::kStages; ++s) mbar_init(smem_u32(&empty_bar[s]), 1);
    for (int g = 0; g < C::kGroups; ++g) mbar_init(smem_u32(&full_bar[g]), C::kGroup);
    for (int s = 0; s < 2; ++s) mbar_init(smem_u32(&tfull_bar[s]), 1), mbar_init(smem_u32(&tempty_bar[s]), 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);   // uniform register, see conv_halo.cu

  if (warp == 0) {
    if (lane == 0) {
      griddep_wait();     // both operands are activations of earlier kernels
      griddep_launch();
      int stage = 0, grp = 0, gcnt = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
        const int b = tile / tiles_per_batch, r = tile - b * tiles_per_batch;
        const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
        for (int kb = 0; kb < kblocks; ++kb) {
          ok = mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.err, 11);
          if (!ok) break;
          const uint32_t fb = smem_u32(&full_bar[grp]);
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          mbar_expect_tx(fb, C::kStageBytes);
          tma_load_3d(sa, &tmA, fb, kb * kBK, mt * kBM, p.a_batched ? b : 0);
          tma_load_3d(sa + C::kABytes, &tmB, fb, kb * kBK, nt * BN, p.b_batched ? b : 0);
          if (++stage == C::kStages) stage = 0, phase ^= 1;
          if (++gcnt == C::kGroup) gcnt = 0, grp = grp + 1 == C::kGroups ? 0 : grp + 1;
        }
      }
      for (; ok && gcnt != 0 && gcnt < C::kGroup; ++gcnt) mbar_arrive(smem_u32(&full_bar[grp]));   // complete the last group
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
      int stage = 0, acc = 0, grp = 0, gcnt = 0;
      uint32_t gph = 0, acc_phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
        ok = mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1, p.err, 12);
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          if (gcnt == 0) {
            ok = mbar_wait(smem_u32(&full_bar[grp]), gph, p.err, 13);
            if (!ok) break;
            tc_fence_after();
          }
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + C::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == C::kStages) stage = 0;
          if (++gcnt == C::kGroup) {
            gcnt = 0;
            if (++grp == C::kGroups) grp = 0, gph ^= 1;
          }
        }
        umma_commit(smem_u32(&tfull_bar[acc]));
        if (++acc == 2) acc = 0, acc_phase ^= 1;
      }
    }
  } else {
    griddep_wait();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool ok = true;
    uint4* stage = reinterpret_cast<uint4*>(smem_epi + (warp - 2) * 4096);
    for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
      const int b = tile / tiles_per_batch, r = tile - b * tiles_per_batch;
      const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
      ok = mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase, p.err, 14);
      if (!ok) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
      const int m = mt * kBM + row;
      if (p.row_softmax) {
        // The tile holds whole rows (n_tiles == 1) and a thread owns one row: three passes over its accumulator row in
        // tensor memory - max, sum of exp, normalised bf16 probabilities - and the scores are never written anywhere.
        const float sc = p.alpha * 1.4426950408889634f;   // softmax(alpha*s) = exp2((s - max) * alpha * log2 e) / sum
        float mx = -INFINITY;
#pragma unroll 1
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
        float sum = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) sum += exp2f((__uint_as_float(v[j]) - mx) * sc);
        }
        const float inv = 1.0f / sum;
        bf16* crow = static_cast<bf16*>(p.C) + b * p.sC + (long long)m * p.ldc;
#pragma unroll 1
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          uint4 o[2];
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            h[j] = __floats2bfloat162_rn(exp2f((__uint_as_float(v[2 * j]) - mx) * sc) * inv,
                                         exp2f((__uint_as_float(v[2 * j + 1]) - mx) * sc) * inv);
          if (m < p.M) {
            *reinterpret_cast<uint4*>(crow + c0) = o[0];
            *reinterpret_cast<uint4*>(crow + c0 + 8) = o[1];
          }
        }
      } else if (p.c_f32) {
        float* crow = static_cast<float*>(p.C) + b * p.sC + (long long)m * p.ldc + nt * BN;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (m < p.M) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              if (nt * BN + c0 + j < p.N)
                *reinterpret_cast<float4*>(crow + c0 + j) =
                    make_float4(__uint_as_float(v[j]) * p.alpha, __uint_as_float(v[j + 1]) * p.alpha,
                                __uint_as_float(v[j + 2]) * p.alpha, __uint_as_float(v[j + 3]) * p.alpha);
            }
          }
        }
      } else {
        // bf16 output through the coalesced staged epilogue (rows = matrix rows, "Cout" = ldc)
        EpiP e;
        e.N_img = 1, e.H = 1, e.W = 1, e.Cout = (int)p.ldc;
        e.bias = nullptr, e.nbias = nullptr, e.nbs = 0, e.nb_t = nullptr, e.nb_ts = 0;
        e.act = ACT_NONE, e.scale = p.alpha, e.resid = nullptr;
        e.out = static_cast<bf16*>(p.C) + b * p.sC, e.out_layout = L_NHWC, e.clamp01 = 0;
        e.stats = nullptr, e.stats_slots = 0;
        float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
        const int m0 = mt * kBM, M = p.M;
        auto pix = [&](int R, int& pn, long long& pm) {
          pn = (m0 + R) < M ? 0 : 1;   // rows past M are masked like images past N_img
          pm = m0 + R;
        };
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64)
          if (nt * BN + c0 < p.N) epilogue_rows64(e, nullptr, taddr + c0, quarter, lane, nt * BN + c0, stage, pix, st);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      if (++acc == 2) acc = 0, acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

int encode_3d(CUtensorMap* map, const void* base, int K, int rows, int batch, long long ld, long long stride, int box_rows) {
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? stride : (long long)rows * ld) * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = host().encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    HSIDM_FAIL(HSIDM_CUDA_ERROR, "cuTensorMapEncodeTiled(gemm operand %dx%dx%d ld %lld) failed: %d", batch, rows, K, ld, (int)r);
  return HSIDM_OK;
}

template <int BN>
int launch(const GemmTcOp& op, cudaStream_t stream) {
  using C = GCfg<BN>;
  GemmP p;
  p.M = op.M, p.N = op.N, p.K = op.K, p.batch = op.batch;
  p.m_tiles = (int)ceil_div(op.M, kBM), p.n_tiles = (int)ceil_div(op.N, BN);
  p.a_batched = op.sA != 0, p.b_batched = op.sB != 0;
  p.alpha = op.alpha, p.C = op.C, p.ldc = op.ldc, p.sC = op.sC, p.c_f32 = op.c_f32, p.row_softmax = op.row_softmax, p.err = host().err_flag;
  CUtensorMap tmA, tmB;
  HSIDM_TRY(encode_3d(&tmA, op.A, op.K, op.M, p.a_batched ? op.batch : 1, op.lda, op.sA, kBM));
  HSIDM_TRY(encode_3d(&tmB, op.B, op.K, op.N, p.b_batched ? op.batch : 1, op.ldb, op.sB, BN));
  const int grid = std::min(p.m_tiles * p.n_tiles * p.batch, host().num_sms);
  char tag[96];
  snprintf(tag, sizeof(tag), "gemm_tc BN%d m%d n%d k%d b%d%s", BN, op.M, op.N, op.K, op.batch, op.row_softmax ? " +softmax" : "");
  ProfScope prof(PROF_GEMM, 2.0 * op.M * (double)op.N * op.K * op.batch, stream, tag);
  HSIDM_CUDA(launch_pdl(gemm_tc_kernel<BN>, dim3(grid), dim3(kThreads), C::kSmemBytes, stream, 1, tmA, tmB, p));
  return after_launch("gemm_tc_kernel");
}

}  // namespace

int gemm_tc_init() {
  HSIDM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GCfg<64>::kSmemBytes));
  HSIDM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GCfg<128>::kSmemBytes));
  HSIDM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GCfg<256>::kSmemBytes));
  return HSIDM_OK;
}

bool gemm_tc_supported(const GemmTcOp& op) {
  if (op.K % 8 || op.lda % 8 || op.ldb % 8 || op.sA % 8 || op.sB % 8) return false;   // 16-byte TMA strides
  if (op.N % 64) return false;
  if (!op.c_f32 && (op.ldc % 8 || op.sC % 8)) return false;
  // row softmax: whole rows in one tile (N is exactly one of the tile widths); the max is taken before scaling by alpha
  if (op.row_softmax && (op.c_f32 || !(op.N == 64 || op.N == 128 || op.N == 256) || op.alpha <= 0.f)) return false;
  if (op.c_f32 && (op.ldc % 4 || op.sC % 4)) return false;
  return true;
}

int gemm_tc(const GemmTcOp& op, cudaStream_t stream) {
  HSIDM_TRY(conv_tc_init());
  if (!gemm_tc_supported(op))
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "gemm_tc: M=%d N=%d K=%d lda=%lld ldb=%lld does not fit the tensor-core GEMM", op.M, op.N, op.K,
               (long long)op.lda, (long long)op.ldb);
  if (op.N % 256 == 0) return launch<256>(op, stream);
  if (op.N % 128 == 0) return launch<128>(op, stream);
  return launch<64>(op, stream);
}

}  // namespace hsidm
