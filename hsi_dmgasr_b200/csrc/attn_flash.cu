// Flash-style attention core on tcgen05 for the low-resolution UNet stages (unet.py:133-140, n_head = 1, S <= 256 tokens):
//   Y[b] = softmax(alpha * Q[b] K[b]^T) V[b]         Q, K: [S][Ck] bf16 row-major, V given TRANSPOSED as Vt [Cv][S]
// in ONE kernel: the scores never leave tensor memory, the probabilities never leave shared memory.
// (With the projections folded at commit - unet.cu attention_folded - Q = Xn Mqk^T, K = Xn, Vt = Xn^T.)
//
// Work item = (image b, tile of 128 queries).  Per item, in order:
//   1. scores  S[128 x S] = Q_tile K^T        K-loop over Ck in 64-wide blocks, M = 128, N = S, accumulators in TMEM columns [0, S)
//   2. softmax per row (thread = row): max, sum of exp2, bf16 probabilities written as the K-major SWIZZLE_128B A operand
//      of step 3 into shared memory (S/64 blocks of 128 rows x 128 B)
//   3. output  O[128 x Cv] = P Vt^T           in slabs of 256 output channels, K-loop over the S keys; slab h accumulates in
//      TMEM columns [256, 512) (h even) or [0, 256) (h odd: the scores are dead once P is in shared memory)
//   4. each slab -> bf16 rows of Y
// Warp roles (192 threads): warp 0 = TMA producer (one ring of 48 KB stages: Q block + K block in step 1, a Vt block in
// step 3), warp 1 = MMA issuer, warps 2..5 = softmax + epilogue, one per TMEM lane quarter.
// Same arithmetic, same K order as the two-kernel form (gemm_tc with row_softmax, then gemm_tc): bitwise the same Y.
#include "tc_common.cuh"

namespace hsidm {
namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kSlab = 256;                       // output channels per accumulation slab
constexpr int kQBytes = kBM * kBK * 2;           // 16 KB: 128 queries x 64 channels
constexpr int kKBytes = 256 * kBK * 2;           // 32 KB: up to 256 keys x 64 channels, or 256 output channels x 64 keys
constexpr int kStageBytes = kQBytes + kKBytes;   // 48 KB
constexpr int kStages = 3;
constexpr int kPBytes = kBM * 256 * 2;           // 64 KB: probabilities of one tile, 4 k-blocks
constexpr int kSmemBytes = kStages * kStageBytes + kPBytes + 1024 + 256;

struct FlashP {
  int S, Ck, Cv, batch, m_tiles;
  float alpha;
  bf16* Y;
  long long ldy, sY;
  int* err;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
attn_flash_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const FlashP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_p = smem + kStages * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_p + kPBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* s_full = empty_bar + kStages;     // scores complete
  uint64_t* p_ready = s_full + 1;             // probabilities in shared memory (4 warps arrive)
  uint64_t* p_empty = p_ready + 1;            // the P.V MMAs of the tile have read them
  uint64_t* o_full = p_empty + 1;             // [2] slab in TMEM region 0 / 1 complete
  uint64_t* o_empty = o_full + 2;             // [2] region drained by the epilogue (4 warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.batch * p.m_tiles;
  const int kb_qk = p.Ck / kBK, kb_pv = p.S / kBK, slabs = p.Cv / kSlab;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&full_bar[s]), 1), mbar_init(smem_u32(&empty_bar[s]), 1);
    mbar_init(smem_u32(s_full), 1);
    mbar_init(smem_u32(p_ready), 4);
    mbar_init(smem_u32(p_empty), 1);
    for (int r = 0; r < 2; ++r) mbar_init(smem_u32(&o_full[r]), 1), mbar_init(smem_u32(&o_empty[r]), 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);   // uniform register, see conv_halo.cu

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      griddep_wait();
      griddep_launch();
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      auto next = [&]() { if (++stage == kStages) stage = 0, phase ^= 1; };
      for (int it = blockIdx.x; it < items && ok; it += gridDim.x) {
        const int b = it / p.m_tiles, mt = it - b * p.m_tiles;
        for (int kb = 0; kb < kb_qk && ok; ++kb) {   // step 1: Q block + K block
          ok = mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.err, 21);
          if (!ok) break;
          const uint32_t fb = smem_u32(&full_bar[stage]), sa = smem_u32(smem + stage * kStageBytes);
          mbar_expect_tx(fb, kQBytes + p.S * kBK * 2);
          tma_load_3d(sa, &tmQ, fb, kb * kBK, mt * kBM, b);
          tma_load_3d(sa + kQBytes, &tmK, fb, kb * kBK, 0, b);
          next();
        }
        for (int h = 0; h < slabs && ok; ++h)        // step 3: Vt blocks, slab by slab
          for (int kb = 0; kb < kb_pv && ok; ++kb) {
            ok = mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.err, 22);
            if (!ok) break;
            const uint32_t fb = smem_u32(&full_bar[stage]), sa = smem_u32(smem + stage * kStageBytes);
            mbar_expect_tx(fb, kKBytes);
            tma_load_3d(sa + kQBytes, &tmV, fb, kb * kBK, h * kSlab, b);
            next();
          }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(kBM, p.S);   // N = S for the scores
      constexpr uint32_t idesc_o = umma_idesc_bf16(kBM, kSlab);
      int stage = 0, used[2] = {0, 0};   // used[r]: output slabs produced into TMEM region r so far
      uint32_t phase = 0;
      bool ok = true;
      auto next = [&]() { if (++stage == kStages) stage = 0, phase ^= 1; };
      // region r may be overwritten once the epilogue has drained the slab last produced into it
      auto wait_drained = [&](int r, int code) {
        if (used[r] > 0) {
          ok = mbar_wait(smem_u32(&o_empty[r]), (uint32_t)((used[r] - 1) & 1), p.err, code);
          tc_fence_after();
        }
      };
      int n_item = 0;
      for (int it = blockIdx.x; it < items && ok; it += gridDim.x, ++n_item) {
        wait_drained(0, 23);   // the scores go to region 0 (columns [0, 256)), where the last odd slab lived
        if (!ok) break;
        for (int kb = 0; kb < kb_qk && ok; ++kb) {
          ok = mbar_wait(smem_u32(&full_bar[stage]), phase, p.err, 24);
          if (!ok) break;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + kQBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc_s, (kb | k) ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[stage]));
          next();
        }
        if (!ok) break;
        umma_commit(smem_u32(s_full));
        // the softmax warps turn the scores into P (shared memory) and thereby free region 0
        ok = mbar_wait(smem_u32(p_ready), (uint32_t)(n_item & 1), p.err, 25);
        if (!ok) break;
        tc_fence_after();
        for (int h = 0; h < slabs && ok; ++h) {
          const int reg = (h & 1) ^ 1;   // slab 0 -> region 1 (columns [256, 512)), slab 1 -> region 0, ...
          if (h != 1) wait_drained(reg, 26);   // slab 1 follows the scores in region 0: drained before them, consumed by the softmax
          if (!ok) break;
          const uint32_t d = tmem_base + reg * kSlab;
          for (int kb = 0; kb < kb_pv && ok; ++kb) {
            ok = mbar_wait(smem_u32(&full_bar[stage]), phase, p.err, 27);
            if (!ok) break;
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(smem_u32(smem_p + kb * kQBytes));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + stage * kStageBytes) + kQBytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc_o, (kb | k) ? 1u : 0u);
            umma_commit(smem_u32(&empty_bar[stage]));
            next();
          }
          if (!ok) break;
          umma_commit(smem_u32(&o_full[reg]));
          ++used[reg];
        }
        if (!ok) break;
        umma_commit(smem_u32(p_empty));   // all P.V MMAs of the tile have read the probabilities
      }
    }
  } else {
    // =============================== softmax + epilogue (warps 2..5) ===============================
    griddep_wait();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    int got[2] = {0, 0};   // got[r]: output slabs drained from TMEM region r so far
    bool ok = true;
    int n_item = 0;
    for (int it = blockIdx.x; it < items && ok; it += gridDim.x, ++n_item) {
      const int b = it / p.m_tiles, mt = it - b * p.m_tiles;
      const int m = mt * kBM + row;
      ok = mbar_wait(smem_u32(s_full), (uint32_t)(n_item & 1), p.err, 28);
      if (!ok) break;
      tc_fence_after();
      // three passes over this thread's score row in tensor memory: max, sum of exp2, normalised bf16 probabilities
      const uint32_t taddr = tmem_base + lane_addr;
      const float sc = p.alpha * 1.4426950408889634f;
      float mx = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < p.S; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      float sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < p.S; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) sum += exp2f((__uint_as_float(v[j]) - mx) * sc);
      }
      const float inv = 1.0f / sum;
      // the previous tile's P.V MMAs must have read the probabilities before they are overwritten
      if (n_item > 0) {
        ok = mbar_wait(smem_u32(p_empty), (uint32_t)((n_item - 1) & 1), p.err, 29);
        if (!ok) break;
      }
      const uint32_t prow = smem_u32(smem_p) + (uint32_t)(row * 128);
#pragma unroll 1
      for (int c0 = 0; c0 < p.S; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        uint4 o[2];
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          h[j] = __floats2bfloat162_rn(exp2f((__uint_as_float(v[2 * j]) - mx) * sc) * inv,
                                       exp2f((__uint_as_float(v[2 * j + 1]) - mx) * sc) * inv);
        // keys c0 .. c0+15 = 16-byte chunks (c0 % 64) / 8 and + 1 of k-block c0 / 64; 128B swizzle by row
        const uint32_t blk = prow + (uint32_t)((c0 >> 6) * kQBytes);
        const int j0 = (c0 & 63) >> 3;
        sts128(blk + (uint32_t)(((j0 ^ (row & 7)) << 4)), o[0]);
        sts128(blk + (uint32_t)((((j0 + 1) ^ (row & 7)) << 4)), o[1]);
      }
      fence_proxy_async_smem();   // the MMAs read P through the async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(p_ready));   // P is ready AND region 0 (the scores) is free
      // drain the output slabs
      for (int h = 0; h < slabs && ok; ++h) {
        const int reg = (h & 1) ^ 1;
        ok = mbar_wait(smem_u32(&o_full[reg]), (uint32_t)(got[reg] & 1), p.err, 30);
        if (!ok) break;
        ++got[reg];
        tc_fence_after();
        bf16* yrow = p.Y + b * p.sY + (long long)m * p.ldy + h * kSlab;
#pragma unroll 1
        for (int c0 = 0; c0 < kSlab; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + reg * kSlab + lane_addr + c0, v);
          tmem_ld_wait();
          uint4 o[2];
          __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
          for (int j = 0; j < 8; ++j) hh[j] = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          if (m < p.S) {
            *reinterpret_cast<uint4*>(yrow + c0) = o[0];
            *reinterpret_cast<uint4*>(yrow + c0 + 8) = o[1];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&o_empty[reg]));   // the region's next user (a slab, or the next tile's scores) may write
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int encode_3d(CUtensorMap* map, const void* base, int K, int rows, int batch, long long ld, long long stride, int box_rows) {
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = host().encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    HSIDM_FAIL(HSIDM_CUDA_ERROR, "cuTensorMapEncodeTiled(attention operand %dx%dx%d ld %lld) failed: %d", batch, rows, K, ld, (int)r);
  return HSIDM_OK;
}

}  // namespace

bool attn_flash_supported(const AttnFlashOp& op) {
  if (!(op.S == 64 || op.S == 128 || op.S == 256)) return false;        // a score row is one accumulator tile
  if (op.Ck <= 0 || op.Ck % kBK || op.Cv <= 0 || op.Cv % kSlab) return false;
  if (op.ldq % 8 || op.ldk % 8 || op.ldy % 8 || op.sQ % 8 || op.sK % 8 || op.sVt % 8 || op.sY % 8) return false;
  return op.alpha > 0.f && op.batch > 0;
}

int attn_flash_init() {
  HSIDM_CUDA(cudaFuncSetAttribute(attn_flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  return HSIDM_OK;
}

int attn_flash(const AttnFlashOp& op, cudaStream_t stream) {
  HSIDM_TRY(conv_tc_init());
  if (!attn_flash_supported(op))
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "attn_flash: S=%d Ck=%d Cv=%d does not fit the fused attention kernel", op.S, op.Ck, op.Cv);
  FlashP p;
  p.S = op.S, p.Ck = op.Ck, p.Cv = op.Cv, p.batch = op.batch, p.m_tiles = (int)ceil_div(op.S, kBM);
  p.alpha = op.alpha, p.Y = static_cast<bf16*>(op.Y), p.ldy = op.ldy, p.sY = op.sY, p.err = host().err_flag;
  CUtensorMap tmQ, tmK, tmV;
  HSIDM_TRY(encode_3d(&tmQ, op.Q, op.Ck, op.S, op.batch, op.ldq, op.sQ, kBM));
  HSIDM_TRY(encode_3d(&tmK, op.K, op.Ck, op.S, op.batch, op.ldk, op.sK, op.S));
  HSIDM_TRY(encode_3d(&tmV, op.Vt, op.S, op.Cv, op.batch, op.S, op.sVt, kSlab));
  const int grid = std::min(p.batch * p.m_tiles, host().num_sms);
  char tag[96];
  snprintf(tag, sizeof(tag), "attn_flash s%d ck%d cv%d b%d", op.S, op.Ck, op.Cv, op.batch);
  ProfScope prof(PROF_GEMM, 2.0 * op.S * (double)op.S * (op.Ck + op.Cv) * op.batch, stream, tag);
  HSIDM_CUDA(launch_pdl(attn_flash_kernel, dim3(grid), dim3(kThreads), kSmemBytes, stream, 1, tmQ, tmK, tmV, p));
  return after_launch("attn_flash_kernel");
}

}  // namespace hsidm
