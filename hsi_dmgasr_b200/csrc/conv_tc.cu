// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// GEMM view of a 3x3 (pad 1) or 1x1 stride-1 convolution over NHWC activations:
//   D[m, co] = sum_{tap, c} X[n, y+ky-1, x+kx-1, c] * Wt[co, tap*Ctot + c]       m = (n, y, x)
// A tile  : 128 output pixels x 64 channels of ONE tap = one 4-D TMA box {64, bw, bh, bn} (bw*bh*bn = 128) of the
//           NHWC tensor at coordinates shifted by the tap; out-of-bounds rows/cols are zero-filled by the TMA unit,
//           which is exactly the conv padding.  No im2col buffer exists anywhere.  Rows land in shared memory as
//           128-byte lines with the 128B swizzle -> canonical K-major SWIZZLE_128B UMMA operand.
// B tile  : BN output channels x 64 k of the pre-packed K-major weight matrix, one 2-D TMA box.
// MMA     : tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16 (x4 per 64-wide k-block), issued by one thread.
// D       : fp32 in TMEM, double buffered (2*BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Epilogue: 4 warps, tcgen05.ld 32x32b -> +bias +noise-embedding bias -> activation -> *scale -> +residual ->
//           bf16 NHWC (or fp32 NCHW for the last layer) straight to global memory.
// Two NHWC sources may be given; their channels are consumed as consecutive K ranges (virtual torch.cat, unet.py:259).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
// Persistent: grid = min(#tiles, #SMs); tiles are walked n-tile-fastest so the CTAs in flight share A tiles in L2.
#include <cstdlib>
#include <mutex>

#include "tc_common.cuh"

namespace hsidm {
namespace {

using namespace tc;

constexpr int kEpiWarps = 8;      // two per TMEM lane quarter
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemBudget = 196 * 1024;

struct TcP {
  int M;
  int bw, bh, bn;          // box extents: x, y, image
  int tiles_x, tiles_y;    // tiles per image along x / y (bn == 1)
  int m_tiles, n_tiles;
  int taps;                // 1 or 9
  int bw_shift, ppi_shift; // log2(bw), log2(bw*bh): tile extents are powers of two below 128
  int chunks0, chunks1;    // 64-channel chunks taken from source 0 / source 1
  EpiP e;
  int* err;                // device flag set when a barrier wait times out (never in a healthy run)
};

template <int BN>
struct Cfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // Stages fill in groups of kGroup that share one full barrier: each barrier wait of the MMA issuer idles the tensor
  // pipe for ~160 clk (scripts/mma_rate.cu), so it waits once per group; stages are still freed one by one.
  static constexpr int kGroup = 2;
  static constexpr int kGroups = ((kSmemBudget / kStageBytes) > 8 ? 8 : (kSmemBudget / kStageBytes)) / kGroup;
  static constexpr int kStages = kGroups * kGroup;
  static_assert(kGroups >= 2, "pipeline too shallow");
  static constexpr int kTmemCols = (2 * BN) < 32 ? 32 : 2 * BN;
  static constexpr int kEpiBytes = BN % 64 == 0 ? kEpiWarps * 4096 : 0;   // staging for the coalesced epilogue
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO, const __grid_constant__ TcP p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_epi = smem + C::kStages * C::kStageBytes;
  uint8_t* tail = smem_epi + C::kEpiBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;           // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int kblocks = p.taps * (p.chunks0 + p.chunks1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.chunks1) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < C::kStages; ++s) mbar_init(smem_u32(&empty_bar[s]), 1);
    for (int g = 0; g < C::kGroups; ++g) mbar_init(smem_u32(&full_bar[g]), C::kGroup);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), kEpiWarps);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);   // uniform register, see conv_halo.cu

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      griddep_wait();     // activations come from the previous kernel(s) of the stream
      griddep_launch();
      int stage = 0, grp = 0, gcnt = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
        int n0, y0, x0;
        if (p.bn > 1) {
          n0 = mt * p.bn, y0 = 0, x0 = 0;
        } else {
          const int tpi = p.tiles_x * p.tiles_y;
          n0 = mt / tpi;
          const int r = mt - n0 * tpi;
          y0 = (r / p.tiles_x) * p.bh;
          x0 = (r % p.tiles_x) * p.bw;
        }
        int kb = 0;
        for (int tap = 0; tap < p.taps && ok; ++tap) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
          for (int ch = 0; ch < p.chunks0 + p.chunks1; ++ch, ++kb) {
            ok = mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.err, 1);
            if (!ok) break;
            const uint32_t fb = smem_u32(&full_bar[grp]);
            const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
            mbar_expect_tx(fb, C::kStageBytes);
            if (ch < p.chunks0)
              tma_load_4d(sa, &tmA0, fb, ch * kBK, x0 + dx, y0 + dy, n0);
            else
              tma_load_4d(sa, &tmA1, fb, (ch - p.chunks0) * kBK, x0 + dx, y0 + dy, n0);
            tma_load_2d(sa + C::kABytes, &tmB, fb, kb * kBK, nt * BN);
            if (++stage == C::kStages) stage = 0, phase ^= 1;
            if (++gcnt == C::kGroup) gcnt = 0, grp = grp + 1 == C::kGroups ? 0 : grp + 1;
          }
        }
      }
      // the issuer waits for whole groups: complete the last, partly filled one
      for (; ok && gcnt != 0 && gcnt < C::kGroup; ++gcnt) mbar_arrive(smem_u32(&full_bar[grp]));
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
      int stage = 0, grp = 0, gcnt = 0;
      uint32_t gph = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
        ok = mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1, p.err, 2);
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          if (gcnt == 0) {
            ok = mbar_wait(smem_u32(&full_bar[grp]), gph, p.err, 3);
            if (!ok) break;
            tc_fence_after();
          }
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + C::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)  // +32 bytes (encoded +2) per K=16 slice inside the swizzle line
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
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
    // =============================== epilogue (warps 2..5) ===============================
    griddep_wait();
    const int quarter = warp & 3;            // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;        // which of the two warps of this quarter
    const int row = quarter * 32 + lane;     // accumulator row = pixel within the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    const float* nbias = p.e.nbias ? p.e.nbias + (p.e.nb_t ? (long long)(*p.e.nb_t) * p.e.nb_ts : 0) : nullptr;
    const int tpi = p.tiles_x * p.tiles_y;
    const int bw_mask = p.bw - 1, ppi_mask = (1 << p.ppi_shift) - 1;
    bool ok = true;
    for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
      const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
      // tile origin: image n0 and pixel (y0, x0); rows then decode with shifts only
      int n0, y0, x0;
      if (p.bn > 1) {
        n0 = mt * p.bn, y0 = 0, x0 = 0;
      } else {
        n0 = mt / tpi;
        const int q = mt - n0 * tpi;
        y0 = (q / p.tiles_x) * p.bh, x0 = (q % p.tiles_x) * p.bw;
      }
      auto pix = [&](int R, int& pn, long long& pm) {
        pn = n0 + (R >> p.ppi_shift);
        const int q = R & ppi_mask;
        pm = ((long long)pn * p.e.H + y0 + (q >> p.bw_shift)) * p.e.W + x0 + (q & bw_mask);
      };
      ok = mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase, p.err, 4);
      if (!ok) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
      bool staged = false;
      if constexpr (BN % 64 == 0) {
        if (p.e.out_layout == L_NHWC) {
          staged = true;
          const uint32_t stage = smem_u32(smem_epi + (warp - 2) * 4096);
          // A tile's 128 accumulator rows are 128 CONSECUTIVE pixels of the NHWC tensor (x fastest, then y, then image;
          // tiles span the full width whenever W < 128), so row R of this tile is flat pixel m_tile + R.
          const long long m_tile = ((long long)n0 * p.e.H + y0) * p.e.W + x0;
          const long long m_warp = m_tile + quarter * 32;
          const long long m_total = (long long)p.e.N_img * p.e.H * p.e.W;
          const int valid_rows = (int)max(0LL, min(32LL, m_total - m_warp));
          // statistics slot of this warp's 32 rows (all of one image, see pertap_stats_slots)
          int sn, sslot;
          if (p.bn > 1) {
            const int qpi = 4 / p.bn;             // lane quarters per image
            sn = n0 + quarter / qpi, sslot = quarter % qpi;
          } else {
            sn = n0, sslot = (mt - n0 * tpi) * 4 + quarter;
          }
          // a tile of several small images may reach past the batch: those rows are masked (valid_rows / stats_store),
          // but their noise-bias row must not be read - it does not exist
          const float* cb = nbias ? nbias + (long long)min(sn, p.e.N_img - 1) * p.e.nbs : p.e.bias;
          const float* cb2 = nbias ? p.e.bias : nullptr;
          constexpr int nC = BN / 64;
          // the two warps of a quarter alternate over the 64-channel chunks (a single chunk goes to the first warp)
#pragma unroll 1
          for (int ci = half; ci < nC; ci += 2) {
            const int co0 = nt * BN + ci * 64;
            if (co0 >= p.e.Cout) break;
            float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bn <= 4) {
              // this lane's (row sub, 16-byte chunk) element of the warp's 32 consecutive pixels in the residual tensor
              const bf16* resid_lane =
                  p.e.resid ? p.e.resid + (m_warp + (lane >> 3)) * p.e.Cout + co0 + (lane & 7) * 8 : nullptr;
              epilogue_tma64<2>(p.e, &tmO, 0u, taddr + ci * 64, lane, stage, co0, (int)m_warp, 0, 0, resid_lane, 8LL * p.e.Cout,
                                4LL * p.e.Cout, st, cb, cb2, valid_rows);
              if (p.e.stats) stats_store(p.e, sn, sslot, co0, lane, st);
            } else {
              // images smaller than 32 pixels: a warp's rows span several images, so the noise bias is per row
              epilogue_rows64(p.e, nbias, taddr + ci * 64, quarter, lane, co0, reinterpret_cast<uint4*>(smem_epi + (warp - 2) * 4096),
                              pix, st);
            }
          }
        }
      }
      if (!staged && half == 0) {
        int n;
        long long m;
        pix(row, n, m);
        const int q = row & ppi_mask;
        const int y = y0 + (q >> p.bw_shift), x = x0 + (q & bw_mask);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
          const int co0 = nt * BN + c0;
          if (n < p.e.N_img && co0 < p.e.Cout) epilogue16(p.e, nbias, n, y, x, co0, r);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      if (++acc == 2) acc = 0, acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all();   // the staging tiles must outlive the stores reading them
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
std::once_flag g_once;
int g_init_status = HSIDM_OK;

int do_init() {
  Host& h = host();
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  HSIDM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess)
    HSIDM_FAIL(HSIDM_CUDA_ERROR, "cuTensorMapEncodeTiled is not available from this driver");
  h.encode = reinterpret_cast<EncodeTiledFn>(fn);
  int dev = 0;
  HSIDM_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  HSIDM_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "tensor-core path needs sm_100 (found sm_%d%d)", prop.major, prop.minor);
  h.num_sms = prop.multiProcessorCount;
  HSIDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<16>::kSmemBytes));
  HSIDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::kSmemBytes));
  HSIDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::kSmemBytes));
  HSIDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::kSmemBytes));
  HSIDM_CUDA(cudaMalloc(&h.err_flag, sizeof(int)));
  HSIDM_CUDA(cudaMemset(h.err_flag, 0, sizeof(int)));
  HSIDM_TRY(conv_halo_init());
  HSIDM_TRY(attn_flash_init());
  return gemm_tc_init();
}

bool tile_geometry(int H, int W, int* bw, int* bh, int* bn) {
  if (W >= kBM) {
    if (W % kBM) return false;
    *bw = kBM, *bh = 1, *bn = 1;
    return true;
  }
  if (W <= 0 || (kBM % W)) return false;
  *bw = W;
  int rows = kBM / W;
  if (H >= rows) {
    if (H % rows) return false;
    *bh = rows, *bn = 1;
    return true;
  }
  if (rows % H) return false;
  *bh = H, *bn = rows / H;
  return true;
}

}  // namespace

namespace tc {

Host& host() {
  static Host h;
  return h;
}

int encode_act_map(CUtensorMap* map, const void* base, int N, int H, int W, int C, int bw, int bh, int bn) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = host().encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) HSIDM_FAIL(HSIDM_CUDA_ERROR, "cuTensorMapEncodeTiled(activation %dx%dx%dx%d) failed: %d", N, H, W, C, (int)r);
  return HSIDM_OK;
}

int encode_rows_map(CUtensorMap* map, void* base, long long pixels, int C) {
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)pixels};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, 32};
  cuuint32_t es[2] = {1, 1};
  CUresult r = host().encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) HSIDM_FAIL(HSIDM_CUDA_ERROR, "cuTensorMapEncodeTiled(output rows %lld x %d) failed: %d", pixels, C, (int)r);
  return HSIDM_OK;
}

int encode_phase_map(CUtensorMap* map, const void* base, int N, int H, int W, int C, int py, int px, int bw, int bh) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)(W / 2), (cuuint64_t)(H / 2), (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)2 * C * 2, (cuuint64_t)2 * W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const void* origin = static_cast<const bf16*>(base) + ((long long)py * W + px) * C;
  CUresult r = host().encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(origin), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) HSIDM_FAIL(HSIDM_CUDA_ERROR, "cuTensorMapEncodeTiled(phase %d,%d of %dx%dx%dx%d) failed: %d", py, px, N, H, W, C, (int)r);
  return HSIDM_OK;
}

int encode_out_map(CUtensorMap* map, void* base, int N, int H, int W, int C, int scale, int oy, int ox) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)(W / scale), (cuuint64_t)(H / scale), (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)scale * C * 2, (cuuint64_t)scale * W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBK, 8, 4, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  void* origin = static_cast<bf16*>(base) + ((long long)oy * W + ox) * C;
  CUresult r = host().encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, origin, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) HSIDM_FAIL(HSIDM_CUDA_ERROR, "cuTensorMapEncodeTiled(output %dx%dx%dx%d) failed: %d", N, H, W, C, (int)r);
  return HSIDM_OK;
}

int encode_weight_map(CUtensorMap* map, const void* base, int K, int rows, int bn_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)bn_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = host().encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) HSIDM_FAIL(HSIDM_CUDA_ERROR, "cuTensorMapEncodeTiled(weights %dx%d) failed: %d", rows, K, (int)r);
  return HSIDM_OK;
}

void fill_epilogue(EpiP* e, const ConvOp& op) {
  e->N_img = op.N, e->H = op.Hout, e->W = op.Wout, e->Cout = op.Cout;
  e->bias = op.bias, e->nbias = op.nbias, e->nbs = op.nbias_stride, e->nb_t = op.nbias_t, e->nb_ts = op.nbias_t_stride;
  e->act = op.act, e->scale = op.scale, e->resid = static_cast<const bf16*>(op.resid), e->out = op.out;
  e->out_layout = op.out_layout, e->clamp01 = op.clamp01;
  e->stats = op.stats_out, e->stats_slots = op.stats_slots;
}

int pertap_stats_slots(const ConvOp& op) {
  int bw, bh, bn;
  if (op.out_layout != L_NHWC || op.Cout % 64 || !tile_geometry(op.Hin, op.Win, &bw, &bh, &bn)) return 0;
  if (bn > 1) return (bn <= 4) ? 4 / bn : 0;     // a warp's 32 rows must stay inside one image
  return (op.Win / bw) * (op.Hin / bh) * 4;
}

int pick_bn(int Cout) {
  if (Cout <= 16) return 16;
  if (Cout % 256 == 0) return 256;
  if (Cout % 128 == 0) return 128;
  if (Cout % 64 == 0) return 64;
  return 0;
}

}  // namespace tc

int conv_tc_init() {
  std::call_once(g_once, [] {
    g_init_status = do_init();
    if (std::getenv("HSIDM_NO_HALO")) tc::host().no_halo = 1;   // A/B switches for profiling runs
    if (std::getenv("HSIDM_NO_PAIRS")) tc::host().pairs_ok = 0;
    if (const char* v = std::getenv("HSIDM_VARIANT")) tc::host().variant = std::atoi(v);   // hsidm_debug_conv_mode's variant bits
  });
  return g_init_status;
}

bool conv_halo_ok(const ConvOp& op) { return !tc::host().no_halo && conv_halo_supported(op); }

int conv_tc_stats_slots(const ConvOp& op) {
  if (!conv_tc_supported(op, HSIDM_BF16)) return 0;
  if (!tc::host().no_halo && conv_halo_supported(op)) return conv_halo_stats_slots(op);
  return tc::pertap_stats_slots(op);
}

static int g_route_gen = 0;
void conv_tc_set_mode(int no_halo, int variant) {
  tc::host().no_halo = no_halo;
  tc::host().variant = variant;
  ++g_route_gen;   // cached CUDA graphs were captured under the previous routing
}
int conv_tc_route_gen() { return g_route_gen; }

int conv_tc_variant() { return tc::host().variant; }

int conv_tc_bn_rows(int Cout) { return pick_bn(Cout); }

bool conv_tc_supported(const ConvOp& op, int prec) {
  if (prec != HSIDM_BF16 || !op.w_bf16) return false;
  if (op.stride != 1 || op.up || (op.ksize != 1 && op.ksize != 3)) return false;
  if (op.src[0].layout != L_NHWC || op.src[0].C % kBK) return false;
  if (op.src[1].C && (op.src[1].layout != L_NHWC || op.src[1].C % kBK)) return false;
  if (op.up_parity >= 0) {
    if (op.up_parity > 3 || op.ksize != 3 || op.Hout != 2 * op.Hin || op.Wout != 2 * op.Win) return false;
  } else if (op.Hout != op.Hin || op.Wout != op.Win) {
    return false;
  }
  if (pick_bn(op.Cout) == 0) return false;
  if (op.out_layout == L_NHWC && op.Cout % 16) return false;
  if (op.out_layout != L_NHWC && op.resid) return false;
  int bw, bh, bn;
  return tile_geometry(op.Hin, op.Win, &bw, &bh, &bn);
}

int conv_tc(const ConvOp& op, cudaStream_t stream) {
  HSIDM_TRY(conv_tc_init());
  if (!conv_tc_supported(op, HSIDM_BF16))
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "conv_tc: op (Cin %d+%d, Cout %d, k%d s%d, %dx%d) does not fit the tensor-core kernel",
               op.src[0].C, op.src[1].C, op.Cout, op.ksize, op.stride, op.Hin, op.Win);
  if (!host().no_halo && conv_halo_supported(op)) return conv_halo(op, stream);
  if (op.rsrc[0].C || op.rsrc[1].C || op.up_parity >= 0 || op.gn.on() || op.s2)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "conv_tc: fused shortcut sources / sub-pixel upsampling / fused input GroupNorm need the halo kernel, which does not take this shape");
  TcP p;
  tile_geometry(op.Hin, op.Win, &p.bw, &p.bh, &p.bn);
  int BN = pick_bn(op.Cout);
  fill_epilogue(&p.e, op);
  p.M = op.N * op.Hin * op.Win;
  p.tiles_x = op.Win / p.bw;
  p.tiles_y = op.Hin / p.bh;
  p.m_tiles = p.bn > 1 ? (int)ceil_div(op.N, p.bn) : op.N * p.tiles_x * p.tiles_y;
  // small batches: narrower channel tiles while the grid still fits one wave - a 5-latent 8x8 conv is 3 pixel tiles, and
  // with BN = 256 six CTAs would walk K = 4608 while 142 SMs idle (the packed weight rows are padded to the widest BN)
  while (BN > 64 && p.m_tiles * (int)ceil_div(op.Cout, BN) * 2 <= host().num_sms) BN >>= 1;
  p.n_tiles = (int)ceil_div(op.Cout, BN);
  p.taps = op.ksize * op.ksize;
  p.bw_shift = 0, p.ppi_shift = 0;
  while ((1 << p.bw_shift) < p.bw) ++p.bw_shift;
  while ((1 << p.ppi_shift) < p.bw * p.bh) ++p.ppi_shift;
  p.chunks0 = op.src[0].C / kBK;
  p.chunks1 = op.src[1].C / kBK;
  p.err = host().err_flag;

  CUtensorMap tmA0, tmA1, tmB;
  HSIDM_TRY(encode_act_map(&tmA0, op.src[0].p, op.N, op.Hin, op.Win, op.src[0].C, p.bw, p.bh, p.bn));
  if (p.chunks1)
    HSIDM_TRY(encode_act_map(&tmA1, op.src[1].p, op.N, op.Hin, op.Win, op.src[1].C, p.bw, p.bh, p.bn));
  else
    tmA1 = tmA0;
  const int K = op.K();
  HSIDM_TRY(encode_weight_map(&tmB, op.w_bf16, K, p.n_tiles * BN, BN));
  CUtensorMap tmO = tmB;   // only the NHWC bf16 outputs with 64-multiple channels store through it
  if (op.out_layout == L_NHWC && op.Cout % 64 == 0)
    HSIDM_TRY(encode_rows_map(&tmO, op.out, (long long)op.N * op.Hout * op.Wout, op.Cout));

  const int grid = std::min(p.m_tiles * p.n_tiles, host().num_sms);
  char tag[96];
  snprintf(tag, sizeof(tag), "pertap BN%d k%d cin%d+%d cout%d %dx%d n%d", BN, op.ksize, op.src[0].C, op.src[1].C, op.Cout, op.Hin,
           op.Win, op.N);
  ProfScope prof(PROF_CONV_TC, 2.0 * p.M * (double)op.Cout * K, stream, tag);
  switch (BN) {
    case 16: HSIDM_CUDA(launch_pdl(conv_tc_kernel<16>, dim3(grid), dim3(kThreads), Cfg<16>::kSmemBytes, stream, 1, tmA0, tmA1, tmB, tmO, p)); break;
    case 64: HSIDM_CUDA(launch_pdl(conv_tc_kernel<64>, dim3(grid), dim3(kThreads), Cfg<64>::kSmemBytes, stream, 1, tmA0, tmA1, tmB, tmO, p)); break;
    case 128: HSIDM_CUDA(launch_pdl(conv_tc_kernel<128>, dim3(grid), dim3(kThreads), Cfg<128>::kSmemBytes, stream, 1, tmA0, tmA1, tmB, tmO, p)); break;
    default: HSIDM_CUDA(launch_pdl(conv_tc_kernel<256>, dim3(grid), dim3(kThreads), Cfg<256>::kSmemBytes, stream, 1, tmA0, tmA1, tmB, tmO, p)); break;
  }
  return after_launch("conv_tc_kernel");
}

// Reads and clears the barrier-timeout flag (0 = healthy). Synchronises the device; test/debug use only.
int conv_tc_error_flag(int* value) {
  *value = 0;
  int* flag = tc::host().err_flag;
  if (!flag) return HSIDM_OK;
  HSIDM_CUDA(cudaDeviceSynchronize());
  HSIDM_CUDA(cudaMemcpy(value, flag, sizeof(int), cudaMemcpyDeviceToHost));
  HSIDM_CUDA(cudaMemset(flag, 0, sizeof(int)));
  return HSIDM_OK;
}

}  // namespace hsidm
