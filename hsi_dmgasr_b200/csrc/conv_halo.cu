// tcgen05 3x3 convolution with the input halo resident in shared memory ("halo kernel").
//
// The per-tap kernel of conv_tc.cu re-fetches its A tile from L2 for each of the nine taps and streams one B tile
// per 128 output pixels; at 1.4 PFLOP/s that is ~95 B/clk/SM of L2->SMEM traffic against the ~45 B/clk/SM the L2 can
// deliver, which is what capped it at 20-47 % of tensor peak (profiles/r1_conv_pertap_ncu.md).  This kernel raises the
// arithmetic intensity per byte fetched:
//   * a CTA owns MT sub-tiles of 8 x 16 output pixels side by side (8*MT x 16 pixels) and BN output channels;
//   * per 64-channel chunk ONE TMA box {64 ch, 8*MT+2, 18, 1} brings the tile plus its 1-pixel halo into smem (zero
//     filled outside the image = conv padding); all 9 taps x MT sub-tiles read it in place: the A operand of tap
//     (dy,dx), sub-tile s is the same buffer addressed from row (1+dy)*PW + 8*s+1+dx with a 16-group stride of one
//     halo row (SBO = PW*128 B).  8 pixels of one image row form one 8-row swizzle group, so the canonical K-major
//     SWIZZLE_128B layout still applies; the hardware swizzle is a function of the shared-memory address bits, so a
//     start address that is a multiple of 128 B (not 1024 B) addresses the rows TMA wrote.
//   * one B tile (BN x 64 weights of one tap) feeds MT*4 MMAs instead of 4.
// L2->SMEM bytes per MMA cycle drop from ~95 to ~35-40 B/clk/SM.
// Accumulators: MT x BN fp32 columns per tile, double buffered (2*MT*BN <= 512 TMEM columns).
//
// Forms (template NT): 9 = 3x3 stride 1; 4 = the sub-pixel form of nearest-2x upsampling + 3x3 (a 2x2-tap conv over the
// low-res source per output parity); 1 = a short-K 1x1 conv (centre tap); 0 = 3x3 stride 2 over the four phase
// lattices of the input.  Optional per op: GroupNorm(+Swish) of the input applied in place to each halo tile
// (gn), a 1x1 shortcut over other tensors as extra centre-tap K columns (rsrc), per-slot GroupNorm partial sums
// of the output (stats).
//
// Warp roles (15 warps): 0 = TMA producer of the halo ring (2-3 stages), 1 = TMA producer of the weight ring (stages
// released per tap, full barriers shared by groups of 2-3 taps: every barrier wait of the issuer idles the tensor pipe
// for ~160 clk), 2 = MMA issuer (+ TMEM allocation), 3..10 = epilogue (two per TMEM lane quarter: tcgen05.ld -> bias
// -> act -> residual -> bf16 staging tile -> statistics -> one TMA store), 11..14 = in-place GroupNorm of halo tiles.
//
// PAIR = true runs the same pipeline on a cluster of two CTAs (cta_group::2): adjacent pixel tiles, one channel tile,
// the weight rows split between the CTAs, MMAs (M = 256) issued by the leader and completed on both CTAs' barriers.
#include <cmath>
#include <cstdlib>

#include "gn_fold.cuh"
#include "tc_common.cuh"

namespace hsidm {
namespace {

using namespace tc;

constexpr int kEpiWarps = 8;        // two warps per TMEM lane quarter
constexpr int kFirstEpiWarp = 3;    // warp 0: A (halo) producer, warp 1: B (weights) producer, warp 2: MMA issuer
// warps 11..: GroupNorm(+Swish) of the halo tile in place, when the op asks for it.  Four of them keep up with the MMAs of
// the wide tiles; the Cout <= 16 instantiation (the network's last conv) has almost no MMA work per tile, its tile time IS
// the transform, so it takes eight.
constexpr int xform_warps(int BN) { return BN == 16 ? 8 : 4; }
constexpr int kFirstXformWarp = kFirstEpiWarp + kEpiWarps;
constexpr int halo_threads(int BN) { return 32 * (kFirstXformWarp + xform_warps(BN)); }
constexpr int kRows = 16;          // output rows per tile
constexpr int kHaloRows = kRows + 2;

struct HaloP {
  int tiles_x, tiles_y;   // tiles per image
  int m_tiles, n_tiles;
  int chunks0, chunks1;   // 64-channel chunks of source 0 / 1
  int rchunks0, rchunks1; // 64-channel chunks of the shortcut sources (centre tap only; K columns after the 3x3 part)
  int dy0, dx0;           // halo offset of tap (0,0): 0,0 for 3x3 (NT = 9); (py,px) for the 2x2 sub-pixel form (NT = 4);
                          // 1,1 for a 1x1 conv (NT = 1, the centre tap)
  int kb0;                // first 64-wide k-block of this launch's weight columns (parity * 4 * chunks)
  int oscale, oy, ox;     // output pixel of source-tile pixel (i,j) = (oscale*i + oy, oscale*j + ox)
  int slot_base;          // first statistics slot of this launch
  int seg_max;            // statistics slots (x 2 for a pair) per (image, channel tile) group and launch: max CTA segments of a group
  int Hin, Win;           // source image size (the fused GroupNorm must leave the zero padding at zero)
  // fused input GroupNorm (kernels.cuh GnIn): the transform warps fold the statistics of the images this CTA works on
  int gn_on;
  const float* gn_part0; const float* gn_part1;   // [N][slots][C][2] partial sums of source 0 / 1
  int gn_slots0, gn_slots1;
  const float* gn_stats;  // alternative: precomputed (mean, rstd) [N][groups][2]
  const float* gn_gamma; const float* gn_beta;
  int gn_groups, gn_cpg;
  float gn_eps, gn_inv_cnt;
  int gn_swish;
  EpiP e;
  int* err;
  int tiles_q, tiles_r;   // tiles per CTA (pair): quotient and remainder of tiles / CTAs (the first tiles_r get one more)
  int variant;            // developer experiments (timing only, results invalid): 1 skip epilogue, 2 skip B loads, 4 skip A loads
  long long* dbg;         // developer timing probe (null in production): per-CTA wait cycles of every role
};

// Wait that also accumulates the cycles it took when the timing probe is on.
static __device__ __forceinline__ bool timed_wait(uint32_t bar, uint32_t parity, int* err, int code, const long long* dbg,
                                                  long long& acc) {
  if (!dbg) return mbar_wait(bar, parity, err, code);
  const long long t0 = clock64();
  const bool ok = mbar_wait(bar, parity, err, code);
  acc += clock64() - t0;
  return ok;
}

template <int MT, int BN, bool PAIR = false>
struct HCfg {
  static constexpr int kPW = 8 * MT + 2;                                   // halo row pitch in pixels
  // Halo stages: a tile's load (+ in-place GroupNorm) must hide behind the MMAs of the stages before it; with the short
  // shortcut chunks in the ring two stages leave it exposed, so the narrow tiles take three (the wide ones do not fit).
  static constexpr int kAStages = MT <= 2 ? 3 : 2;
  static constexpr int kABox = kHaloRows * kPW * 128;                      // bytes one TMA box writes
  static constexpr int kRBox = kRows * 8 * MT * 128;                        // ... for a shortcut chunk: the tile without halo
  static constexpr int kAStage = (kABox + 1023) / 1024 * 1024;             // keep every stage 1024-B aligned
  static constexpr int kBRows = PAIR ? BN / 2 : BN;                         // weight rows this CTA holds (a pair splits them)
  static constexpr int kBStage = kBRows * 128;
  static constexpr int kStageBytes = kEpiWarps * 4096;                     // epilogue staging, 4 KB per epilogue warp
  // Per-tile fold of the epilogue warps' GroupNorm partial sums: 512 B per warp and 64-channel chunk.  With at most one
  // chunk per warp (BN <= 128) the warps' bias slots double as the fold area; BN = 256 (two chunks per warp) has its own.
  static constexpr int kFoldBytes = BN / 64 > 2 ? kEpiWarps * 1024 : 0;
  static constexpr int kBStagesRaw = (232448 - 1536 - 2048 - kEpiWarps * 512 - kFoldBytes - kStageBytes - kAStages * kAStage) / kBStage;
  // Weight tiles land in groups of kBGroup taps that share ONE full barrier: every barrier wait of the MMA issuer
  // stalls the tensor pipe for ~160 clk (measured, scripts/mma_rate.cu), so it waits once per group, not per tap.
  // Stages are still released (tcgen05.commit, free) and refilled one tap at a time.
  static constexpr int kBGroup = kBStagesRaw >= 6 ? 3 : 2;
  static constexpr int kBGroups = (kBStagesRaw > 9 ? 9 : kBStagesRaw) / kBGroup;
  static constexpr int kBStages = kBGroups * kBGroup;
  static constexpr int kTmemCols = 2 * MT * BN < 32 ? 32 : 2 * MT * BN;
  static constexpr int kBiasBytes = kEpiWarps * 512;                       // 2 x 64 bias floats per epilogue warp
  static constexpr int kStatBytes = 2048;   // (mean, rstd) of a window of images x groups for the fused input GroupNorm
  static constexpr int kSmemBytes = kAStages * kAStage + kBStages * kBStage + kStageBytes + kBiasBytes + kFoldBytes + kStatBytes + 1024 + 512;
  static_assert(kBGroups >= 2, "not enough shared memory for the B ring");
  static_assert(2 * MT * BN <= 512, "accumulators do not fit TMEM");
};

template <int MT, int BN, int NT, bool PAIR>
__global__ void __launch_bounds__(halo_threads(BN), 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmR0, const __grid_constant__ CUtensorMap tmR1,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO, const HaloP p) {
  using C = HCfg<MT, BN, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + C::kAStages * C::kAStage;
  uint8_t* smem_stage = smem_b + C::kBStages * C::kBStage;
  uint8_t* smem_bias = smem_stage + C::kStageBytes;
  uint8_t* smem_fold = smem_bias + C::kBiasBytes;   // BN = 256 only; otherwise the bias slots are reused
  uint8_t* smem_stat = smem_fold + C::kFoldBytes;
  uint8_t* tail = smem_stat + C::kStatBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* a_empty = a_full + C::kAStages;
  uint64_t* a_ready = a_empty + C::kAStages;   // halo tile normalised in place (fused GroupNorm only)
  uint64_t* b_full = a_ready + C::kAStages;
  uint64_t* b_empty = b_full + C::kBStages;
  uint64_t* tfull_bar = b_empty + C::kBStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA pair (PAIR): the two CTAs of a cluster work on two adjacent pixel tiles (m-tiles 2q, 2q+1) of the SAME channel
  // tile; each holds its own halo tiles and half of the weight rows, and the leader (rank 0) issues M = 256
  // cta_group::2 MMAs over both.  Per MMA a CTA then reads its A slice plus HALF a B slice from shared memory.
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  // Tile order: every CTA (pair) walks a CONTIGUOUS range of tiles.  It then touches only a few consecutive images, whose
  // GroupNorm statistics its transform warps fold once (window in shared memory) - no finalize kernel in front.  (Groups
  // of CTAs striding through a shared range measured slower the larger the group: more images per CTA, more refills.)
  // The tile loops below run over an ARGUMENT-ONLY trip count (tiles_q + 1) and leave through a break: with a loop bound
  // derived from blockIdx ptxas takes every loop-carried ring index of the MMA issuer out of the uniform datapath and
  // the issuer pays five R2UR moves per tcgen05.mma (measured: 86 -> 122 clk per MMA on the 64-channel layers).
  const int cta = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile0 = cta * p.tiles_q + min(cta, p.tiles_r);
  const int total_tiles = tile0 + p.tiles_q + (cta < p.tiles_r ? 1 : 0);   // end of this CTA's range
  // tile -> (image n, channel tile nt, pixel tile r of the image): the tiles of one (image, channel tile) GROUP are
  // consecutive, pixel tile fastest - and the pixel tiles run DOWN the image first (r = tx * tiles_y + ty), so that a CTA's
  // next tile shares its two halo rows with the one it just loaded (an L2 hit; in row-major order the vertical neighbour
  // came tiles_x tiles later and its halo rows had been evicted: ncu showed DRAM reads at 1.2x the tensor size) (a pair's CTAs take the pixel tiles 2*rp and 2*rp + 1), so that an epilogue warp
  // can keep the group's GroupNorm partial sums in registers from tile to tile (see the epilogue)
  const int tpi = p.tiles_x * p.tiles_y;
  const int tpp = PAIR ? tpi >> 1 : tpi;   // tiles (pair tiles) per group
  auto decode = [&](int tile, int& n, int& nt, int& r) {
    const int u = tile / tpp, rp = tile - u * tpp;
    n = u / p.n_tiles, nt = u - n * p.n_tiles;
    r = PAIR ? 2 * rp + (int)rank : rp;
  };
  const int chunks = p.chunks0 + p.chunks1;
  const int rchunks = p.rchunks0 + p.rchunks1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.chunks1 || NT == 0) tma_prefetch_desc(&tmA1);
    if (p.rchunks0 || NT == 0) tma_prefetch_desc(&tmR0);
    if (p.rchunks1 || NT == 0) tma_prefetch_desc(&tmR1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < C::kAStages; ++s)
      mbar_init(smem_u32(&a_full[s]), 1), mbar_init(smem_u32(&a_empty[s]), 1),
          mbar_init(smem_u32(&a_ready[s]), PAIR ? 2 * xform_warps(BN) : xform_warps(BN));   // a pair's leader hears both CTAs
    for (int s = 0; s < C::kBStages; ++s) mbar_init(smem_u32(&b_empty[s]), 1);
    for (int g = 0; g < C::kBGroups; ++g) mbar_init(smem_u32(&b_full[g]), C::kBGroup);
    for (int s = 0; s < 2; ++s)
      mbar_init(smem_u32(&tfull_bar[s]), 1), mbar_init(smem_u32(&tempty_bar[s]), PAIR ? 2 * kEpiWarps : kEpiWarps);
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_pair(smem_u32(tmem_slot), C::kTmemCols);
    else tmem_alloc(smem_u32(tmem_slot), C::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  // REDUX leaves the (identical) value in a UNIFORM register: otherwise ptxas may treat the address read back from shared
  // memory as divergent and re-broadcast it (R2UR) in front of every tcgen05.mma - 10 clk per MMA on the issuer's path
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);

  if (warp == 0) {
    // =============================== TMA producer: halo tiles (A) ===============================
    if (lane == 0) {
      griddep_wait();     // activations come from the previous kernel(s) of the stream
      griddep_launch();   // ... and from here on the next kernel's CTAs may take the SMs this grid frees
      int as = 0;
      uint32_t aph = 0;
      bool ok = true;
      long long w_ae = 0;
      for (int kt = 0; kt <= p.tiles_q && ok; ++kt) {   // argument-only trip count, see tile0
        const int tile = tile0 + kt;
        if (tile >= total_tiles) break;
        int n, nt_unused, r;
        decode(tile, n, nt_unused, r);
        const int y0 = (r % p.tiles_y) * kRows, x0 = (r / p.tiles_y) * (8 * MT);
        for (int ch = 0; ch < chunks + rchunks; ++ch) {
          ok = timed_wait(smem_u32(&a_empty[as]), aph ^ 1, p.err, 1, p.dbg, w_ae);
          if (!ok) break;
          const uint32_t fb = smem_u32(&a_full[as]);
          const uint32_t dst = smem_u32(smem + as * C::kAStage);
          if (p.variant & 4) { mbar_arrive(fb); if (++as == C::kAStages) as = 0, aph ^= 1; continue; }
          mbar_expect_tx(fb, (NT != 0 && ch >= chunks) ? C::kRBox : C::kABox);
          if constexpr (NT == 0) {
            // stride-2 form: chunk = (phase, channel block); the four maps are the phase lattices (1,1) (1,0) (0,1) (0,0)
            const int cpp = chunks >> 2, ph = ch / cpp;
            const CUtensorMap* map = ph == 0 ? &tmA0 : ph == 1 ? &tmA1 : ph == 2 ? &tmR0 : &tmR1;
            tma_load_4d(dst, map, fb, (ch - ph * cpp) * kBK, x0 - 1, y0 - 1, n);
          } else if (ch < p.chunks0)
            tma_load_4d(dst, &tmA0, fb, ch * kBK, x0 - 1, y0 - 1, n);
          else if (ch < chunks)
            tma_load_4d(dst, &tmA1, fb, (ch - p.chunks0) * kBK, x0 - 1, y0 - 1, n);
          else if (ch < chunks + p.rchunks0)   // shortcut sources: centre tap only, so no halo (box {64, 8*MT, 16, 1})
            tma_load_4d(dst, &tmR0, fb, (ch - chunks) * kBK, x0, y0, n);
          else
            tma_load_4d(dst, &tmR1, fb, (ch - chunks - p.rchunks0) * kBK, x0, y0, n);
          if (++as == C::kAStages) as = 0, aph ^= 1;
        }
      }
      if (p.dbg) p.dbg[blockIdx.x * 8 + 0] = w_ae;
    }
  } else if (warp == 1) {
    // =============================== TMA producer: weight tiles (B) ===============================
    if (lane == 0) {
      int bs = 0, grp = 0, gcnt = 0;
      uint32_t bph = 0;
      bool ok = true;
      long long w_be = 0;
      // one weight tile: k-block kb of this tile's BN output channels, into the next stage of the ring
      auto load_b = [&](int kb, int nt) {
        ok = timed_wait(smem_u32(&b_empty[bs]), bph ^ 1, p.err, 5, p.dbg, w_be);
        if (!ok) return;
        const uint32_t bb = smem_u32(&b_full[grp]);
        if constexpr (PAIR) {
          // each CTA fetches its half of the weight rows; both halves complete on the LEADER's group barrier, which
          // the leader arms for the bytes of both
          if (rank == 0) mbar_expect_tx(bb, 2 * C::kBStage);
          tma_load_2d_pair(smem_u32(smem_b + bs * C::kBStage), &tmB, mapa_cluster(bb, 0), kb * kBK, nt * BN + (int)rank * C::kBRows);
        } else if (p.variant & 2) {
          mbar_arrive(bb);
        } else {
          mbar_expect_tx(bb, C::kBStage);
          tma_load_2d(smem_u32(smem_b + bs * C::kBStage), &tmB, bb, kb * kBK, nt * BN);
        }
        if (++bs == C::kBStages) bs = 0, bph ^= 1;
        if (++gcnt == C::kBGroup) gcnt = 0, grp = grp + 1 == C::kBGroups ? 0 : grp + 1;
      };
      for (int kt = 0; kt <= p.tiles_q && ok; ++kt) {   // argument-only trip count, see tile0
        const int tile = tile0 + kt;
        if (tile >= total_tiles) break;
        const int nt = (tile / tpp) % p.n_tiles;
        if constexpr (NT == 0) {   // stride-2 form: weights are packed in consumption order, 9 k-blocks per channel block
          for (int kb = 0; kb < 9 * (chunks >> 2) && ok; ++kb) load_b(kb, nt);
        }
        for (int ch = 0; ch < chunks && ok; ++ch)
          for (int tap = 0; tap < NT && ok; ++tap) load_b(p.kb0 + tap * chunks + ch, nt);
        for (int rc = 0; rc < rchunks && ok; ++rc) load_b(p.kb0 + NT * chunks + rc, nt);   // shortcut columns follow the 3x3 ones
      }
      // the issuer waits for whole groups: complete the last, partly filled one
      for (; ok && rank == 0 && gcnt != 0 && gcnt < C::kBGroup; ++gcnt) mbar_arrive(smem_u32(&b_full[grp]));
      if (p.dbg) p.dbg[blockIdx.x * 8 + 1] = w_be;
    }
  } else if (warp == 2) {
    // =============================== MMA issuer ===============================
    // The tensor pipe only stays busy while tcgen05.mma instructions arrive back to back: whatever else this thread
    // does between two of them (a barrier wait, address arithmetic) shows up as idle pipe time.  Hence: taps unrolled
    // with compile-time operand offsets, one weight-group wait per kBGroup taps, commits (free) per tap.
    if (lane == 0 && rank == 0) {
      constexpr int TW = NT == 9 ? 3 : 2;
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * kBM : kBM, BN);
      // one MMA / one completion signal, in the single-CTA or the pair form
      auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t acc_flag) {
        if constexpr (PAIR) umma_f16_pair(d, a, b, idesc, acc_flag);
        else umma_f16(d, a, b, idesc, acc_flag);
      };
      auto commit = [](uint32_t bar) {
        if constexpr (PAIR) umma_commit_pair(bar);
        else umma_commit(bar);
      };
      constexpr uint64_t desc_hi = (uint64_t)((C::kPW * 128) >> 4) << 32 | (1ull << 46) | (2ull << 61) | (1ull << 16);
      constexpr uint64_t bdesc_hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t tap0 = (uint32_t)((p.dy0 * C::kPW + p.dx0) * 128);   // halo offset of tap (0,0)
      const uint32_t b_lo = smem_u32(smem_b) >> 4;
      int as = 0, bs = 0, acc = 0, grp = 0, gcnt = 0;
      uint32_t aph = 0, gph = 0, acc_phase = 0;
      bool ok = true;
      long long w_te = 0, w_af = 0, w_bf = 0;
      const long long t_start = p.dbg ? clock64() : 0;
      for (int kt = 0; kt <= p.tiles_q && ok; ++kt) {   // argument-only trip count, see tile0
        const int tile = tile0 + kt;
        if (tile >= total_tiles) break;
        ok = timed_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1, p.err, 2, p.dbg, w_te);
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (MT * BN);
        for (int ch = 0; ch < chunks && ok; ++ch) {
          ok = timed_wait(smem_u32((PAIR || p.gn_on) ? &a_ready[as] : &a_full[as]), aph, p.err, 3, p.dbg, w_af);
          if (!ok) break;
          const uint64_t adesc0 = desc_hi | (uint64_t)((smem_u32(smem + as * C::kAStage) + tap0) >> 4);
          // one tap: (group barrier) -> MT*4 MMAs from the halo tile at byte offset `off` -> release the weight stage
          auto tap_body = [&](int off, bool first) {
            if (gcnt == 0) {
              ok = timed_wait(smem_u32(&b_full[grp]), gph, p.err, 6, p.dbg, w_bf);
              if (!ok) return;
              tc_fence_after();
            }
            const uint64_t bdesc = bdesc_hi | (uint64_t)(b_lo + bs * (C::kBStage >> 4));
            // sub-tile outer / k-step inner.  The other order (consecutive MMAs to different accumulators) measures the
            // same in the micro-benchmark and in the kernel (17.75 vs 17.75 ms per step, profiles/r2_experiments.md).
#pragma unroll
            for (int s = 0; s < MT; ++s) {
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)
                mma(d_tmem + s * BN, adesc0 + (uint64_t)((off + s * 1024) / 16 + 2 * k), bdesc + 2 * k, (first && k == 0) ? 0u : 1u);
            }
            commit(smem_u32(&b_empty[bs]));
            if (++bs == C::kBStages) bs = 0;
            if (++gcnt == C::kBGroup) {
              gcnt = 0;
              if (++grp == C::kBGroups) grp = 0, gph ^= 1;
            }
          };
          if constexpr (NT == 0) {
            // stride-2 3x3 conv over the four phase lattices of the input: phase (py,px) serves the taps whose source
            // row/column has that parity, at lattice offsets -1 / 0 (halo coordinates 0 / 1)
            constexpr int R = C::kPW * 128, X = 128;
            const int ph = ch / (chunks >> 2);
            if (ph == 0) {          // (1,1): taps (ky,kx) in {0,2}x{0,2}
              tap_body(0, ch == 0);
              if (ok) tap_body(X, false);
              if (ok) tap_body(R, false);
              if (ok) tap_body(R + X, false);
            } else if (ph == 1) {   // (1,0): ky in {0,2}, kx = 1
              tap_body(X, false);
              if (ok) tap_body(R + X, false);
            } else if (ph == 2) {   // (0,1): ky = 1, kx in {0,2}
              tap_body(R, false);
              if (ok) tap_body(R + X, false);
            } else {                // (0,0): the centre tap
              tap_body(R + X, false);
            }
            if (!ok) break;
          }
#pragma unroll
          for (int tap = 0; tap < NT; ++tap) {
            tap_body(((tap / TW) * C::kPW + tap % TW) * 128, tap == 0 && ch == 0);   // halo offset: constant once unrolled
            if (!ok) break;
          }
          commit(smem_u32(&a_empty[as]));
          if (++as == C::kAStages) as = 0, aph ^= 1;
        }
        for (int rc = 0; rc < rchunks && ok; ++rc) {   // shortcut: centre tap of the un-normalised block input
          ok = timed_wait(smem_u32((PAIR || p.gn_on) ? &a_ready[as] : &a_full[as]), aph, p.err, 3, p.dbg, w_af);
          if (!ok) break;
          if (gcnt == 0) {
            ok = timed_wait(smem_u32(&b_full[grp]), gph, p.err, 6, p.dbg, w_bf);
            if (!ok) break;
          }
          tc_fence_after();
          // the shortcut tile has no halo: rows of 8*MT pixels, so 8-pixel groups of consecutive image rows are 8*MT*128 B apart
          constexpr uint64_t rdesc_hi = (uint64_t)((8 * MT * 128) >> 4) << 32 | (1ull << 46) | (2ull << 61) | (1ull << 16);
          const uint64_t adesc0 = rdesc_hi | (uint64_t)(smem_u32(smem + as * C::kAStage) >> 4);
          const uint64_t bdesc = bdesc_hi | (uint64_t)(b_lo + bs * (C::kBStage >> 4));
#pragma unroll
          for (int s = 0; s < MT; ++s) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              mma(d_tmem + s * BN, adesc0 + (uint64_t)(s * 1024 / 16 + 2 * k), bdesc + 2 * k, 1u);
          }
          commit(smem_u32(&b_empty[bs]));
          if (++bs == C::kBStages) bs = 0;
          if (++gcnt == C::kBGroup) {
            gcnt = 0;
            if (++grp == C::kBGroups) grp = 0, gph ^= 1;
          }
          commit(smem_u32(&a_empty[as]));
          if (++as == C::kAStages) as = 0, aph ^= 1;
        }
        commit(smem_u32(&tfull_bar[acc]));
        if (++acc == 2) acc = 0, acc_phase ^= 1;
      }
      if (p.dbg) {
        p.dbg[blockIdx.x * 8 + 2] = w_te, p.dbg[blockIdx.x * 8 + 3] = w_af, p.dbg[blockIdx.x * 8 + 4] = w_bf;
        p.dbg[blockIdx.x * 8 + 5] = clock64() - t_start;
      }
    }
  } else if (warp >= kFirstXformWarp) {
    // =============================== fused GroupNorm(+Swish) of the halo tiles ===============================
    // y = act(x*A[n,c] + B[n,c]) in place on every in-image pixel of the tile TMA just delivered; pixels outside the
    // image keep the zeros TMA filled in (the conv pads the NORMALISED tensor).  Thread -> one 16-byte chunk (8
    // channels, so 16 affine coefficients in registers) of rows tid/8, tid/8 + 16, ...; a row's logical chunk j sits
    // at physical chunk j ^ (row & 7) (128B swizzle).  Same arithmetic as gn_apply (norm.cu), so fused and unfused
    // paths agree bit for bit.
    if (PAIR || p.gn_on) {   // a pair always routes halo readiness through these warps (the leader hears both CTAs)
      griddep_wait();          // the statistics are the previous kernels' output
      const int tid = threadIdx.x - 32 * kFirstXformWarp;
      // a_ready of the CTA whose issuer consumes the stage: this CTA's own, or the pair leader's
      auto ready_arrive = [&](int stage) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_cluster(smem_u32(&a_ready[stage]), 0));
        else mbar_arrive(smem_u32(&a_ready[stage]));
      };
      const int j8 = tid & 7, row0 = tid >> 3;
      constexpr int kRowsTot = kHaloRows * C::kPW;
      constexpr int kXformWarps = xform_warps(BN);
      constexpr int kRowStep = 32 * kXformWarps / 8;
      const int C0 = p.chunks0 * kBK, C1 = p.chunks1 * kBK;
      // Statistics window: (mean, rstd) of images [win_lo, win_lo + win_n) x groups in shared memory.  This CTA's tiles
      // are consecutive, so the window moves forward a few times per launch at most.  Thread -> one (image, group), see
      // gn_fold_group.
      float2* win = reinterpret_cast<float2*>(smem_stat);
      const GnFoldP fold{p.gn_part0, p.gn_part1, p.gn_slots0, p.gn_slots1, C0, C1, p.gn_cpg, p.gn_inv_cnt, p.gn_eps};
      const int win_n = max(1, min(C::kStatBytes / 8 / max(1, p.gn_groups), 64));
      int win_lo = -(1 << 30);
      const int n_end = total_tiles > tile0 ? (total_tiles - 1) / tpp / p.n_tiles : 0;   // last image this CTA touches
      auto fill_window = [&](int n0) {
        named_bar_sync(2, 32 * kXformWarps);   // nobody still reads the old window
        // images [n0, n0 + cnt) x groups; with few of them, 2 or 4 adjacent lanes share one (image, group)
        const int cnt = min(win_n, min(p.e.N_img, n_end + 1) - n0), items = cnt * p.gn_groups;
        const int S = p.gn_stats ? 1 : items <= 32 ? 4 : items <= 64 ? 2 : 1, per_pass = 32 * kXformWarps / S;
        for (int base = 0; base < items; base += per_pass) {
          const int i = base + tid / S;
          const bool valid = i < items;
          const int wi = i / p.gn_groups, g = i - wi * p.gn_groups, n = n0 + wi;
          float2 v;
          if (p.gn_stats) v = valid ? *reinterpret_cast<const float2*>(p.gn_stats + ((long long)n * p.gn_groups + g) * 2) : make_float2(0.f, 0.f);
          else v = gn_fold_group(fold, n, g, tid % S, S, valid);
          if (valid && tid % S == 0) win[i] = v;
        }
        named_bar_sync(2, 32 * kXformWarps);
        win_lo = n0;
      };
      int as = 0;
      uint32_t aph = 0;
      bool ok = true;
      for (int kt = 0; kt <= p.tiles_q && ok; ++kt) {   // argument-only trip count, see tile0
        const int tile = tile0 + kt;
        if (tile >= total_tiles) break;
        int n, nt_unused, r;
        decode(tile, n, nt_unused, r);
        const int y0 = (r % p.tiles_y) * kRows - 1, x0 = (r / p.tiles_y) * (8 * MT) - 1;   // image coordinates of halo (0,0)
        if (p.gn_on && n >= win_lo + win_n) fill_window(n);
        const float2* wst = win + (p.gn_on ? (n - win_lo) * p.gn_groups : 0);
        for (int ch = 0; ch < chunks && ok; ++ch) {
          // coefficients of this thread's 8 channels (before the wait: independent of the tile data): A = rstd*gamma,
          // B = beta - mean*A, the expressions gn_apply evaluates
          float4 ab[4];
          if (p.gn_on) {
            const int c0 = ch * kBK + j8 * 8;
            const float4* gp = reinterpret_cast<const float4*>(p.gn_gamma + c0);
            const float4* bp = reinterpret_cast<const float4*>(p.gn_beta + c0);
            const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), b0 = __ldg(bp), b1 = __ldg(bp + 1);
            const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            int g = c0 / p.gn_cpg, rem = c0 - g * p.gn_cpg;
            float2 st = wst[g];
            float A[8], B[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              A[q] = st.y * gam[q], B[q] = bet[q] - st.x * A[q];
              if (++rem == p.gn_cpg && q < 7) rem = 0, st = wst[++g];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) ab[q] = make_float4(A[2 * q], B[2 * q], A[2 * q + 1], B[2 * q + 1]);
          }
          ok = mbar_wait(smem_u32(&a_full[as]), aph, p.err, 7);
          if (!ok) break;
          if (!p.gn_on) {   // pair without a fused GroupNorm: relay only
            __syncwarp();
            if (lane == 0) ready_arrive(as);
            if (++as == C::kAStages) as = 0, aph ^= 1;
            continue;
          }
          const uint32_t base = smem_u32(smem + as * C::kAStage);
          int hy = row0 / C::kPW, hx = row0 % C::kPW;   // halo coordinates of this thread's current row
#pragma unroll 2
          for (int row = row0; row < kRowsTot; row += kRowStep) {
            if ((unsigned)(y0 + hy) < (unsigned)p.Hin && (unsigned)(x0 + hx) < (unsigned)p.Win) {
              const uint32_t addr = base + (uint32_t)(row * 128 + ((j8 ^ (row & 7)) << 4));
              uint4 v = lds128(addr);
              uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 f = bf16x2_to_f2(w[q]);
                float y0f = fmaf(f.x, ab[q].x, ab[q].y), y1f = fmaf(f.y, ab[q].z, ab[q].w);
                if (p.gn_swish) {
                  float t0, t1;
                  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(0.5f * y0f));
                  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(0.5f * y1f));
                  y0f = y0f * fmaf(0.5f, t0, 0.5f), y1f = y1f * fmaf(0.5f, t1, 0.5f);
                }
                const __nv_bfloat162 h = __floats2bfloat162_rn(y0f, y1f);
                w[q] = *reinterpret_cast<const uint32_t*>(&h);
              }
              sts128(addr, v);
            }
            hx += kRowStep % C::kPW, hy += kRowStep / C::kPW;
            if (hx >= C::kPW) hx -= C::kPW, ++hy;
          }
          fence_proxy_async_smem();   // the MMA reads these rows through the async proxy
          __syncwarp();
          if (lane == 0) ready_arrive(as);
          if (++as == C::kAStages) as = 0, aph ^= 1;
        }
        // Shortcut tiles pass through un-normalised, but every use of a stage is still acknowledged through a_ready:
        // the issuer then cannot hand a stage back to the producer before all four warps have seen its phase, so no
        // waiter can fall a ring lap behind (where the parity of an old phase would alias the one it waits for).
        for (int rc = 0; rc < rchunks && ok; ++rc) {
          ok = mbar_wait(smem_u32(&a_full[as]), aph, p.err, 7);
          if (!ok) break;
          __syncwarp();
          if (lane == 0) ready_arrive(as);
          if (++as == C::kAStages) as = 0, aph ^= 1;
        }
      }
    }
  } else {
    // =============================== epilogue (warps 3..10) ===============================
    griddep_wait();   // residual / noise-bias reads and the output writes must follow the previous kernels of the stream
    const int quarter = warp & 3;
    const int half = (warp - kFirstEpiWarp) >> 2;      // which of the two warps of this lane quarter
    const int row = quarter * 32 + lane;   // accumulator row: pixel (row/8, row%8) of a sub-tile
    constexpr int nC = BN / 64;            // 64-channel chunks per tile (0 for the small-Cout instantiation)
    int acc = 0;
    uint32_t acc_phase = 0;
    const float* nbias = p.e.nbias ? p.e.nbias + (p.e.nb_t ? (long long)(*p.e.nb_t) * p.e.nb_ts : 0) : nullptr;
    bool ok = true;
    long long w_tf = 0;
    const long long t_start = p.dbg ? clock64() : 0;
    // GroupNorm partial sums of the output: every warp keeps the (sum, sumsq) of its rows x its 64-channel chunk(s) in
    // registers across the consecutive tiles of one (image, channel tile) group; when the group changes the eight warps
    // fold them through shared memory (two named barriers - per GROUP, not per tile: a barrier per tile puts the warps
    // in lock step and costs the epilogue-bound layers 20-30 %) and publish ONE slot per (group segment, CTA).
    // A group cut by the CTA ranges has one slot per segment; the CTA holding its last tile zero-fills the rest.
    constexpr int kFoldK = C::kFoldBytes ? 2 : 1;   // chunks per warp the fold area holds
    const uint32_t fold_base = smem_u32(C::kFoldBytes ? smem_fold : smem_bias);
    const bool stats_on = nC >= 1 && p.e.stats && !(p.variant & 1);
    float4 gsum[kFoldK];
    int u_cur = -1;
    auto flush_group = [&]() {
      if constexpr (nC >= 1) {
        constexpr bool by_chunk = nC >= 2;
        const int we = warp - kFirstEpiWarp;
#pragma unroll
        for (int k = 0; k < kFoldK; ++k)
          sts128(fold_base + (uint32_t)(((we * kFoldK + k) * 32 + lane) * 16),
                 make_uint4(__float_as_uint(gsum[k].x), __float_as_uint(gsum[k].y), __float_as_uint(gsum[k].z), __float_as_uint(gsum[k].w)));
        named_bar_sync(1, 32 * kEpiWarps);
        if (we < nC) {
          // epilogue warp `we` adds, in warp order, the partial sums of the warps that worked on chunk `we`
          const int w0 = by_chunk ? 4 * (we & 1) : 0, w1 = by_chunk ? w0 + 4 : kEpiWarps, kk = by_chunk ? we >> 1 : 0;
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int w = w0; w < w1; ++w) {
            const uint4 u = lds128(fold_base + (uint32_t)(((w * kFoldK + kk) * 32 + lane) * 16));
            t.x += __uint_as_float(u.x), t.y += __uint_as_float(u.y), t.z += __uint_as_float(u.z), t.w += __uint_as_float(u.w);
          }
          // segment index of this CTA within the group = CTAs between the owner of the group's first tile and this one
          const int first = u_cur * tpp, cut = p.tiles_r * (p.tiles_q + 1);
          const int owner = first < cut ? first / (p.tiles_q + 1) : p.tiles_r + (first - cut) / p.tiles_q;
          const int n_g = u_cur / p.n_tiles, nt_g = u_cur - n_g * p.n_tiles;
          constexpr int R = PAIR ? 2 : 1;
          int slot = p.slot_base + (cta - owner) * R + (int)rank;
          stats_store(p.e, n_g, slot, nt_g * BN + we * 64, lane, t);
          if (total_tiles >= first + tpp)   // this CTA held the group's last tile: the segments nobody wrote count as zero
            for (slot += R; slot < p.slot_base + p.seg_max * R; slot += R)
              stats_store(p.e, n_g, slot, nt_g * BN + we * 64, lane, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        named_bar_sync(1, 32 * kEpiWarps);   // the fold area (= the bias slots) may be overwritten from here on
      }
    };
    for (int kt = 0; kt <= p.tiles_q && ok; ++kt) {   // argument-only trip count, see tile0
      const int tile = tile0 + kt;
      if (tile >= total_tiles) break;
      int n, nt, r;
      decode(tile, n, nt, r);
      if (stats_on && tile / tpp != u_cur) {
        if (u_cur >= 0) flush_group();
        u_cur = tile / tpp;
#pragma unroll
        for (int k = 0; k < kFoldK; ++k) gsum[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const int y = (r % p.tiles_y) * kRows + (row >> 3);
      const int xb = (r / p.tiles_y) * (8 * MT) + (row & 7);
      // this warp's share of the tile: with several 64-channel chunks (BN >= 128) the two warps of a lane quarter take
      // alternate chunks ci = half, half + 2, ... of every sub-tile; with one chunk they take alternate sub-tiles
      constexpr bool by_chunk = nC >= 2;
      const int c_first = by_chunk ? half : 0, c_step = by_chunk ? 2 : 1;
      const int s_first = by_chunk ? 0 : half, s_step = by_chunk ? 1 : 2;
      const int ty0 = (r % p.tiles_y) * kRows, tx0 = (r / p.tiles_y) * (8 * MT);
      const uint32_t bias_base = smem_u32(smem_bias + (warp - kFirstEpiWarp) * 512);
      bool have_bias = false;
      if constexpr (nC >= 1) {
        // park the bias values of this warp's chunks (conv bias and/or this image's noise embedding) in shared memory
        // while the accumulators are still being produced
        const float* cb = nbias ? nbias + (long long)n * p.e.nbs : p.e.bias;
        if (cb && !(p.variant & 1)) {
          have_bias = true;
          const int k = lane >> 4;   // lanes 0-15 park the first chunk, 16-31 the second (if the warp has one)
          const int ci = c_first + k * c_step;
          if (ci < nC) {
            float4 b = __ldg(reinterpret_cast<const float4*>(cb + nt * BN + ci * 64) + (lane & 15));
            if (nbias && p.e.bias) {
              const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.e.bias + nt * BN + ci * 64) + (lane & 15));
              b.x += b2.x, b.y += b2.y, b.z += b2.z, b.w += b2.w;
            }
            sts128(bias_base + k * 256 + (lane & 15) * 16,
                   make_uint4(__float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w)));
          }
          __syncwarp();
        }
      }
      ok = timed_wait(smem_u32(&tfull_bar[acc]), acc_phase, p.err, 4, p.dbg, w_tf);
      if (!ok) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * (MT * BN) + ((uint32_t)(quarter * 32) << 16);
      if (p.variant & 1) {
      } else if constexpr (nC >= 1) {
        static_assert(nC <= 4, "an epilogue warp parks the bias of at most two chunks");
        const uint32_t stage = smem_u32(smem_stage + (warp - kFirstEpiWarp) * 4096);
        const long long xstep = (long long)p.oscale * p.e.Cout, pitch = (long long)p.oscale * p.e.W * p.e.Cout;
#pragma unroll 1
        for (int ci = c_first, k = 0; ci < nC; ci += c_step, ++k) {
          const int co0 = nt * BN + ci * 64;
          // residual layers: this lane's element of the tile in the residual tensor (same lattice as the output)
          const bf16* resid_lane = nullptr;
          if (p.e.resid) {
            const long long m_q = ((long long)n * p.e.H + p.oscale * (ty0 + quarter * 4) + p.oy) * p.e.W + p.oscale * tx0 + p.ox;
            resid_lane = p.e.resid + m_q * p.e.Cout + (lane >> 3) * xstep + co0 + (lane & 7) * 8;
          }
          float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
          for (int s = s_first; s < MT; s += s_step)
            epilogue_tma64(p.e, &tmO, have_bias ? bias_base + k * 256 : 0u, taddr + s * BN + ci * 64, lane, stage, co0, tx0 + 8 * s,
                           ty0 + quarter * 4, n, resid_lane ? resid_lane + 8 * s * xstep : nullptr, pitch, 4 * xstep, st);
          if (stats_on) {   // this tile's sums of chunk ci (the warp's k-th) join the group's
            if (kFoldK == 2 && (k & 1)) gsum[kFoldK - 1].x += st.x, gsum[kFoldK - 1].y += st.y, gsum[kFoldK - 1].z += st.z, gsum[kFoldK - 1].w += st.w;
            else gsum[0].x += st.x, gsum[0].y += st.y, gsum[0].z += st.z, gsum[0].w += st.w;
          }
        }
      } else {
#pragma unroll 1
        for (int s = half; s < MT; s += 2) {
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + s * BN + c0, v);
            tmem_ld_wait();
            const int co0 = nt * BN + c0;
            if (co0 < p.e.Cout) epilogue16(p.e, nbias, n, y, xb + 8 * s, co0, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // the issuer (of the pair's leader) may overwrite this accumulator buffer
        if constexpr (PAIR) mbar_arrive_cluster(mapa_cluster(smem_u32(&tempty_bar[acc]), 0));
        else mbar_arrive(smem_u32(&tempty_bar[acc]));
      }
      if (++acc == 2) acc = 0, acc_phase ^= 1;
    }
    if (stats_on && u_cur >= 0) flush_group();
    if (lane == 0) tma_store_wait_all();   // the staging tiles must outlive the stores reading them
    if (p.dbg && warp == kFirstEpiWarp && lane == 0) p.dbg[blockIdx.x * 8 + 6] = w_tf, p.dbg[blockIdx.x * 8 + 7] = clock64() - t_start;
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // neither CTA may leave while the other can still signal it or read its smem
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// (MT, BN) for an op, or MT = 0 when the halo kernel does not apply; *pair = run it on CTA pairs (cta_group::2).
void pick_shape(const ConvOp& op, int* MT, int* BN, bool* pair) {
  *MT = 0, *BN = 0, *pair = false;
  if (op.stride != 1 || op.up) return;
  // 1x1 convs with a short K (one or two 64-channel chunks) are per-tile-overhead bound in the per-tap kernel; here
  // they run as the centre tap alone over 4x larger tiles.  Longer-K 1x1 convs stay with the per-tap kernel, whose
  // halo-free A tiles and BN = 256 serve them better.
  if (op.ksize == 1 && (op.src[0].C + op.src[1].C > 2 * kBK || op.rsrc[0].C || op.up_parity >= 0)) return;
  if (op.ksize != 3 && op.ksize != 1) return;
  if (op.Hin % kRows) return;
  // CTA pairs need an even number of pixel tiles (the two CTAs take tiles 2q and 2q+1).  Only the N = 256 pair tile is
  // on by default: its MMAs run at the tensor pipe's full rate (128 clk) where the single-CTA N = 128 tile is capped at
  // ~90 clk per 128x128x16 by shared-memory operand bandwidth (deep layers 1.13 -> 1.35 PFLOP/s).  The N = 128 and
  // N = 64 pair tiles gain 12 % / 28 % in the MMA micro-benchmark but measured 15-25 % SLOWER end to end (the pair
  // advances at the pace of its slower CTA on every chunk); variant bits 64 / 128 switch them on for A/B runs, bit 32
  // switches the N = 256 pair off.
  const int var = host().variant ^ (64 | 128);
  // the two CTAs of a pair take the pixel tiles 2*rp and 2*rp + 1 of the SAME image
  auto even_tiles = [&](int mt) { return (((op.Hin / kRows) * (op.Win / (8 * mt))) & 1) == 0; };
  const bool pair_ok = op.ksize == 3 && !op.s2 && host().pairs_ok;
  if (op.s2 && (op.ksize != 3 || op.src[1].C || op.rsrc[0].C || op.up_parity >= 0 || op.gn.on())) return;
  // Candidates in the order of preference at full batch:
  //   <1,256> on CTA pairs - the only shape whose MMAs run at the tensor pipe's full rate;
  //   <2,128>, <4,64> - widest channel tile the layer has, most sub-tiles per weight tile;
  //   <2,64> - widths that are only a multiple of 16 (also the test knob, variant bit 16: measured 10-30 % slower than
  //   <4,64> where both apply: half the weight-tile reuse, twice the per-tile overhead);
  //   <1,128>, <1,64> - small batches only, see below.
  struct Cand {
    int mt, bn;
    bool pair;
  } cands[6];
  int nc = 0;
  const bool k3 = op.ksize == 3 && !op.s2;
  if (op.Cout % 256 == 0 && op.Win % 8 == 0 && pair_ok && !(var & 32) && even_tiles(1)) cands[nc++] = {1, 256, true};
  if (op.Cout % 128 == 0 && op.Win % 16 == 0) cands[nc++] = {2, 128, pair_ok && !(var & 64) && even_tiles(2)};
  if (op.Cout % 64 == 0 && op.Win % 32 == 0 && !(var & 16)) cands[nc++] = {4, 64, pair_ok && !(var & 128) && even_tiles(4)};
  if (op.Cout % 64 == 0 && op.Win % 16 == 0) cands[nc++] = {2, 64, false};
  if (k3 && op.Cout % 128 == 0 && op.Win % 8 == 0) cands[nc++] = {1, 128, false};
  if (k3 && op.Cout % 64 == 0 && op.Win % 8 == 0) cands[nc++] = {1, 64, false};
  if (nc == 0) {
    // Cout <= 16 (the network's last conv).  The narrower <2,16> tile with three halo stages was measured slower
    // (0.44 against 0.35 ms at 176 latents, profiles/r2_experiments.md).
    if (op.Cout <= 16 && op.Win % 32 == 0 && op.ksize == 3 && !op.s2) *MT = 4, *BN = 16;
    return;
  }
  // Pick by a small cost model: rounds of CTAs x cost of one tile.  A tile costs its MMA work over the rate its N
  // sustains from shared memory (N = 64: 58 clk against a 32-clk floor, N = 128: 68 / 64, N = 256: full rate - section 4
  // of DESIGN.md) plus a fixed per-tile overhead (pipeline fill, epilogue tail) that favours the larger tiles.  At 176
  // latents every layer has tens of tiles per SM and the first candidate stands (the model is only consulted when that
  // shape yields less than one round of CTAs); with a small batch the launch
  // becomes as wide as the layer allows (5 latents at 16x16 x 512 channels are 20 <1,256> tiles each walking K = 4608
  // alone - or 80 <1,64> tiles).
  auto cost = [&](const Cand& c) {
    const double ctas = (double)op.N * (op.Hin / kRows) * (op.Win / (8 * c.mt)) * (op.Cout / c.bn), sms = host().num_sms;
    const double rounds = ctas <= 4 * sms ? std::ceil(ctas / sms) : ctas / sms;
    const double rate = c.bn >= 256 ? 1.0 : c.bn == 128 ? 0.94 : 0.55;
    return rounds * (c.mt * c.bn / rate + 40.0);
  };
  int best = 0;
  const bool one_round = (double)op.N * (op.Hin / kRows) * (op.Win / (8 * cands[0].mt)) * (op.Cout / cands[0].bn) < host().num_sms;
  if (one_round && !(var & 1024))   // the model only decides where the full-batch shape leaves SMs idle
    for (int i = 1; i < nc; ++i)
      if (cost(cands[i]) < 0.97 * cost(cands[best])) best = i;   // a later (smaller) tile must win clearly
  *MT = cands[best].mt, *BN = cands[best].bn, *pair = cands[best].pair;
}

// Grid and tile partition of a launch: CTAs (pairs), tiles per CTA (quotient / remainder) and the largest number of CTA
// ranges that can cut one (image, channel tile) group of `tpp` consecutive tiles.
struct Partition {
  int ctas, q, r, seg_max;
};
Partition partition(const ConvOp& op, int MT, int BN, bool pair) {
  const int tpi = (op.Win / (8 * MT)) * (op.Hin / kRows), n_tiles = (int)ceil_div(op.Cout, BN);
  const int tpp = pair ? tpi / 2 : tpi;
  const long long tiles = (long long)op.N * n_tiles * tpp;
  Partition pt;
  pt.ctas = (int)std::min<long long>(tiles, pair ? host().num_sms / 2 : host().num_sms);
  pt.q = (int)(tiles / pt.ctas), pt.r = (int)(tiles % pt.ctas);
  pt.seg_max = (tpp - 1) / pt.q + 2;
  return pt;
}

template <int MT, int BN, int NT, bool PAIR>
int launch(const ConvOp& op, cudaStream_t stream) {
  using C = HCfg<MT, BN, PAIR>;
  HaloP p;
  p.tiles_x = op.Win / (8 * MT);
  p.tiles_y = op.Hin / kRows;
  p.m_tiles = op.N * p.tiles_x * p.tiles_y;
  p.n_tiles = (int)ceil_div(op.Cout, BN);
  p.chunks0 = op.src[0].C / kBK;
  p.chunks1 = op.src[1].C / kBK;
  p.rchunks0 = op.rsrc[0].C / kBK;
  p.rchunks1 = op.rsrc[1].C / kBK;
  fill_epilogue(&p.e, op);
  p.err = host().err_flag;
  p.dbg = host().halo_dbg;
  p.Hin = op.Hin, p.Win = op.Win;
  const GnIn& gn = op.gn;
  p.gn_on = gn.on() ? 1 : 0;
  p.gn_part0 = gn.part[0], p.gn_part1 = gn.part[1], p.gn_slots0 = gn.slots[0], p.gn_slots1 = gn.slots[1];
  p.gn_stats = gn.stats, p.gn_gamma = gn.gamma, p.gn_beta = gn.beta, p.gn_groups = gn.groups, p.gn_swish = gn.swish, p.gn_eps = gn.eps;
  p.gn_cpg = 1, p.gn_inv_cnt = 0.f;
  if (p.gn_on) {
    const int Ct = op.src[0].C + op.src[1].C;
    if (gn.groups <= 0 || Ct % gn.groups || (!gn.stats && ((Ct / gn.groups) & 1)) || 8 * gn.groups > C::kStatBytes ||
        (!gn.stats && (gn.C[0] != op.src[0].C || gn.C[1] != op.src[1].C || (op.src[1].C && !gn.part[1]))))
      HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "conv_halo: fused GroupNorm over %d+%d channels in %d groups is not supported", op.src[0].C, op.src[1].C, gn.groups);
    p.gn_cpg = Ct / gn.groups;
    p.gn_inv_cnt = (float)(1.0 / ((double)p.gn_cpg * op.Hin * op.Win));
  }
  p.variant = host().variant & 7;
  p.dy0 = p.dx0 = NT == 1 ? 1 : 0, p.kb0 = 0, p.oscale = 1, p.oy = p.ox = 0, p.slot_base = 0;
  const Partition pt = partition(op, MT, BN, PAIR);
  p.tiles_q = pt.q, p.tiles_r = pt.r, p.seg_max = pt.seg_max;
  if (op.up_parity >= 0) {
    const int py = op.up_parity >> 1, px = op.up_parity & 1;
    p.dy0 = py, p.dx0 = px, p.kb0 = op.up_parity * 4 * (p.chunks0 + p.chunks1);
    p.oscale = 2, p.oy = py, p.ox = px;
    p.slot_base = op.up_parity * (pt.seg_max * (PAIR ? 2 : 1));
  }
  CUtensorMap tmA0, tmA1, tmB;
  if (NT == 0) {
    // stride-2 form: every 64-channel block is visited once per phase lattice; the four maps are the phases
    // (1,1) (1,0) (0,1) (0,0) of the full-resolution source
    p.chunks0 = 4 * (op.src[0].C / kBK);
    HSIDM_TRY(encode_phase_map(&tmA0, op.src[0].p, op.N, 2 * op.Hin, 2 * op.Win, op.src[0].C, 1, 1, C::kPW, kHaloRows));
    HSIDM_TRY(encode_phase_map(&tmA1, op.src[0].p, op.N, 2 * op.Hin, 2 * op.Win, op.src[0].C, 1, 0, C::kPW, kHaloRows));
  } else
  HSIDM_TRY(encode_act_map(&tmA0, op.src[0].p, op.N, op.Hin, op.Win, op.src[0].C, C::kPW, kHaloRows, 1));
  if (NT == 0) {
  } else if (p.chunks1)
    HSIDM_TRY(encode_act_map(&tmA1, op.src[1].p, op.N, op.Hin, op.Win, op.src[1].C, C::kPW, kHaloRows, 1));
  else
    tmA1 = tmA0;
  CUtensorMap tmR0 = tmA0, tmR1 = tmA0;
  if (NT == 0) {
    HSIDM_TRY(encode_phase_map(&tmR0, op.src[0].p, op.N, 2 * op.Hin, 2 * op.Win, op.src[0].C, 0, 1, C::kPW, kHaloRows));
    HSIDM_TRY(encode_phase_map(&tmR1, op.src[0].p, op.N, 2 * op.Hin, 2 * op.Win, op.src[0].C, 0, 0, C::kPW, kHaloRows));
  }
  if (p.rchunks0) HSIDM_TRY(encode_act_map(&tmR0, op.rsrc[0].p, op.N, op.Hin, op.Win, op.rsrc[0].C, 8 * MT, kRows, 1));
  if (p.rchunks1) HSIDM_TRY(encode_act_map(&tmR1, op.rsrc[1].p, op.N, op.Hin, op.Win, op.rsrc[1].C, 8 * MT, kRows, 1));
  const int K = op.K();
  HSIDM_TRY(encode_weight_map(&tmB, op.w_bf16, K, p.n_tiles * BN, C::kBRows));
  CUtensorMap tmO = tmB;   // the BN = 16 instantiation (fp32 NCHW output) stores from registers and never reads it
  if (BN % 64 == 0) HSIDM_TRY(encode_out_map(&tmO, op.out, op.N, op.Hout, op.Wout, op.Cout, p.oscale, p.oy, p.ox));
  char tag[120];
  snprintf(tag, sizeof(tag), "halo%s MT%d BN%d cin%d+%d cout%d %dx%d n%d%s%s%s%s%s", PAIR ? "2" : "", MT, BN, op.src[0].C, op.src[1].C, op.Cout,
           op.Hin, op.Win, op.N, op.resid ? " +res" : "", op.nbias ? " +nb" : "", op.stats_out ? " +st" : "", op.up_parity >= 0 ? " up2x" : op.s2 ? " s2" : "",
           op.gn.on() ? " +gn" : "");
  // algorithmic FLOPs: for the sub-pixel form, the share of the reference's 3x3 conv over the upsampled tensor
  const double flops = op.up_parity >= 0 ? 2.0 * op.N * op.Hin * (double)op.Win * op.Cout * 9 * (op.src[0].C + op.src[1].C)
                                         : 2.0 * op.N * op.Hin * (double)op.Win * op.Cout * K;
  ProfScope prof(PROF_CONV_TC, flops, stream, tag);
  if constexpr (PAIR) {
    // one cluster of two CTAs per pair of adjacent pixel tiles; an even grid of at most one CTA per SM
    const int pairs = pt.ctas;
    HSIDM_CUDA(launch_pdl(conv_halo_kernel<MT, BN, NT, true>, dim3(2 * pairs), dim3(halo_threads(BN)), C::kSmemBytes, stream, 2, tmA0, tmA1,
                          tmR0, tmR1, tmB, tmO, p));
  } else {
    const int grid = pt.ctas;
    HSIDM_CUDA(launch_pdl(conv_halo_kernel<MT, BN, NT, false>, dim3(grid), dim3(halo_threads(BN)), C::kSmemBytes, stream, 1, tmA0, tmA1, tmR0,
                          tmR1, tmB, tmO, p));
  }
  return after_launch("conv_halo_kernel");
}

template <int MT, int BN, int NT, bool PAIR>
int set_smem() {
  HSIDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<MT, BN, NT, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCfg<MT, BN, PAIR>::kSmemBytes));
  return HSIDM_OK;
}

}  // namespace

int conv_halo_init() {
  HSIDM_TRY((set_smem<2, 128, 9, false>()));
  HSIDM_TRY((set_smem<2, 128, 4, false>()));
  HSIDM_TRY((set_smem<2, 128, 1, false>()));
  HSIDM_TRY((set_smem<4, 64, 9, false>()));
  HSIDM_TRY((set_smem<4, 64, 4, false>()));
  HSIDM_TRY((set_smem<4, 64, 1, false>()));
  HSIDM_TRY((set_smem<2, 64, 9, false>()));
  HSIDM_TRY((set_smem<2, 64, 4, false>()));
  HSIDM_TRY((set_smem<2, 64, 1, false>()));
  HSIDM_TRY((set_smem<4, 16, 9, false>()));
  HSIDM_TRY((set_smem<2, 128, 0, false>()));
  HSIDM_TRY((set_smem<4, 64, 0, false>()));
  HSIDM_TRY((set_smem<2, 64, 0, false>()));
  HSIDM_TRY((set_smem<1, 128, 9, false>()));
  HSIDM_TRY((set_smem<1, 128, 4, false>()));
  HSIDM_TRY((set_smem<1, 64, 9, false>()));
  HSIDM_TRY((set_smem<1, 64, 4, false>()));
  HSIDM_TRY((set_smem<1, 256, 9, true>()));
  HSIDM_TRY((set_smem<1, 256, 4, true>()));
  HSIDM_TRY((set_smem<2, 128, 9, true>()));
  HSIDM_TRY((set_smem<2, 128, 4, true>()));
  HSIDM_TRY((set_smem<4, 64, 9, true>()));
  HSIDM_TRY((set_smem<4, 64, 4, true>()));
  return HSIDM_OK;
}

void conv_halo_set_timing(long long* device_counters) { tc::host().halo_dbg = device_counters; }

int conv_halo_stats_slots(const ConvOp& op) {
  int MT, BN;
  bool pair;
  pick_shape(op, &MT, &BN, &pair);
  if (MT == 0 || BN % 64 || op.out_layout != L_NHWC) return 0;
  // one slot per CTA segment of an (image, channel tile) group (x 2 CTAs of a pair, x 4 output parities of the sub-pixel form)
  return partition(op, MT, BN, pair).seg_max * (pair ? 2 : 1) * (op.up_parity >= 0 ? 4 : 1);
}

// Assumes conv_tc_supported(op) already holds (bf16 NHWC sources with 64-multiple channels, Cout fits an N tile).
bool conv_halo_supported(const ConvOp& op) {
  int MT, BN;
  bool pair;
  pick_shape(op, &MT, &BN, &pair);
  if (MT == 0) return false;
  for (int i = 0; i < 2; ++i)
    if (op.rsrc[i].C && (op.rsrc[i].C % kBK || op.rsrc[i].layout != L_NHWC)) return false;
  if (op.up_parity >= 0 && (BN % 64 || op.rsrc[0].C || op.resid)) return false;
  // the packed weight rows are padded to pick_bn(Cout); the halo kernel's BN must divide that padding
  return tc::pick_bn(op.Cout) % BN == 0;
}

int conv_halo(const ConvOp& op, cudaStream_t stream) {
  int MT, BN;
  bool pair;
  pick_shape(op, &MT, &BN, &pair);
  const bool sub = op.up_parity >= 0, one = op.ksize == 1;
  if (pair) {
    if (MT == 1 && BN == 256) return sub ? launch<1, 256, 4, true>(op, stream) : launch<1, 256, 9, true>(op, stream);
    if (MT == 2 && BN == 128) return sub ? launch<2, 128, 4, true>(op, stream) : launch<2, 128, 9, true>(op, stream);
    if (MT == 4 && BN == 64) return sub ? launch<4, 64, 4, true>(op, stream) : launch<4, 64, 9, true>(op, stream);
  } else if (op.s2) {
    if (MT == 2 && BN == 128) return launch<2, 128, 0, false>(op, stream);
    if (MT == 4 && BN == 64) return launch<4, 64, 0, false>(op, stream);
    if (MT == 2 && BN == 64) return launch<2, 64, 0, false>(op, stream);
  } else {
    if (MT == 2 && BN == 128)
      return one ? launch<2, 128, 1, false>(op, stream) : sub ? launch<2, 128, 4, false>(op, stream) : launch<2, 128, 9, false>(op, stream);
    if (MT == 2 && BN == 64)
      return one ? launch<2, 64, 1, false>(op, stream) : sub ? launch<2, 64, 4, false>(op, stream) : launch<2, 64, 9, false>(op, stream);
    if (MT == 4 && BN == 64)
      return one ? launch<4, 64, 1, false>(op, stream) : sub ? launch<4, 64, 4, false>(op, stream) : launch<4, 64, 9, false>(op, stream);
    if (MT == 4 && BN == 16 && !sub && !one) return launch<4, 16, 9, false>(op, stream);
    if (MT == 1 && BN == 128 && !one) return sub ? launch<1, 128, 4, false>(op, stream) : launch<1, 128, 9, false>(op, stream);
    if (MT == 1 && BN == 64 && !one) return sub ? launch<1, 64, 4, false>(op, stream) : launch<1, 64, 9, false>(op, stream);
  }
  HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "conv_halo: unsupported shape");
}

}  // namespace hsidm
