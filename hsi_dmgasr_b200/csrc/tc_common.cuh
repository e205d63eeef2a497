// Shared device/host pieces of the tcgen05 convolution kernels (conv_tc.cu: per-tap TMA boxes; conv_halo.cu: halo tile
// resident in shared memory and reused by all nine taps).
#pragma once
#include <cuda.h>

#include "kernels.cuh"

namespace hsidm {
namespace tc {

constexpr int kBM = 128;       // UMMA M
constexpr int kBK = 64;        // bf16 elements per k-block = one 128-byte swizzle line
constexpr uint32_t kSpinLimit = 1u << 24;

// Everything the epilogue needs to turn an fp32 accumulator row into output pixels.
struct EpiP {
  int N_img, H, W, Cout;
  const float* bias;
  const float* nbias;
  long long nbs;
  const int* nb_t;
  long long nb_ts;
  int act;
  float scale;
  const bf16* resid;
  void* out;
  int out_layout, clamp01;
  float* stats;      // optional [N][slots][Cout][2] per-slot (sum, sumsq) of the stored bf16 outputs, for the consumer's GroupNorm
  int stats_slots;   // slots per image
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

static __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
static __device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
static __device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a descriptor or protocol bug must not hang the GPU box.  Returns false on timeout.
static __device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  for (uint32_t i = 0; i < kSpinLimit; ++i)
    if (mbar_try_wait(bar, parity)) return true;
  if (err) atomicExch(err, code);
  return false;
}
static __device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
static __device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
static __device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

static __device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
static __device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
static __device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

static __device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
static __device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
static __device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
static __device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
static __device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
static __device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA pair (cta_group::2): two CTAs of a cluster issue one M = 256 MMA over both halves ----
static __device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
static __device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
static __device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
static __device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
static __device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
static __device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
static __device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all MMAs issued so far arrives on the barrier at this offset in BOTH CTAs of the pair
static __device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// 2-D TMA load into THIS CTA's shared memory whose bytes complete on a barrier of the pair's leader (cluster address)
static __device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row groups are 1024 B apart (SBO), LBO unused (=1),
// descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B).
static __device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24.
static __host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// 16 consecutive output channels [co0, co0+16) of pixel (n, y, x): +bias +noise bias -> act -> *scale -> +resid -> store.
static __device__ __forceinline__ void epilogue16(const EpiP& p, const float* nbias, int n, int y, int x, int co0,
                                                  const uint32_t (&r)[16]) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  if (p.out_layout == L_NHWC) {
    const long long m = ((long long)n * p.H + y) * p.W + x;
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + j));
        v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
      }
    }
    if (nbias) {
      const float* nb = nbias + (long long)n * p.nbs + co0;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(nb + j));
        v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
      }
    }
    if (p.act == ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : 0.01f * v[j];
    }
    if (p.scale != 1.0f) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= p.scale;
    }
    if (p.resid) {
      const bf16* rp = p.resid + m * p.Cout + co0;
      float a[8], b[8];
      load8(rp, a);
      load8(rp + 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += a[j], v[8 + j] += b[j];
    }
    bf16* op = static_cast<bf16*>(p.out) + m * p.Cout + co0;
    float lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) lo[j] = v[j], hi[j] = v[8 + j];
    store8(op, lo);
    store8(op + 8, hi);
  } else {
    // fp32 NCHW (last UNet layer / GAE outputs): a handful of channels, scalar tail-safe path
    const long long HW = (long long)p.H * p.W;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int co = co0 + j;
      if (co < p.Cout) {
        float o = v[j];
        if (p.bias) o += __ldg(p.bias + co);
        if (nbias) o += __ldg(nbias + (long long)n * p.nbs + co);
        if (p.act == ACT_LRELU) o = o > 0.f ? o : 0.01f * o;
        o *= p.scale;
        if (p.clamp01) o = fminf(fmaxf(o, 0.f), 1.f);
        static_cast<float*>(p.out)[((long long)n * p.Cout + co) * HW + (long long)y * p.W + x] = o;
      }
    }
  }
}

// Coalesced epilogue for 64 consecutive output channels [co0, co0+64) of the 32 accumulator rows one warp owns.
// `taddr` addresses column 0 of those 64 columns for this warp's TMEM lane quarter; `stage` is this warp's 4 KB
// shared-memory buffer (32 rows x 128 B, 16-byte chunks XOR-swizzled by row to stay bank-conflict free);
// pix(R) maps accumulator row R (0..127) to (image n, flat pixel index m).  Global traffic is fully coalesced:
// 8 lanes cover the 128 contiguous bytes of one pixel, 4 pixels per instruction, for both the residual read and
// the output write; the per-thread row work happens in registers in between.  NHWC bf16 output only.
// If p.stats is set, `st` (x,y = sums of channels co0+2*lane, +1; z,w = sums of squares) accumulates this warp's 32 rows.
template <typename PixFn>
static __device__ __forceinline__ void epilogue_rows64(const EpiP& p, const float* nbias, uint32_t taddr, int quarter,
                                                       int lane, int co0, uint4* stage, PixFn pix, float4& st) {
  const int sub = lane >> 3, chunk = lane & 7;
  // phase 1: residual tile -> staging (coalesced)
  if (p.resid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + sub;
      int n;
      long long m;
      pix(quarter * 32 + r, n, m);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (n < p.N_img) v = __ldg(reinterpret_cast<const uint4*>(p.resid + m * p.Cout + co0) + chunk);
      stage[r * 8 + (chunk ^ (r & 7))] = v;
    }
    __syncwarp();
  }
  // phase 2: own row in registers
  {
    int n;
    long long m;
    pix(quarter * 32 + lane, n, m);
    uint32_t acc[4][16];
#pragma unroll
    for (int q = 0; q < 4; ++q) tmem_ld16(taddr + q * 16, acc[q]);
    tmem_ld_wait();
    const float* nb = nbias ? nbias + (long long)min(n, p.N_img - 1) * p.nbs + co0 : nullptr;   // rows past the batch are masked, not read
#pragma unroll
    for (int c = 0; c < 8; ++c) {   // 8 channels per 16-byte chunk
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[c >> 1][(c & 1) * 8 + j]);
      if (p.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + c * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + c * 8 + 4));
        v[0] += b0.x, v[1] += b0.y, v[2] += b0.z, v[3] += b0.w, v[4] += b1.x, v[5] += b1.y, v[6] += b1.z, v[7] += b1.w;
      }
      if (nb && n < p.N_img) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(nb + c * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(nb + c * 8 + 4));
        v[0] += b0.x, v[1] += b0.y, v[2] += b0.z, v[3] += b0.w, v[4] += b1.x, v[5] += b1.y, v[6] += b1.z, v[7] += b1.w;
      }
      if (p.act == ACT_LRELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : 0.01f * v[j];
      }
      if (p.scale != 1.0f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= p.scale;
      }
      uint4* slot = &stage[lane * 8 + (c ^ (lane & 7))];
      if (p.resid) {
        const uint4 rv = *slot;
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          v[2 * j] += f.x, v[2 * j + 1] += f.y;
        }
      }
      uint4 o;
      __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      if (n >= p.N_img) o = make_uint4(0, 0, 0, 0);   // rows past the batch must not reach the statistics
      *slot = o;
    }
    __syncwarp();
  }
  // phase 2b: column sums of the stored values (lane -> channel pair 2*lane, 2*lane+1); conflict-free word reads
  if (p.stats) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(stage);
    const int cw = lane >> 2, ww = lane & 3;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const uint32_t u = w[r * 32 + ((cw ^ (r & 7)) << 2) + ww];
      const float a = __uint_as_float(u << 16), b = __uint_as_float(u & 0xffff0000u);
      st.x += a, st.y += b, st.z = fmaf(a, a, st.z), st.w = fmaf(b, b, st.w);
    }
  }
  // phase 3: staging -> global (coalesced)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + sub;
    int n;
    long long m;
    pix(quarter * 32 + r, n, m);
    const uint4 v = stage[r * 8 + (chunk ^ (r & 7))];
    if (n < p.N_img) *(reinterpret_cast<uint4*>(static_cast<bf16*>(p.out) + m * p.Cout + co0) + chunk) = v;
  }
  __syncwarp();
}

// ---- shared-memory accessors with 32-bit addresses (the staging buffer must not decay to generic LD/ST) ----
static __device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
static __device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
static __device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
static __device__ __forceinline__ float2 bf16x2_to_f2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// ---- TMA store of a staged tile ----
static __device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
static __device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
static __device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
static __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
static __device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Epilogue of the tensor-core conv kernels for 64 output channels of one warp's 32 accumulator rows:
// TMEM -> registers -> +bias (64 floats the warp parked in shared memory at `bias_smem`, or read from global `cb` /
// `cb2` when bias_smem is 0) -> act/scale -> +residual -> bf16 rows in the 128B-swizzled staging tile -> column
// statistics -> ONE TMA store of the tile: DIMS = 4, box {64 ch, 8 px, 4 rows, 1 image} of the output lattice map
// (halo kernel: the rows are 4 image rows x 8 pixels); DIMS = 2, box {64 ch, 32 pixels} of the [pixels, channels] view
// (per-tap kernel: the rows are 32 consecutive pixels; rows past the tensor are clipped by the TMA unit).
// valid_rows: accumulator rows that map to real pixels; the others are neither read as residual nor counted.
// Against per-lane stores the write-back costs one instruction instead of 8 LDS + 8 STG with 64-bit address arithmetic,
// and with the bias in shared memory instead of 64 registers the warp stays below the spill limit.
template <int DIMS = 4>
static __device__ __forceinline__ void epilogue_tma64(const EpiP& p, const CUtensorMap* tmO, uint32_t bias_smem, uint32_t taddr,
                                                      int lane, uint32_t stage, int c0, int c1, int c2, int c3,
                                                      const bf16* __restrict__ resid_lane, long long pitch, long long odd_off,
                                                      float4& st, const float* __restrict__ cb = nullptr,
                                                      const float* __restrict__ cb2 = nullptr, int valid_rows = 32) {
  if (lane == 0) tma_store_wait_read();   // the previous store out of `stage` must have read it
  __syncwarp();
  // phase 1 (residual layers only): residual tile -> staging, coalesced; resid_lane already points at this lane's
  // (row sub, 16-byte chunk) element of the tile
  const int sub = lane >> 3, chunk = lane & 7;
  if (resid_lane) {
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = (4 * i + sub) < valid_rows ? __ldg(reinterpret_cast<const uint4*>(resid_lane + (i >> 1) * pitch + (i & 1) * odd_off))
                                        : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r7 = (4 * (i & 1) + sub) & 7;
      sts128(stage + (uint32_t)((4 * i + sub) * 128 + ((chunk ^ r7) << 4)), v[i]);
    }
    __syncwarp();
  }
  // phase 2: own row in registers
  {
    uint32_t acc[4][16];
#pragma unroll
    for (int q = 0; q < 4; ++q) tmem_ld16(taddr + q * 16, acc[q]);
    tmem_ld_wait();
    const uint32_t my_row = stage + (uint32_t)(lane * 128);
    const int l7 = lane & 7;
#pragma unroll
    for (int c = 0; c < 8; ++c) {   // 8 channels per 16-byte chunk
      float2 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        v[j] = make_float2(__uint_as_float(acc[c >> 1][(c & 1) * 8 + 2 * j]), __uint_as_float(acc[c >> 1][(c & 1) * 8 + 2 * j + 1]));
      if (bias_smem) {
        const uint4 b0 = lds128(bias_smem + c * 32), b1 = lds128(bias_smem + c * 32 + 16);   // broadcast reads
        v[0] = __fadd2_rn(v[0], make_float2(__uint_as_float(b0.x), __uint_as_float(b0.y)));
        v[1] = __fadd2_rn(v[1], make_float2(__uint_as_float(b0.z), __uint_as_float(b0.w)));
        v[2] = __fadd2_rn(v[2], make_float2(__uint_as_float(b1.x), __uint_as_float(b1.y)));
        v[3] = __fadd2_rn(v[3], make_float2(__uint_as_float(b1.z), __uint_as_float(b1.w)));
      } else if (cb) {   // bias vectors straight from global memory (L1-resident: every lane reads the same 32 bytes)
        float4 b0 = __ldg(reinterpret_cast<const float4*>(cb + c0) + 2 * c), b1 = __ldg(reinterpret_cast<const float4*>(cb + c0) + 2 * c + 1);
        if (cb2) {
          const float4 e0 = __ldg(reinterpret_cast<const float4*>(cb2 + c0) + 2 * c), e1 = __ldg(reinterpret_cast<const float4*>(cb2 + c0) + 2 * c + 1);
          b0.x += e0.x, b0.y += e0.y, b0.z += e0.z, b0.w += e0.w, b1.x += e1.x, b1.y += e1.y, b1.z += e1.z, b1.w += e1.w;
        }
        v[0] = __fadd2_rn(v[0], make_float2(b0.x, b0.y));
        v[1] = __fadd2_rn(v[1], make_float2(b0.z, b0.w));
        v[2] = __fadd2_rn(v[2], make_float2(b1.x, b1.y));
        v[3] = __fadd2_rn(v[3], make_float2(b1.z, b1.w));
      }
      if (p.act == ACT_LRELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j].x = v[j].x > 0.f ? v[j].x : 0.01f * v[j].x, v[j].y = v[j].y > 0.f ? v[j].y : 0.01f * v[j].y;
      }
      if (p.scale != 1.0f) {
        const float2 sc = make_float2(p.scale, p.scale);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __fmul2_rn(v[j], sc);
      }
      const uint32_t slot = my_row + (uint32_t)((c ^ l7) << 4);
      if (resid_lane) {
        const uint4 rv = lds128(slot);
        v[0] = __fadd2_rn(v[0], bf16x2_to_f2(rv.x));
        v[1] = __fadd2_rn(v[1], bf16x2_to_f2(rv.y));
        v[2] = __fadd2_rn(v[2], bf16x2_to_f2(rv.z));
        v[3] = __fadd2_rn(v[3], bf16x2_to_f2(rv.w));
      }
      uint4 o;
      __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(v[j].x, v[j].y);
      if (lane >= valid_rows) o = make_uint4(0, 0, 0, 0);   // keep rows past the batch out of the statistics
      sts128(slot, o);
    }
    fence_proxy_async_smem();   // the rows just written are read by the TMA unit (async proxy)
    __syncwarp();
  }
  if (lane == 0) {
    if constexpr (DIMS == 4) tma_store_4d(tmO, stage, c0, c1, c2, c3);
    else tma_store_2d(tmO, stage, c0, c1);
  }
  // column sums of the stored values (lane -> channel pair 2*lane, 2*lane+1); conflict-free word reads
  if (p.stats) {
    const int cw = lane >> 2, ww = lane & 3;
    uint32_t base[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) base[k] = stage + (uint32_t)(((cw ^ k) << 4) + ww * 4);
    float2 s = make_float2(st.x, st.y), q = make_float2(st.z, st.w);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const float2 a = bf16x2_to_f2(lds32(base[r & 7] + r * 128));
      s = __fadd2_rn(s, a);
      q = __ffma2_rn(a, a, q);
    }
    st = make_float4(s.x, s.y, q.x, q.y);
  }
}

// Publishes one warp's accumulated statistics for 64 channels into its slot.
static __device__ __forceinline__ void stats_store(const EpiP& p, int n, int slot, int co0, int lane, const float4& st) {
  if (n < p.N_img)
    *reinterpret_cast<float4*>(p.stats + (((long long)n * p.stats_slots + slot) * p.Cout + co0 + 2 * lane) * 2) =
        make_float4(st.x, st.z, st.y, st.w);   // (sum, sumsq) of channel 2*lane, then of channel 2*lane+1
}

static __device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- host side shared state ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Host {
  EncodeTiledFn encode = nullptr;
  int num_sms = 0;
  int* err_flag = nullptr;
  int no_halo = 0;           // test knob: 1 -> always use the per-tap kernel of conv_tc.cu
  int variant = 0;           // test knob: kernel-variant selector for A/B runs (0 = default routing)
  int pairs_ok = 1;          // HSIDM_NO_PAIRS=1 (or a device without cluster launch) keeps every conv on single CTAs
  long long* halo_dbg = nullptr;   // developer timing probe of the halo kernel (device buffer, 8 counters per CTA)
};
Host& host();
// NHWC bf16 activation [N,H,W,C] as a 4-D tensor map with box {64, bw, bh, bn} and 128B swizzle (OOB reads are zero).
int encode_act_map(CUtensorMap* map, const void* base, int N, int H, int W, int C, int bw, int bh, int bn);
// Load map of the phase lattice (2i + py, 2j + px) of an NHWC bf16 tensor [N,H,W,C] (H, W even): a 4-D tensor
// {C, W/2, H/2, N} with box {64, bw, bh, 1}; used by the stride-2 form of the halo kernel.
int encode_phase_map(CUtensorMap* map, const void* base, int N, int H, int W, int C, int py, int px, int bw, int bh);
// Store map of an NHWC bf16 output seen as [pixels, channels]: 2-D tensor {C, N*H*W}, box {64, 32}, 128B swizzle.
int encode_rows_map(CUtensorMap* map, void* base, long long pixels, int C);
// Store map of an NHWC bf16 output [N,H,W,C] restricted to the pixel lattice (scale*i + oy, scale*j + ox): a 4-D tensor
// {C, W/scale, H/scale, N} with box {64, 8, 4, 1} (one epilogue warp's tile) and 128B swizzle.
int encode_out_map(CUtensorMap* map, void* base, int N, int H, int W, int C, int scale, int oy, int ox);
// K-major bf16 weights [rows][K] with box {64, bn_rows}.
int encode_weight_map(CUtensorMap* map, const void* base, int K, int rows, int bn_rows);
void fill_epilogue(EpiP* e, const ConvOp& op);
int pertap_stats_slots(const ConvOp& op);   // statistic slots per image the per-tap kernel produces (0 = cannot)
int pick_bn(int Cout);

}  // namespace tc

// conv_halo.cu
bool conv_halo_supported(const ConvOp& op);
int conv_halo_stats_slots(const ConvOp& op);
int conv_halo(const ConvOp& op, cudaStream_t stream);
int conv_halo_init();

}  // namespace hsidm
