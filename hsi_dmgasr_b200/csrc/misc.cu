// Small HBM-bound kernels around the contractions: fused DDPM posterior step, noise-level embedding,
// nearest-2x upsample, stride-2 im2col, row softmax, and the GAE channel-attention / overlap-average pieces.
#include <algorithm>

#include "kernels.cuh"

namespace hsidm {
namespace {

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10, counter-based: element e of loop index i of stream `seed` is reproducible and needs no state.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
}
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t step, uint64_t idx4) {
  uint32_t c[4] = {(uint32_t)idx4, (uint32_t)(idx4 >> 32), step, 0x9E3779B9u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  // Box-Muller on two pairs
  const float s = 2.3283064365386963e-10f;  // 2^-32
  float u0 = ((float)c[0] + 0.5f) * s, u1 = ((float)c[1] + 0.5f) * s;
  float u2 = ((float)c[2] + 0.5f) * s, u3 = ((float)c[3] + 0.5f) * s;
  u0 = fmaxf(u0, 1e-12f), u2 = fmaxf(u2, 1e-12f);
  float r0 = sqrtf(-2.0f * __logf(u0)), r1 = sqrtf(-2.0f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// One fused pass: read x_t, eps, (noise) ; write x_{t-1} (and a snapshot).  16 B/element fp32 traffic with a tape,
// 12 B/element with the in-kernel generator.  4 elements per thread, 128-bit accesses.
__global__ void posterior_kernel(PosteriorArgs a) {
  griddep_wait();
  griddep_launch();
  const int t = a.t >= 0 ? a.t : *a.t_dev;
  const float k_recip = a.coef[t * 5 + 0], k_recipm1 = a.coef[t * 5 + 1];
  const float k_c1 = a.coef[t * 5 + 2], k_c2 = a.coef[t * 5 + 3], k_sigma = a.coef[t * 5 + 4];
  const bool add_noise = t > 0 && (a.noise != nullptr || a.use_philox);
  float* snap = nullptr;
  if (a.snapshot_base && (t % a.inter) == 0) {
    // snapshots are stored in loop order: the k-th loop index with i % inter == 0, counting down from T-1
    const int k = (a.T - 1) / a.inter - t / a.inter;
    snap = a.snapshot_base + (int64_t)k * a.n;
  }
  const int64_t n4 = a.n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a.x_t)[i];
    const float4 e = reinterpret_cast<const float4*>(a.eps)[i];
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (add_noise) {
      if (a.noise) {
        const int64_t el = i << 2;
        const int64_t img = el / a.per_image, r = el - img * a.per_image;
        const int64_t j = a.T - 1 - t;
        z = *reinterpret_cast<const float4*>(a.noise + img * a.tape_image_stride + j * a.tape_step_stride + r);
      } else {
        z = philox_normal4(a.seed_dev ? (uint64_t)a.seed_dev[0] : a.seed, (uint32_t)t, (uint64_t)i + (a.seed_dev ? (uint64_t)a.seed_dev[1] : 0ull));
      }
    }
    float4 o;
#define HSIDM_POST(f)                                           \
  {                                                             \
    float x0 = k_recip * x.f - k_recipm1 * e.f;                 \
    x0 = fminf(fmaxf(x0, -1.f), 1.f);                           \
    o.f = (k_c1 * x0 + k_c2 * x.f) + z.f * k_sigma;             \
  }
    HSIDM_POST(x) HSIDM_POST(y) HSIDM_POST(z) HSIDM_POST(w)
#undef HSIDM_POST
    reinterpret_cast<float4*>(a.x_prev)[i] = o;
    if (snap) reinterpret_cast<float4*>(snap)[i] = o;
  }
}

__global__ void dec_kernel(int* t) {
  griddep_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) *t -= 1;
}
__global__ void state_set_kernel(int* state, int t, unsigned long long seed, unsigned long long offset4) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    state[0] = t;
    reinterpret_cast<unsigned long long*>(state + 2)[0] = seed;
    reinterpret_cast<unsigned long long*>(state + 2)[1] = offset4;
  }
}
__global__ void randn_kernel(float4* __restrict__ out, int64_t n4, unsigned long long seed, uint32_t step, unsigned long long first4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = philox_normal4(seed, step, first4 + (uint64_t)i);
}

// ---------------------------------------------------------------------------------------------------------
// noise embedding: one block per image. pe[dim] -> h[4*dim] -> t[dim] -> every layer's Linear(dim -> C_l).
__global__ void noise_embed_kernel(const float* __restrict__ level, int level_stride, int dim,
                                   const float* __restrict__ w1, const float* __restrict__ b1,
                                   const float* __restrict__ w3, const float* __restrict__ b3,
                                   const NoiseLayer* __restrict__ layers, int n_layers, int total,
                                   float* __restrict__ nbias) {
  extern __shared__ float sm[];  // pe[dim], h[4*dim], t[dim]
  float* pe = sm;
  float* h = sm + dim;
  float* tv = h + 4 * dim;
  const int n = blockIdx.x;
  const float lv = level[(int64_t)n * level_stride];
  const int half = dim / 2;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float step = (float)i / (float)half;
    const float enc = lv * expf(-9.210340371976184f * step);  // ln(1e4)
    pe[i] = sinf(enc);
    pe[half + i] = cosf(enc);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 4 * dim; o += blockDim.x) {
    float acc = b1[o];
    for (int k = 0; k < dim; ++k) acc = fmaf(w1[o * dim + k], pe[k], acc);
    h[o] = acc / (1.0f + expf(-acc));
  }
  __syncthreads();
  for (int o = threadIdx.x; o < dim; o += blockDim.x) {
    float acc = b3[o];
    for (int k = 0; k < 4 * dim; ++k) acc = fmaf(w3[o * 4 * dim + k], h[k], acc);
    tv[o] = acc;
  }
  __syncthreads();
  for (int l = 0; l < n_layers; ++l) {
    const NoiseLayer L = layers[l];
    for (int o = threadIdx.x; o < L.C; o += blockDim.x) {
      float acc = L.b[o];
      for (int k = 0; k < dim; ++k) acc = fmaf(L.w[o * dim + k], tv[k], acc);
      if (L.cbias) acc += L.cbias[o];
      nbias[(int64_t)n * total + L.off + o] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
template <typename AT>
__global__ void upsample2x_kernel(const AT* __restrict__ x, AT* __restrict__ out, int H, int W, int CV, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int ox = (int)(r % (2 * W));
    r /= (2 * W);
    const int oy = (int)(r % (2 * H));
    const int64_t n = r / (2 * H);
    float v[8];
    load8(x + ((n * H + (oy >> 1)) * W + (ox >> 1)) * (int64_t)CV * 8 + cv * 8, v);
    store8(out + i * 8, v);
  }
}

// out[n, oy, ox, tap*C + c] = x[n, 2*oy+ky-1, 2*ox+kx-1, c] (zero outside)
__global__ void im2col_s2_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int H, int W, int CV,
                                 int64_t total) {
  griddep_wait();
  griddep_launch();
  const int Ho = H / 2, Wo = W / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int tap = (int)(r % 9);
    r /= 9;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho);
    const int64_t n = r / Ho;
    const int iy = 2 * oy + tap / 3 - 1, ix = 2 * ox + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = *reinterpret_cast<const uint4*>(x + ((n * H + iy) * W + ix) * (int64_t)CV * 8 + cv * 8);
    *reinterpret_cast<uint4*>(out + i * 8) = v;
  }
}

// one thread per output pixel gathers its row (64 bf16 = 128 B): the 3x3 neighbourhood of every source plane is read with
// coalesced loads (x is the fastest index across threads) and three row pointers per plane.  The block's 128 rows are
// contiguous in the output, so they go out through shared memory as whole 512-byte runs per warp instruction (a thread
// storing its own row writes 16 bytes into each of 32 different lines per instruction: 2.8 TB/s measured).
template <int CT>
__global__ void __launch_bounds__(128)
im2col_small_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1, bf16* __restrict__ out,
                    int H, int W, int64_t total) {
  __shared__ uint4 tile[128][8];   // [row][16-byte chunk ^ (row & 7)]
  griddep_wait();
  griddep_launch();
  const int64_t plane = (int64_t)H * W;
  const int tid = threadIdx.x;
  for (int64_t base = blockIdx.x * (int64_t)128; base < total; base += (int64_t)gridDim.x * 128) {
    const int64_t i = base + tid;
    if (i < total) {
      const int x = (int)(i % W);
      const int y = (int)((i / W) % H);
      const int64_t n = i / plane;
      const bool xl = x > 0, xr = x + 1 < W, yt = y > 0, yb = y + 1 < H;
      float nb[9][CT];   // [tap][channel]
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const float* p = (c < C0 ? x0 + (n * C0 + c) * plane : x1 + (n * C1 + (c - C0)) * plane) + (int64_t)y * W + x;
        nb[0][c] = (yt && xl) ? __ldg(p - W - 1) : 0.f;
        nb[1][c] = yt ? __ldg(p - W) : 0.f;
        nb[2][c] = (yt && xr) ? __ldg(p - W + 1) : 0.f;
        nb[3][c] = xl ? __ldg(p - 1) : 0.f;
        nb[4][c] = __ldg(p);
        nb[5][c] = xr ? __ldg(p + 1) : 0.f;
        nb[6][c] = (yb && xl) ? __ldg(p + W - 1) : 0.f;
        nb[7][c] = yb ? __ldg(p + W) : 0.f;
        nb[8][c] = (yb && xr) ? __ldg(p + W + 1) : 0.f;
      }
      const float* flat = &nb[0][0];   // k = tap*CT + c, exactly the packed weight order
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = q * 8 + 2 * j;
          const float a = k < 9 * CT ? flat[k] : 0.f, b = k + 1 < 9 * CT ? flat[k + 1] : 0.f;
          __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
          w[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        tile[tid][q ^ (tid & 7)] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    __syncthreads();
    const int rows = (int)(total - base < 128 ? total - base : 128);
    uint4* o = reinterpret_cast<uint4*>(out + base * 64);
#pragma unroll 4
    for (int r = tid >> 3; r < rows; r += 16) o[r * 8 + (tid & 7)] = tile[r][(tid & 7) ^ (r & 7)];
    __syncthreads();
  }
}

// one warp per row, in place
__global__ void softmax_kernel(float* __restrict__ x, int64_t rows, int cols) {
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* p = x + row * cols;
  float m = -INFINITY;
  for (int i = lane; i < cols; i += 32) m = fmaxf(m, p[i]);
  m = warp_max(m);
  float s = 0.f;
  for (int i = lane; i < cols; i += 32) {
    const float e = expf(p[i] - m);
    p[i] = e;
    s += e;
  }
  s = warp_sum(s);
  const float inv = 1.0f / s;
  for (int i = lane; i < cols; i += 32) p[i] *= inv;
}

// one warp per row: fp32 scores in, normalised bf16 probabilities out (operand of the P*V tensor-core GEMM)
__global__ void softmax_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ out, int64_t rows, int cols) {
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* p = reinterpret_cast<const float4*>(x + row * cols);
  const int n4 = cols >> 2;
  float m = -INFINITY;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = p[i];
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  m = warp_max(m);
  float s = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = p[i];
    s += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
  }
  s = warp_sum(s);
  const float inv = 1.0f / s;
  uint2* o = reinterpret_cast<uint2*>(out + row * cols);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = p[i];
    uint2 w;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&w);
    h[0] = __floats2bfloat162_rn(__expf(v.x - m) * inv, __expf(v.y - m) * inv);
    h[1] = __floats2bfloat162_rn(__expf(v.z - m) * inv, __expf(v.w - m) * inv);
    o[i] = w;
  }
}

// ---------------------------------------------------------------------------------------------------------
// GAE: CALayer global average pool. One block per image (deterministic: fixed-order fold, no float atomics);
// thread -> (lane, channel), C <= 256.
template <typename AT>
__global__ void channel_mean_kernel(const AT* __restrict__ x, int HW, int C, float inv_hw, float* __restrict__ mean) {
  __shared__ float part[256];
  const int n = blockIdx.x;
  const int lanes = blockDim.x / C;
  const int c = threadIdx.x % C, lane = threadIdx.x / C;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (lane < lanes) {
    const AT* base = x + (int64_t)n * HW * C + c;
    int p = lane;
    for (; p + 3 * lanes < HW; p += 4 * lanes) {
      s0 += to_f32(base[(int64_t)p * C]);
      s1 += to_f32(base[(int64_t)(p + lanes) * C]);
      s2 += to_f32(base[(int64_t)(p + 2 * lanes) * C]);
      s3 += to_f32(base[(int64_t)(p + 3 * lanes) * C]);
    }
    for (; p < HW; p += lanes) s0 += to_f32(base[(int64_t)p * C]);
  }
  part[threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (threadIdx.x < C) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += part[l * C + threadIdx.x];
    mean[(int64_t)n * C + threadIdx.x] = acc * inv_hw;
  }
}

__global__ void ca_gate_kernel(const float* __restrict__ mean, int C, int Cr, const float* __restrict__ w0,
                               const float* __restrict__ b0, const float* __restrict__ w1,
                               const float* __restrict__ b1, float* __restrict__ gate) {
  extern __shared__ float sm[];  // m[C], h[Cr]
  float* m = sm;
  float* h = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) m[c] = mean[(int64_t)n * C + c];
  __syncthreads();
  for (int r = threadIdx.x; r < Cr; r += blockDim.x) {
    float acc = b0[r];
    for (int c = 0; c < C; ++c) acc = fmaf(w0[r * C + c], m[c], acc);
    h[r] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = b1[c];
    for (int r = 0; r < Cr; ++r) acc = fmaf(w1[c * Cr + r], h[r], acc);
    gate[(int64_t)n * C + c] = 1.0f / (1.0f + expf(-acc));
  }
}

template <typename AT>
__global__ void scale_residual_kernel(const AT* __restrict__ x, const float* __restrict__ gate, float scale,
                                      const AT* __restrict__ resid, AT* __restrict__ out, int64_t HW, int C,
                                      int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t n = i / ((int64_t)HW * C);
    const float g = gate ? gate[n * C + c] : 1.0f;
    out[i] = from_f32<AT>(to_f32(x[i]) * g * scale + to_f32(resid[i]));
  }
}

template <typename AT>
__global__ void overlap_average_kernel(const AT* __restrict__ dec, int G, int64_t HW, int n_subs, int n_colors,
                                       const int* __restrict__ start, const float* __restrict__ inv_count,
                                       AT* __restrict__ y, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int band = (int)(i % n_colors);
    const int64_t r = i / n_colors;
    const int64_t pix = r % HW, b = r / HW;
    float acc = 0.f;
    for (int g = 0; g < G; ++g) {
      const int j = band - start[g];
      if (j >= 0 && j < n_subs) acc += to_f32(dec[(((b * G + g) * HW) + pix) * n_subs + j]);
    }
    y[i] = from_f32<AT>(acc * inv_count[band]);
  }
}

inline int grid_for(int64_t work, int threads) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(work, threads), 148 * 32));
}

}  // namespace

int posterior_step(const PosteriorArgs& a, cudaStream_t stream) {
  if (a.n % 4 || a.per_image % 4) HSIDM_FAIL(HSIDM_BAD_SHAPE, "posterior_step: element counts must be multiples of 4");
  if (((uintptr_t)a.x_t | (uintptr_t)a.eps | (uintptr_t)a.x_prev | (uintptr_t)a.noise) & 15)
    HSIDM_FAIL(HSIDM_BAD_ARG, "posterior_step: pointers must be 16-byte aligned");
  ProfScope prof(PROF_POSTERIOR, (double)a.n * 4 * (a.noise ? 4 : 3), stream);
  HSIDM_CUDA(launch_pdl(posterior_kernel, dim3(grid_for(a.n / 4, 256)), dim3(256), 0, stream, 1, a));
  return after_launch("posterior_kernel");
}

int step_counter_dec(int* t_dev, cudaStream_t stream) {
  HSIDM_CUDA(launch_pdl(dec_kernel, dim3(1), dim3(32), 0, stream, 1, t_dev));
  return after_launch("dec_kernel");
}

int sampler_state_set(int* state, int t, uint64_t seed, uint64_t offset4, cudaStream_t stream) {
  state_set_kernel<<<1, 32, 0, stream>>>(state, t, (unsigned long long)seed, (unsigned long long)offset4);
  return after_launch("state_set_kernel");
}

int randn_fill(float* out, int64_t n, uint64_t seed, uint32_t step, uint64_t first4, cudaStream_t stream) {
  if (n % 4 || ((uintptr_t)out & 15)) HSIDM_FAIL(HSIDM_BAD_ARG, "randn_fill: count and pointer must be multiples of 4 floats");
  if (n == 0) return HSIDM_OK;
  randn_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(reinterpret_cast<float4*>(out), n / 4, (unsigned long long)seed, step,
                                                          (unsigned long long)first4);
  return after_launch("randn_kernel");
}

int noise_embed(const float* level, int level_stride, int n, int dim, const float* w1, const float* b1, const float* w3,
                const float* b3, const NoiseLayer* layers_dev, int n_layers, int total, float* nbias,
                cudaStream_t stream) {
  size_t smem = sizeof(float) * 6 * dim;
  noise_embed_kernel<<<n, 256, smem, stream>>>(level, level_stride, dim, w1, b1, w3, b3, layers_dev, n_layers, total,
                                               nbias);
  return after_launch("noise_embed_kernel");
}

int upsample2x(const void* x, void* out, int N, int H, int W, int C, int prec, cudaStream_t stream) {
  if (C % 8) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "upsample2x: C=%d not a multiple of 8", C);
  const int64_t total = (int64_t)N * 4 * H * W * (C / 8);
  ProfScope prof(PROF_OTHER, (double)total * 8 * (prec == HSIDM_BF16 ? 2 : 4) * 1.25, stream, "upsample2x");
  if (prec == HSIDM_BF16)
    upsample2x_kernel<bf16><<<grid_for(total, 256), 256, 0, stream>>>((const bf16*)x, (bf16*)out, H, W, C / 8, total);
  else
    upsample2x_kernel<float><<<grid_for(total, 256), 256, 0, stream>>>((const float*)x, (float*)out, H, W, C / 8, total);
  return after_launch("upsample2x_kernel");
}

int im2col_small(const float* x0, int C0, const float* x1, int C1, void* out, int N, int H, int W, cudaStream_t stream) {
  if (9 * (C0 + C1) > 64) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "im2col_small: 9*(%d+%d) > 64", C0, C1);
  const int64_t total = (int64_t)N * H * W;
  ProfScope prof(PROF_OTHER, (double)total * (128 + 4.0 * (C0 + C1)), stream, "im2col_small");
  bf16* o = static_cast<bf16*>(out);
  const int grid = grid_for(total, 128);
  switch (C0 + C1) {
    case 1: HSIDM_CUDA(launch_pdl(im2col_small_kernel<1>, dim3(grid), dim3(128), 0, stream, 1, x0, C0, x1, C1, o, H, W, total)); break;
    case 2: HSIDM_CUDA(launch_pdl(im2col_small_kernel<2>, dim3(grid), dim3(128), 0, stream, 1, x0, C0, x1, C1, o, H, W, total)); break;
    case 3: HSIDM_CUDA(launch_pdl(im2col_small_kernel<3>, dim3(grid), dim3(128), 0, stream, 1, x0, C0, x1, C1, o, H, W, total)); break;
    case 4: HSIDM_CUDA(launch_pdl(im2col_small_kernel<4>, dim3(grid), dim3(128), 0, stream, 1, x0, C0, x1, C1, o, H, W, total)); break;
    case 5: HSIDM_CUDA(launch_pdl(im2col_small_kernel<5>, dim3(grid), dim3(128), 0, stream, 1, x0, C0, x1, C1, o, H, W, total)); break;
    case 6: HSIDM_CUDA(launch_pdl(im2col_small_kernel<6>, dim3(grid), dim3(128), 0, stream, 1, x0, C0, x1, C1, o, H, W, total)); break;
    default: HSIDM_CUDA(launch_pdl(im2col_small_kernel<7>, dim3(grid), dim3(128), 0, stream, 1, x0, C0, x1, C1, o, H, W, total)); break;
  }
  return after_launch("im2col_small_kernel");
}

int im2col_s2(const void* x, void* out, int N, int H, int W, int C, cudaStream_t stream) {
  if (C % 8) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "im2col_s2: C=%d not a multiple of 8", C);
  const int64_t total = (int64_t)N * (H / 2) * (W / 2) * 9 * (C / 8);
  ProfScope prof(PROF_OTHER, (double)total * 16 * 1.45, stream, "im2col_s2");
  HSIDM_CUDA(launch_pdl(im2col_s2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, stream, 1, (const bf16*)x, (bf16*)out, H, W, C / 8, total));
  return after_launch("im2col_s2_kernel");
}

int softmax_rows(float* x, int64_t rows, int cols, cudaStream_t stream) {
  softmax_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, stream>>>(x, rows, cols);
  return after_launch("softmax_kernel");
}

int softmax_rows_bf16(const float* x, void* p_bf16, int64_t rows, int cols, cudaStream_t stream) {
  if (cols % 4) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "softmax_rows_bf16: cols=%d not a multiple of 4", cols);
  ProfScope prof(PROF_OTHER, (double)rows * cols * 6, stream, "softmax");
  softmax_bf16_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, stream>>>(x, static_cast<bf16*>(p_bf16), rows, cols);
  return after_launch("softmax_bf16_kernel");
}

int channel_mean(const void* x, int N, int HW, int C, float* mean, int prec, cudaStream_t stream) {
  if (C > 256) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "channel_mean: C=%d > 256", C);
  if (prec == HSIDM_BF16)
    channel_mean_kernel<bf16><<<N, 256, 0, stream>>>((const bf16*)x, HW, C, 1.0f / HW, mean);
  else
    channel_mean_kernel<float><<<N, 256, 0, stream>>>((const float*)x, HW, C, 1.0f / HW, mean);
  return after_launch("channel_mean_kernel");
}

int ca_gate(const float* mean, int N, int C, int Cr, const float* w0, const float* b0, const float* w1, const float* b1,
            float* gate, cudaStream_t stream) {
  ca_gate_kernel<<<N, 128, sizeof(float) * (C + Cr), stream>>>(mean, C, Cr, w0, b0, w1, b1, gate);
  return after_launch("ca_gate_kernel");
}

int scale_residual(const void* x, const float* gate, float scale, const void* resid, void* out, int N, int HW, int C,
                   int prec, cudaStream_t stream) {
  const int64_t total = (int64_t)N * HW * C;
  if (prec == HSIDM_BF16)
    scale_residual_kernel<bf16><<<grid_for(total, 256), 256, 0, stream>>>((const bf16*)x, gate, scale, (const bf16*)resid,
                                                                          (bf16*)out, HW, C, total);
  else
    scale_residual_kernel<float><<<grid_for(total, 256), 256, 0, stream>>>((const float*)x, gate, scale,
                                                                           (const float*)resid, (float*)out, HW, C, total);
  return after_launch("scale_residual_kernel");
}

int overlap_average(const void* dec, int B, int G, int HW, int n_subs, int n_colors, const int* start_dev,
                    const float* inv_count_dev, void* y, int prec, cudaStream_t stream) {
  const int64_t total = (int64_t)B * HW * n_colors;
  if (prec == HSIDM_BF16)
    overlap_average_kernel<bf16><<<grid_for(total, 256), 256, 0, stream>>>((const bf16*)dec, G, HW, n_subs, n_colors,
                                                                           start_dev, inv_count_dev, (bf16*)y, total);
  else
    overlap_average_kernel<float><<<grid_for(total, 256), 256, 0, stream>>>((const float*)dec, G, HW, n_subs, n_colors,
                                                                            start_dev, inv_count_dev, (float*)y, total);
  return after_launch("overlap_average_kernel");
}

}  // namespace hsidm
