// Training step of the SR3 UNet behind hsidm_train_forward / hsidm_train_backward.
//
// Reference: GaussianDiffusion.p_losses (model/sr3_modules/diffusion.py:222-250) = q_sample at a per-sample continuous
// noise level -> denoise_fn(cat([SR, x_noisy]), level) -> L1 / L2 sum against the injected noise, and
// DDPM.optimize_parameters (model/model.py:49-59) = loss.sum()/(b*c*h*w) -> backward -> Adam step.
//
// This first version is fp32 on CUDA cores end to end (parity gate: loss 1e-5, gradients 2e-4 against the unmodified
// reference, tests/golden/train_step.npz): the forward reuses the fp32 implicit-GEMM convolution and batched GEMM of the
// inference path and saves what the backward needs; the backward is hand-written - data gradients as convolutions with
// flipped/transposed weights, weight gradients as a split-M implicit GEMM with a fixed-order fold, GroupNorm+Swish(+dropout)
// backward, softmax / attention backward, the noise-level MLP.  No atomics anywhere: every reduction has a fixed order, so a
// step is bit-reproducible.  The tensor-core (bf16 dgrad / wgrad on the halo kernel) version is the next step.
//
// Memory: activations and their gradients live in two bump arenas sized by a dry pass (nothing is freed inside a step).
#include <cmath>
#include <functional>
#include <vector>

#include "unet.cuh"

namespace hsidm {
namespace {

constexpr float kGnEps = 1e-5f;

// ---- tensors and arenas -----------------------------------------------------------------------------------------------
struct T4 {          // NHWC fp32 activation and its gradient buffer (in the gradient arena)
  float* p = nullptr;
  float* g = nullptr;
  int N = 0, H = 0, W = 0, C = 0;
  int64_t numel() const { return (int64_t)N * H * W * C; }
};

struct Bump {
  char* base = nullptr;
  int64_t cap = 0, top = 0, peak = 0;
  bool dry = false;
  int64_t take(int64_t bytes) {
    bytes = round_up(bytes < 1 ? 1 : bytes, 256);
    const int64_t off = top;
    top += bytes;
    if (top > peak) peak = top;
    return off;
  }
};

struct TrainState {
  Bump act, grad;                 // activations + scratch / gradients of the activations (zeroed at the start of a backward)
  int N = 0, H = 0, W = 0;
  std::vector<std::function<void()>> tape;   // backward closures, replayed in reverse
  // per-step inputs / outputs
  float* loss_parts = nullptr;    // device: per-block partial sums of the loss
  int loss_blocks = 0;
  float* loss_sum_dev = nullptr;
  float* grads = nullptr;         // caller's gradient slab (ParamStore layout)
  int loss_type = 0;
  float inv_count = 0.f;          // 1 / (b*c*h*w)
  T4 eps;                         // network output
  const float* noise = nullptr;   // caller's noise (NCHW)
  uint64_t dropout_seed = 0;
  int dropout_layer = 0;
  cudaStream_t stream = nullptr;
  bool have_forward = false;
  int status = HSIDM_OK;
};

void free_train(void* p) {
  TrainState* t = static_cast<TrainState*>(p);
  if (t->act.base) cudaFree(t->act.base);
  if (t->grad.base) cudaFree(t->grad.base);
  if (t->loss_parts) cudaFree(t->loss_parts);
  if (t->loss_sum_dev) cudaFree(t->loss_sum_dev);
  delete t;
}

struct Ctx {   // what the layer functions need
  hsidm_ctx* c;
  TrainState* t;
  bool dry() const { return t->act.dry; }
  cudaStream_t st() const { return t->stream; }
  T4 alloc(int N, int H, int W, int C) {
    T4 x;
    x.N = N, x.H = H, x.W = W, x.C = C;
    const int64_t off = t->act.take(x.numel() * 4);
    const int64_t goff = t->grad.take(x.numel() * 4);
    if (!dry()) x.p = reinterpret_cast<float*>(t->act.base + off), x.g = reinterpret_cast<float*>(t->grad.base + goff);
    return x;
  }
  float* scratch(int64_t floats) {   // activation-arena scratch without a gradient twin
    const int64_t off = t->act.take(floats * 4);
    return dry() ? nullptr : reinterpret_cast<float*>(t->act.base + off);
  }
  template <typename F>
  void run(F&& f) {
    if (dry() || t->status != HSIDM_OK) return;
    const int s = f();
    if (s != HSIDM_OK) t->status = s;
  }
  template <typename F>
  void back(F&& f) {
    if (!dry()) t->tape.emplace_back(std::forward<F>(f));
  }
  float* grad_of(int param) const {   // slot of a parameter in the caller's gradient slab
    return t->grads + (c->ps.dev(param) - c->ps.dev(0));
  }
};

inline int grid1(int64_t n, int threads = 256) { return (int)std::min<int64_t>(ceil_div(n, threads), 4096); }

// ---- elementwise / layout kernels -------------------------------------------------------------------------------------
// x_in[n,y,x,0:3] = sr, [3:6] = level*hr + sqrt(1-level^2)*noise   (diffusion.py:213-220, :246-247); inputs NCHW
__global__ void qsample_cat_kernel(const float* __restrict__ hr, const float* __restrict__ sr, const float* __restrict__ noise,
                                   const float* __restrict__ level, int C, int HW, int64_t total, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (2 * C));
    const int64_t r = i / (2 * C);
    const int px = (int)(r % HW);
    const int n = (int)(r / HW);
    float v;
    if (c < C) {
      v = sr[((int64_t)n * C + c) * HW + px];
    } else {
      const int64_t j = ((int64_t)n * C + (c - C)) * HW + px;
      const float l = level[n];
      v = l * hr[j] + sqrtf(1.0f - l * l) * noise[j];
    }
    out[i] = v;
  }
}

__global__ void concat_kernel(const float* __restrict__ a, int Ca, const float* __restrict__ b, int Cb, int64_t pixels,
                              float* __restrict__ out) {
  const int C = Ca + Cb;
  const int64_t total = pixels * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t px = i / C;
    out[i] = c < Ca ? a[px * Ca + c] : b[px * Cb + (c - Ca)];
  }
}
// ga += gcat[:, :Ca], gb += gcat[:, Ca:]
__global__ void split_add_kernel(const float* __restrict__ gcat, float* __restrict__ ga, int Ca, float* __restrict__ gb, int Cb,
                                 int64_t pixels) {
  const int C = Ca + Cb;
  const int64_t total = pixels * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t px = i / C;
    if (c < Ca) ga[px * Ca + c] += gcat[i];
    else gb[px * Cb + (c - Ca)] += gcat[i];
  }
}
__global__ void add_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] += src[i];
}
// h[n,px,c] += nb[n,c]
__global__ void add_rowbias_kernel(float* __restrict__ h, const float* __restrict__ nb, int HW, int C, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int n = (int)(i / ((int64_t)HW * C));
    h[i] += nb[(int64_t)n * C + c];
  }
}
// z[n, 2y, 2x, c] = g[n, y, x, c], zero elsewhere: turns the data gradient of a stride-2 conv into a stride-1 conv
__global__ void zero_insert_kernel(const float* __restrict__ g, int Ho, int Wo, int C, int64_t total, float* __restrict__ z) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int x = (int)(r % (2 * Wo));
    r /= 2 * Wo;
    const int y = (int)(r % (2 * Ho));
    const int64_t n = r / (2 * Ho);
    z[i] = ((x | y) & 1) ? 0.f : g[((n * Ho + (y >> 1)) * Wo + (x >> 1)) * C + c];
  }
}
// gx[n,y,x,c] += sum of the 2x2 block of gu (backward of nearest-2x upsampling)
__global__ void sumpool2_add_kernel(const float* __restrict__ gu, int H, int W, int C, int64_t total, float* __restrict__ gx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const int64_t n = r / H;
    const float* b = gu + (((n * 2 * H + 2 * y) * 2 * W) + 2 * x) * C + c;
    gx[i] += (b[0] + b[C]) + (b[(int64_t)2 * W * C] + b[(int64_t)2 * W * C + C]);
  }
}

// ---- dropout mask: counter-based, so the backward regenerates it instead of storing it -------------------------------
__device__ __forceinline__ float keep_scale(uint64_t seed, uint32_t layer, uint64_t idx, float p) {
  if (p <= 0.f) return 1.0f;
  // splitmix-style hash of (seed, layer, idx) -> uniform in [0,1)
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1) + ((uint64_t)layer << 40);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
  return u < p ? 0.f : 1.0f / (1.0f - p);
}

// ---- GroupNorm(+Swish)(+dropout) forward / backward: one block per (group, image) ---------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];   // fixed order
  return t;
}

__global__ void __launch_bounds__(256) gn_fwd_train_kernel(const float* __restrict__ x, int HW, int C, int groups,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, int swish,
                                                           float drop_p, uint64_t seed, uint32_t layer, float* __restrict__ stats,
                                                           float* __restrict__ out) {
  __shared__ double sh[8];
  const int g = blockIdx.x, n = blockIdx.y, cpg = C / groups;
  const int64_t m = (int64_t)HW * cpg;
  const float* xb = x + (int64_t)n * HW * C + g * cpg;
  double s = 0.0, q = 0.0;
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
    const float v = xb[(i / cpg) * C + (i % cpg)];
    s += v, q += (double)v * v;
  }
  s = block_sum(s, sh);
  q = block_sum(q, sh);
  const double mean = s / (double)m;
  double var = q / (double)m - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + (double)kGnEps));
  if (threadIdx.x == 0) stats[((int64_t)n * groups + g) * 2] = mu, stats[((int64_t)n * groups + g) * 2 + 1] = rstd;
  float* ob = out + (int64_t)n * HW * C + g * cpg;
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
    const int j = (int)(i % cpg);
    const int64_t off = (i / cpg) * C + j;
    float y = (xb[off] - mu) * rstd * gamma[g * cpg + j] + beta[g * cpg + j];
    if (swish) y = y / (1.0f + __expf(-y));
    y *= keep_scale(seed, layer, (uint64_t)n * HW * C + off + g * cpg, drop_p);
    ob[off] = y;
  }
}

// dA = gradient w.r.t. the kernel's output; accumulates the input gradient into gx and writes per-image partial sums of the
// affine gradients: part[n][c] = (sum dy*xhat, sum dy).
__global__ void __launch_bounds__(256) gn_bwd_train_kernel(const float* __restrict__ x, const float* __restrict__ dA, int HW, int C,
                                                           int groups, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int swish, float drop_p, uint64_t seed, uint32_t layer,
                                                           const float* __restrict__ stats, float* __restrict__ gx,
                                                           float* __restrict__ part) {
  __shared__ double sh[8];
  extern __shared__ float chsum[];   // [lanes][cpg][2]
  const int g = blockIdx.x, n = blockIdx.y, cpg = C / groups;
  const int lanes = blockDim.x / cpg;
  const int j = threadIdx.x % cpg, lane = threadIdx.x / cpg;
  const bool active = lane < lanes;
  const float mu = stats[((int64_t)n * groups + g) * 2], rstd = stats[((int64_t)n * groups + g) * 2 + 1];
  const int64_t base = (int64_t)n * HW * C + g * cpg;
  const float gm = active ? gamma[g * cpg + j] : 0.f, bt = active ? beta[g * cpg + j] : 0.f;
  auto dy_of = [&](int px, float& xh) {
    const int64_t off = base + (int64_t)px * C + j;
    xh = (x[off] - mu) * rstd;
    const float y = xh * gm + bt;
    float d = dA[off] * keep_scale(seed, layer, (uint64_t)off, drop_p);
    if (swish) {
      const float sg = 1.0f / (1.0f + __expf(-y));
      d *= sg * (1.0f + y * (1.0f - sg));
    }
    return d;
  };
  double s1 = 0.0, s2 = 0.0;
  float cg = 0.f, cb = 0.f;
  if (active)
    for (int px = lane; px < HW; px += lanes) {
      float xh;
      const float d = dy_of(px, xh);
      s1 += (double)d * gm, s2 += (double)d * gm * xh;
      cg += d * xh, cb += d;
    }
  s1 = block_sum(s1, sh);
  s2 = block_sum(s2, sh);
  // per-channel partials: fold the lanes in order
  if (active) chsum[(lane * cpg + j) * 2] = cg, chsum[(lane * cpg + j) * 2 + 1] = cb;
  __syncthreads();
  if (threadIdx.x < cpg) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) a += chsum[(l * cpg + threadIdx.x) * 2], b += chsum[(l * cpg + threadIdx.x) * 2 + 1];
    part[((int64_t)n * C + g * cpg + threadIdx.x) * 2] = a;
    part[((int64_t)n * C + g * cpg + threadIdx.x) * 2 + 1] = b;
  }
  const float m_inv = 1.0f / ((float)HW * cpg);
  const float k1 = (float)s1 * m_inv, k2 = (float)s2 * m_inv;
  if (active)
    for (int px = lane; px < HW; px += lanes) {
      float xh;
      const float d = dy_of(px, xh);
      gx[base + (int64_t)px * C + j] += rstd * (d * gm - k1 - xh * k2);
    }
}

// out[c] (+)= sum_n in[n][c*stride + which]   (fixed order over n)
__global__ void reduce_rows_kernel(const float* __restrict__ in, int N, int C, int stride, int which, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int n = 0; n < N; ++n) a += in[((int64_t)n * C + c) * stride + which];
  out[c] = a;
}

// per-image column sums: out[n][c] = sum_px g[n,px,c]; block (32 channels x 8 pixel lanes)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ g, int HW, int C, float* __restrict__ out) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), lane = threadIdx.x >> 5, n = blockIdx.y;
  float a = 0.f;
  if (c < C)
    for (int px = lane; px < HW; px += 8) a += g[((int64_t)n * HW + px) * C + c];
  sh[lane][threadIdx.x & 31] = a;
  __syncthreads();
  if (lane == 0 && c < C) {
    float t = 0.f;
    for (int l = 0; l < 8; ++l) t += sh[l][threadIdx.x & 31];
    out[(int64_t)n * C + c] = t;
  }
}

// ---- weight gradient of a convolution: split-M implicit GEMM -----------------------------------------------------------
// part[split][tap][ci][co] = sum over this split's output pixels of A[n, iy, ix, ci] * dY[n, oy, ox, co]
struct WgP {
  const float* a;     // conv input NHWC [N, Hin, Win, Cin] (pre-upsampling when up)
  const float* dy;    // [N, Ho, Wo, Cout]
  int N, Hin, Win, Cin, Ho, Wo, Cout, ks, stride, up;
  int splits, pix_per_split;
  float* part;
};
__global__ void __launch_bounds__(256) wgrad_kernel(const WgP p) {
  __shared__ float As[16][33], Ds[16][33];
  const int co0 = blockIdx.x * 32;
  const int ci_tiles = (p.Cin + 31) / 32;
  const int tap = blockIdx.y / ci_tiles, ci0 = (blockIdx.y % ci_tiles) * 32;
  const int split = blockIdx.z;
  const int ky = tap / p.ks, kx = tap % p.ks, pad = p.ks >> 1;
  const int He = p.up ? 2 * p.Hin : p.Hin, We = p.up ? 2 * p.Win : p.Win;
  const int64_t M = (int64_t)p.N * p.Ho * p.Wo;
  const int64_t m_begin = (int64_t)split * p.pix_per_split, m_end = min(M, m_begin + p.pix_per_split);
  const int tc = threadIdx.x & 31, tr = threadIdx.x >> 5;   // tr: 8 rows of 4 input channels
  const int lp = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 2;   // loader: pixel lp (0..15), channels lc, lc+1
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t m0 = m_begin; m0 < m_end; m0 += 16) {
    const int64_t m = m0 + lp;
    float a0 = 0.f, a1 = 0.f, d0 = 0.f, d1 = 0.f;
    if (m < m_end) {
      const int n = (int)(m / ((int64_t)p.Ho * p.Wo));
      const int r = (int)(m - (int64_t)n * p.Ho * p.Wo);
      const int oy = r / p.Wo, ox = r - oy * p.Wo;
      int iy = oy * p.stride + ky - pad, ix = ox * p.stride + kx - pad;
      if (iy >= 0 && iy < He && ix >= 0 && ix < We) {
        if (p.up) iy >>= 1, ix >>= 1;
        const float* src = p.a + (((int64_t)n * p.Hin + iy) * p.Win + ix) * p.Cin;
        if (ci0 + lc < p.Cin) a0 = src[ci0 + lc];
        if (ci0 + lc + 1 < p.Cin) a1 = src[ci0 + lc + 1];
      }
      const float* dsrc = p.dy + m * p.Cout;
      if (co0 + lc < p.Cout) d0 = dsrc[co0 + lc];
      if (co0 + lc + 1 < p.Cout) d1 = dsrc[co0 + lc + 1];
    }
    __syncthreads();
    As[lp][lc] = a0, As[lp][lc + 1] = a1, Ds[lp][lc] = d0, Ds[lp][lc + 1] = d1;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float d = Ds[q][tc];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(As[q][tr * 4 + r], d, acc[r]);
    }
  }
  const int co = co0 + tc;
  if (co < p.Cout)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ci = ci0 + tr * 4 + r;
      if (ci < p.Cin) p.part[(((int64_t)split * p.ks * p.ks + tap) * p.Cin + ci) * p.Cout + co] = acc[r];
    }
}
// dW[co][ci][tap] = sum_split part[split][tap][ci][co]  (reference parameter layout)
__global__ void wgrad_fold_kernel(const float* __restrict__ part, int splits, int taps, int Cin, int Cout, float* __restrict__ dw) {
  const int64_t total = (int64_t)taps * Cin * Cout;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const int ci = (int)((i / Cout) % Cin);
    const int tap = (int)(i / ((int64_t)Cout * Cin));
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += part[(int64_t)s * total + i];
    dw[((int64_t)co * Cin + ci) * taps + tap] = a;
  }
}
// data-gradient weights: wd[(tap*Cout + co)*Cin + ci] = w[co][ci][taps-1-tap]  (a convolution of dY with the flipped kernel)
__global__ void pack_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, float* __restrict__ wd) {
  const int64_t total = (int64_t)taps * Cout * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int tap = (int)(i / ((int64_t)Cin * Cout));
    wd[i] = w[((int64_t)co * Cin + ci) * taps + (taps - 1 - tap)];
  }
}

// ---- batched C[k][n] = alpha * sum_m A[m][k] * B[m][n] ("TN"), fp32 ------------------------------------------------------
struct TnP {
  const float* A;
  const float* B;
  float* C;
  int M, K, N;
  int64_t lda, ldb, ldc, sA, sB, sC;
  float alpha;
};
__global__ void __launch_bounds__(256) gemm_tn_kernel(const TnP p) {
  __shared__ float As[16][33], Bs[16][33];
  const int b = blockIdx.z;
  const float* A = p.A + b * p.sA;
  const float* B = p.B + b * p.sB;
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tc = threadIdx.x & 31, tr = threadIdx.x >> 5;
  const int lp = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 2;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int m0 = 0; m0 < p.M; m0 += 16) {
    const int m = m0 + lp;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    if (m < p.M) {
      if (k0 + lc < p.K) a0 = A[(int64_t)m * p.lda + k0 + lc];
      if (k0 + lc + 1 < p.K) a1 = A[(int64_t)m * p.lda + k0 + lc + 1];
      if (n0 + lc < p.N) b0 = B[(int64_t)m * p.ldb + n0 + lc];
      if (n0 + lc + 1 < p.N) b1 = B[(int64_t)m * p.ldb + n0 + lc + 1];
    }
    __syncthreads();
    As[lp][lc] = a0, As[lp][lc + 1] = a1, Bs[lp][lc] = b0, Bs[lp][lc + 1] = b1;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float d = Bs[q][tc];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(As[q][tr * 4 + r], d, acc[r]);
    }
  }
  const int n = n0 + tc;
  if (n < p.N)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k = k0 + tr * 4 + r;
      if (k < p.K) p.C[b * p.sC + (int64_t)k * p.ldc + n] = p.alpha * acc[r];
    }
}

// dS = P * (dP - rowsum(dP * P)) * scale, in place on dP; one warp per row
__global__ void softmax_bwd_kernel(const float* __restrict__ P, float* __restrict__ dP, int64_t rows, int cols, float scale) {
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* p = P + row * cols;
  float* d = dP + row * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += d[c] * p[c];
  s = warp_sum(s);
  for (int c = lane; c < cols; c += 32) d[c] = p[c] * (d[c] - s) * scale;
}

// ---- noise-level embedding + MLP (unet.py:18-31, 182-187) and the per-layer FeatureWiseAffine linears (unet.py:34-50) --------
// one block per image: pe[dim] -> h1[4dim] -> a1 = swish(h1) -> t[dim]; saved for the backward
__global__ void noise_mlp_fwd_kernel(const float* __restrict__ level, int dim, const float* __restrict__ w1, const float* __restrict__ b1,
                                     const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ pe,
                                     float* __restrict__ h1, float* __restrict__ a1, float* __restrict__ temb) {
  extern __shared__ float sm[];   // pe[dim], a1[4dim]
  const int n = blockIdx.x, half = dim / 2;
  const float lv = level[n];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float e = lv * expf(-logf(1e4f) * ((float)i / (float)half));
    sm[i] = sinf(e), sm[half + i] = cosf(e);
    pe[(int64_t)n * dim + i] = sm[i], pe[(int64_t)n * dim + half + i] = sm[half + i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 4 * dim; j += blockDim.x) {
    float a = b1[j];
    for (int i = 0; i < dim; ++i) a = fmaf(w1[(int64_t)j * dim + i], sm[i], a);
    h1[(int64_t)n * 4 * dim + j] = a;
    sm[dim + j] = a / (1.0f + expf(-a));
    a1[(int64_t)n * 4 * dim + j] = sm[dim + j];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    float a = b3[d];
    for (int j = 0; j < 4 * dim; ++j) a = fmaf(w3[(int64_t)d * 4 * dim + j], sm[dim + j], a);
    temb[(int64_t)n * dim + d] = a;
  }
}
// out[n][o] = b[o] + sum_i W[o][i] * in[n][i]
__global__ void linear_fwd_kernel(const float* __restrict__ in, int N, int I, const float* __restrict__ W, const float* __restrict__ b, int O,
                                  float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * O) return;
  const int n = idx / O, o = idx % O;
  float a = b ? b[o] : 0.f;
  for (int i = 0; i < I; ++i) a = fmaf(W[(int64_t)o * I + i], in[(int64_t)n * I + i], a);
  out[idx] = a;
}
// dW[o][i] = sum_n dout[n][o] * in[n][i] ; db[o] = sum_n dout[n][o]
__global__ void linear_bwd_w_kernel(const float* __restrict__ dout, const float* __restrict__ in, int N, int I, int O, float* __restrict__ dW,
                                    float* __restrict__ db) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= O * (I + 1)) return;
  const int o = idx / (I + 1), i = idx % (I + 1);
  float a = 0.f;
  if (i < I) {
    for (int n = 0; n < N; ++n) a = fmaf(dout[(int64_t)n * O + o], in[(int64_t)n * I + i], a);
    dW[(int64_t)o * I + i] = a;
  } else {
    for (int n = 0; n < N; ++n) a += dout[(int64_t)n * O + o];
    if (db) db[o] = a;
  }
}
// din[n][i] += sum_o dout[n][o] * W[o][i]
__global__ void linear_bwd_x_kernel(const float* __restrict__ dout, const float* __restrict__ W, int N, int I, int O, float* __restrict__ din) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * I) return;
  const int n = idx / I, i = idx % I;
  float a = 0.f;
  for (int o = 0; o < O; ++o) a = fmaf(dout[(int64_t)n * O + o], W[(int64_t)o * I + i], a);
  din[idx] += a;
}
// dh1 = da1 * swish'(h1)
__global__ void swish_bwd_kernel(const float* __restrict__ h, float* __restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float y = h[i], sg = 1.0f / (1.0f + expf(-y));
    d[i] *= sg * (1.0f + y * (1.0f - sg));
  }
}

// ---- loss (diffusion.py:249, model.py:53-55): partial sums per block, gradient w.r.t. eps ----------------------------------
// eps NHWC [N,HW,C], noise NCHW; geps = dL/deps for L = sum(|noise-eps|)/count (l1) or sum((noise-eps)^2)/count (l2)
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ eps, const float* __restrict__ noise, int C, int HW, int64_t total,
                                                   int l2, float inv_count, float* __restrict__ parts, float* __restrict__ geps) {
  __shared__ double sh[8];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    const int px = (int)(r % HW);
    const int64_t n = r / HW;
    const float d = noise[(n * C + c) * HW + px] - eps[i];
    if (l2) {
      s += (double)d * d;
      geps[i] = -2.0f * d * inv_count;
    } else {
      s += fabsf(d);
      geps[i] = d > 0.f ? -inv_count : (d < 0.f ? inv_count : 0.f);
    }
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) parts[blockIdx.x] = (float)s;
}
__global__ void loss_fold_kernel(const float* __restrict__ parts, int n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += parts[i];
    out[0] = (float)s;
  }
}
__global__ void scale_kernel(float* __restrict__ x, float s, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= s;
}

// ======================================================================================================================
// layer functions: each enqueues its forward kernels and records the closure that enqueues its backward
// ======================================================================================================================
int launch_conv_fwd(const float* src, int N, int Hin, int Win, int Cin, int up, int stride, int ks, const float* w_f32, const float* bias,
                    int Cout, int Ho, int Wo, const float* resid, float* out, cudaStream_t st) {
  ConvOp op;
  op.src[0].p = src, op.src[0].C = Cin, op.src[0].layout = L_NHWC;
  op.N = N, op.Hin = Hin, op.Win = Win, op.up = up, op.stride = stride, op.ksize = ks;
  op.Hout = Ho, op.Wout = Wo, op.Cout = Cout;
  op.w_f32 = w_f32, op.bias = bias, op.resid = resid, op.out = out;
  return conv_simt(op, HSIDM_F32, st);
}

// y = conv(a) (+ bias); backward: ga += dgrad(gy), dW, db
T4 conv_layer(Ctx& cx, const T4& a, const ConvW& w, int stride = 1, int up = 0) {
  const int He = up ? 2 * a.H : a.H, We = up ? 2 * a.W : a.W;
  const int Ho = stride == 2 ? (He + 1) / 2 : He, Wo = stride == 2 ? (We + 1) / 2 : We;
  T4 y = cx.alloc(a.N, Ho, Wo, w.Cout);
  const float* bias = cx.c->ps.dev(w.pb);
  const int taps = w.ks * w.ks;
  cx.run([&] { return launch_conv_fwd(a.p, a.N, a.H, a.W, a.C, up, stride, w.ks, w.w_f32, bias, w.Cout, Ho, Wo, nullptr, y.p, cx.st()); });
  // backward scratch, reserved now so that the dry pass counts it
  float* wd = cx.scratch((int64_t)taps * w.Cout * w.Cin);
  const int64_t M = (int64_t)a.N * Ho * Wo;
  const int tiles = (int)(ceil_div(w.Cout, 32) * ceil_div(w.Cin, 32) * taps);
  int splits = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(592, tiles), ceil_div(M, 256)));
  const int pps = (int)round_up(ceil_div(M, splits), 16);
  splits = (int)ceil_div(M, pps);
  float* part = cx.scratch((int64_t)splits * taps * w.Cin * w.Cout);
  float* zins = stride == 2 ? cx.scratch((int64_t)a.N * a.H * a.W * w.Cout) : nullptr;
  float* gup = up ? cx.scratch((int64_t)a.N * He * We * a.C) : nullptr;
  float* bsum = w.pb >= 0 ? cx.scratch((int64_t)a.N * w.Cout) : nullptr;
  hsidm_ctx* c = cx.c;
  TrainState* t = cx.t;
  float* gw = cx.dry() ? nullptr : cx.grad_of(w.pw);
  float* gb = (cx.dry() || w.pb < 0) ? nullptr : cx.grad_of(w.pb);
  const bool need_dx = a.g != nullptr;
  cx.back([=]() {
    cudaStream_t st = t->stream;
    auto chk = [&](int s) {
      if (s != HSIDM_OK && t->status == HSIDM_OK) t->status = s;
    };
    // weight gradient
    WgP p{a.p, y.g, a.N, a.H, a.W, a.C, Ho, Wo, w.Cout, w.ks, stride, up, splits, pps, part};
    wgrad_kernel<<<dim3((unsigned)ceil_div(w.Cout, 32), (unsigned)(ceil_div(w.Cin, 32) * taps), (unsigned)splits), 256, 0, st>>>(p);
    {
      const int ws = after_launch("wgrad_kernel");
      if (ws != HSIDM_OK)
        set_last_error("wgrad_kernel launch failed: grid (%d,%d,%d) Cin %d Cout %d ks %d stride %d up %d N %d %dx%d -> %dx%d pps %d [%s]",
                       (int)ceil_div(w.Cout, 32), (int)(ceil_div(w.Cin, 32) * taps), splits, w.Cin, w.Cout, w.ks, stride, up, a.N, a.H, a.W, Ho,
                       Wo, pps, g_last_error.c_str());
      chk(ws);
    }
    wgrad_fold_kernel<<<grid1((int64_t)taps * w.Cin * w.Cout), 256, 0, st>>>(part, splits, taps, w.Cin, w.Cout, gw);
    chk(after_launch("wgrad_fold_kernel"));
    if (gb) {
      colsum_kernel<<<dim3((unsigned)ceil_div(w.Cout, 32), (unsigned)a.N), 256, 0, st>>>(y.g, Ho * Wo, w.Cout, bsum);
      chk(after_launch("colsum_kernel"));
      reduce_rows_kernel<<<(unsigned)ceil_div(w.Cout, 128), 128, 0, st>>>(bsum, a.N, w.Cout, 1, 0, gb);
      chk(after_launch("reduce_rows_kernel"));
    }
    if (!need_dx) return;
    // data gradient: a stride-1 convolution of (zero-inserted) gy with the flipped, transposed kernel
    pack_dgrad_kernel<<<grid1((int64_t)taps * w.Cout * w.Cin), 256, 0, st>>>(c->ps.dev(w.pw), w.Cout, w.Cin, taps, wd);
    chk(after_launch("pack_dgrad_kernel"));
    const float* gy = y.g;
    int Hg = Ho, Wg = Wo;
    if (stride == 2) {
      zero_insert_kernel<<<grid1((int64_t)a.N * a.H * a.W * w.Cout), 256, 0, st>>>(y.g, Ho, Wo, w.Cout, (int64_t)a.N * a.H * a.W * w.Cout, zins);
      chk(after_launch("zero_insert_kernel"));
      gy = zins, Hg = a.H, Wg = a.W;
    }
    if (up) {
      chk(launch_conv_fwd(gy, a.N, Hg, Wg, w.Cout, 0, 1, w.ks, wd, nullptr, a.C, Hg, Wg, nullptr, gup, st));
      sumpool2_add_kernel<<<grid1(a.numel()), 256, 0, st>>>(gup, a.H, a.W, a.C, a.numel(), a.g);
      chk(after_launch("sumpool2_add_kernel"));
    } else {
      chk(launch_conv_fwd(gy, a.N, Hg, Wg, w.Cout, 0, 1, w.ks, wd, nullptr, a.C, Hg, Wg, a.g, a.g, st));   // ga += conv(gy)
    }
  });
  return y;
}

// a = dropout(swish?(GroupNorm(x))); backward accumulates into x.g and writes the affine gradients
T4 gn_layer(Ctx& cx, const T4& x, int pw, int pb, bool swish, float drop_p) {
  hsidm_ctx* c = cx.c;
  TrainState* t = cx.t;
  const int groups = c->cfg.norm_groups, cpg = x.C / groups, HW = x.H * x.W;
  T4 a = cx.alloc(x.N, x.H, x.W, x.C);
  float* stats = cx.scratch((int64_t)2 * x.N * groups);
  float* part = cx.scratch((int64_t)2 * x.N * x.C);
  const float* gamma = c->ps.dev(pw);
  const float* beta = c->ps.dev(pb);
  const uint32_t layer = (uint32_t)t->dropout_layer++;
  const uint64_t seed = t->dropout_seed;
  cx.run([&] {
    gn_fwd_train_kernel<<<dim3(groups, x.N), 256, 0, cx.st()>>>(x.p, HW, x.C, groups, gamma, beta, swish ? 1 : 0, drop_p, seed, layer, stats, a.p);
    return after_launch("gn_fwd_train_kernel");
  });
  float* ggam = cx.dry() ? nullptr : cx.grad_of(pw);
  float* gbet = cx.dry() ? nullptr : cx.grad_of(pb);
  cx.back([=]() {
    cudaStream_t st = t->stream;
    auto chk = [&](int s) {
      if (s != HSIDM_OK && t->status == HSIDM_OK) t->status = s;
    };
    const int threads = 256;
    const int lanes = threads / cpg;
    gn_bwd_train_kernel<<<dim3(groups, x.N), threads, sizeof(float) * 2 * lanes * cpg, st>>>(x.p, a.g, HW, x.C, groups, gamma, beta, swish ? 1 : 0,
                                                                                           drop_p, seed, layer, stats, x.g, part);
    chk(after_launch("gn_bwd_train_kernel"));
    reduce_rows_kernel<<<(unsigned)ceil_div(x.C, 128), 128, 0, st>>>(part, x.N, x.C, 2, 0, ggam);
    chk(after_launch("reduce_rows_kernel"));
    reduce_rows_kernel<<<(unsigned)ceil_div(x.C, 128), 128, 0, st>>>(part, x.N, x.C, 2, 1, gbet);
    chk(after_launch("reduce_rows_kernel"));
  });
  return a;
}

T4 concat_layer(Ctx& cx, const T4& a, const T4& b) {
  T4 y = cx.alloc(a.N, a.H, a.W, a.C + b.C);
  const int64_t pixels = (int64_t)a.N * a.H * a.W;
  cx.run([&] {
    concat_kernel<<<grid1(y.numel()), 256, 0, cx.st()>>>(a.p, a.C, b.p, b.C, pixels, y.p);
    return after_launch("concat_kernel");
  });
  TrainState* t = cx.t;
  cx.back([=]() {
    split_add_kernel<<<grid1(y.numel()), 256, 0, t->stream>>>(y.g, a.g, a.C, b.g, b.C, pixels);
    const int s = after_launch("split_add_kernel");
    if (s != HSIDM_OK && t->status == HSIDM_OK) t->status = s;
  });
  return y;
}

// y += x (in place on y); backward: x.g += y.g
void add_into(Ctx& cx, T4& y, const T4& x) {
  cx.run([&] {
    add_kernel<<<grid1(y.numel()), 256, 0, cx.st()>>>(y.p, x.p, y.numel());
    return after_launch("add_kernel");
  });
  TrainState* t = cx.t;
  const T4 yy = y;
  cx.back([=]() {
    add_kernel<<<grid1(yy.numel()), 256, 0, t->stream>>>(x.g, yy.g, yy.numel());
    const int s = after_launch("add_kernel");
    if (s != HSIDM_OK && t->status == HSIDM_OK) t->status = s;
  });
}

struct NoiseEmb {
  float* pe = nullptr;     // [N][dim]
  float* h1 = nullptr;     // [N][4dim]
  float* a1 = nullptr;     // [N][4dim] swish(h1)
  float* temb = nullptr;   // [N][dim]
  float* gtemb = nullptr;  // [N][dim], accumulated by the layers
  int N = 0, dim = 0;
};

// h += Linear_l(temb)[n, :, None, None]  (FeatureWiseAffine, unet.py:45-49)
void noise_bias_layer(Ctx& cx, T4& h, const NoiseEmb& ne, int pw, int pb) {
  hsidm_ctx* c = cx.c;
  TrainState* t = cx.t;
  float* nb = cx.scratch((int64_t)ne.N * h.C);
  float* gnb = cx.scratch((int64_t)ne.N * h.C);
  const float* W = c->ps.dev(pw);
  const float* b = c->ps.dev(pb);
  const int HW = h.H * h.W;
  cx.run([&] {
    linear_fwd_kernel<<<(unsigned)ceil_div((int64_t)ne.N * h.C, 128), 128, 0, cx.st()>>>(ne.temb, ne.N, ne.dim, W, b, h.C, nb);
    HSIDM_TRY(after_launch("linear_fwd_kernel"));
    add_rowbias_kernel<<<grid1(h.numel()), 256, 0, cx.st()>>>(h.p, nb, HW, h.C, h.numel());
    return after_launch("add_rowbias_kernel");
  });
  float* gW = cx.dry() ? nullptr : cx.grad_of(pw);
  float* gb = cx.dry() ? nullptr : cx.grad_of(pb);
  const T4 hh = h;
  const NoiseEmb e = ne;
  cx.back([=]() {
    cudaStream_t st = t->stream;
    auto chk = [&](int s) {
      if (s != HSIDM_OK && t->status == HSIDM_OK) t->status = s;
    };
    colsum_kernel<<<dim3((unsigned)ceil_div(hh.C, 32), (unsigned)hh.N), 256, 0, st>>>(hh.g, HW, hh.C, gnb);
    chk(after_launch("colsum_kernel"));
    linear_bwd_w_kernel<<<(unsigned)ceil_div((int64_t)hh.C * (e.dim + 1), 128), 128, 0, st>>>(gnb, e.temb, e.N, e.dim, hh.C, gW, gb);
    chk(after_launch("linear_bwd_w_kernel"));
    linear_bwd_x_kernel<<<(unsigned)ceil_div((int64_t)e.N * e.dim, 128), 128, 0, st>>>(gnb, W, e.N, e.dim, hh.C, e.gtemb);
    chk(after_launch("linear_bwd_x_kernel"));
  });
}

// SelfAttention (unet.py:124-143), n_head = 1
T4 attention_layer(Ctx& cx, const ResW& r, const T4& x) {
  hsidm_ctx* c = cx.c;
  TrainState* t = cx.t;
  const int C = x.C, S = x.H * x.W, N = x.N;
  T4 nrm = gn_layer(cx, x, r.an_w, r.an_b, false, 0.f);
  T4 qkv = conv_layer(cx, nrm, r.qkv);                       // [N,S,3C]: q | k | v channel chunks
  float* P = cx.scratch((int64_t)N * S * S);
  float* dP = cx.scratch((int64_t)N * S * S);
  T4 av = cx.alloc(N, x.H, x.W, C);
  const float scale = 1.0f / std::sqrt((float)C);
  cx.run([&] {
    GemmOp qk;
    qk.A = qkv.p, qk.B = qkv.p + C, qk.C = P, qk.M = S, qk.N = S, qk.K = C, qk.lda = qk.ldb = 3 * C, qk.ldc = S;
    qk.sA = qk.sB = (int64_t)S * 3 * C, qk.sC = (int64_t)S * S, qk.batch = N, qk.transB = 1, qk.c_f32 = 1, qk.alpha = scale;
    HSIDM_TRY(gemm_simt(qk, HSIDM_F32, cx.st()));
    HSIDM_TRY(softmax_rows(P, (int64_t)N * S, S, cx.st()));
    GemmOp pv;
    pv.A = P, pv.B = qkv.p + 2 * C, pv.C = av.p, pv.M = S, pv.N = C, pv.K = S, pv.lda = S, pv.ldb = 3 * C, pv.ldc = C;
    pv.sA = (int64_t)S * S, pv.sB = (int64_t)S * 3 * C, pv.sC = (int64_t)S * C, pv.batch = N, pv.transB = 0, pv.a_f32 = 1, pv.c_f32 = 1;
    return gemm_simt(pv, HSIDM_F32, cx.st());
  });
  cx.back([=]() {
    cudaStream_t st = t->stream;
    auto chk = [&](int s) {
      if (s != HSIDM_OK && t->status == HSIDM_OK) t->status = s;
    };
    // dP = dO v^T
    GemmOp g1;
    g1.A = av.g, g1.B = qkv.p + 2 * C, g1.C = dP, g1.M = S, g1.N = S, g1.K = C, g1.lda = C, g1.ldb = 3 * C, g1.ldc = S;
    g1.sA = (int64_t)S * C, g1.sB = (int64_t)S * 3 * C, g1.sC = (int64_t)S * S, g1.batch = N, g1.transB = 1, g1.a_f32 = 1, g1.c_f32 = 1;
    chk(gemm_simt(g1, HSIDM_F32, st));
    // dv = P^T dO  -> qkv.g[..., 2C:3C]   (qkv.g is zero before: the conv above has a single consumer)
    TnP tv{P, av.g, qkv.g + 2 * C, S, S, C, S, C, 3 * C, (int64_t)S * S, (int64_t)S * C, (int64_t)S * 3 * C, 1.0f};
    gemm_tn_kernel<<<dim3((unsigned)ceil_div(C, 32), (unsigned)ceil_div(S, 32), (unsigned)N), 256, 0, st>>>(tv);
    chk(after_launch("gemm_tn_kernel"));
    // dS = P * (dP - rowsum(dP*P)) / sqrt(C)
    softmax_bwd_kernel<<<(unsigned)ceil_div((int64_t)N * S, 8), 256, 0, st>>>(P, dP, (int64_t)N * S, S, scale);
    chk(after_launch("softmax_bwd_kernel"));
    // dq = dS k -> qkv.g[..., 0:C]
    GemmOp g2;
    g2.A = dP, g2.B = qkv.p + C, g2.C = qkv.g, g2.M = S, g2.N = C, g2.K = S, g2.lda = S, g2.ldb = 3 * C, g2.ldc = 3 * C;
    g2.sA = (int64_t)S * S, g2.sB = (int64_t)S * 3 * C, g2.sC = (int64_t)S * 3 * C, g2.batch = N, g2.transB = 0, g2.a_f32 = 1, g2.c_f32 = 1;
    chk(gemm_simt(g2, HSIDM_F32, st));
    // dk = dS^T q -> qkv.g[..., C:2C]
    TnP tk{dP, qkv.p, qkv.g + C, S, S, C, S, 3 * C, 3 * C, (int64_t)S * S, (int64_t)S * 3 * C, (int64_t)S * 3 * C, 1.0f};
    gemm_tn_kernel<<<dim3((unsigned)ceil_div(C, 32), (unsigned)ceil_div(S, 32), (unsigned)N), 256, 0, st>>>(tk);
    chk(after_launch("gemm_tn_kernel"));
  });
  T4 out = conv_layer(cx, av, r.aout);
  add_into(cx, out, x);   // residual is the un-normalised input (unet.py:143)
  (void)c;
  return out;
}

// ResnetBlocWithAttn (unet.py:94-111, 146-159)
T4 res_layer(Ctx& cx, const ResW& r, const T4& x, const NoiseEmb& ne, float drop_p) {
  T4 a1 = gn_layer(cx, x, r.gn1_w, r.gn1_b, true, 0.f);       // block1 has no dropout (unet.py:101)
  T4 h = conv_layer(cx, a1, r.c1);
  noise_bias_layer(cx, h, ne, r.nf_w, r.nf_b);
  T4 a2 = gn_layer(cx, h, r.gn2_w, r.gn2_b, true, drop_p);    // block2 = GN -> Swish -> Dropout -> conv (unet.py:102)
  T4 y = conv_layer(cx, a2, r.c2);
  if (r.has_res) {
    T4 sc = conv_layer(cx, x, r.rc);
    add_into(cx, y, sc);
  } else {
    add_into(cx, y, x);
  }
  return r.attn ? attention_layer(cx, r, y) : y;
}

// the whole network; returns eps (NHWC)
T4 unet_train_forward(Ctx& cx, const float* hr, const float* sr, const float* noise, const float* level_dev, int N, int H, int W) {
  hsidm_ctx* c = cx.c;
  TrainState* t = cx.t;
  const int ch = c->cfg.out_channel, dim = c->cfg.inner_channel;
  const float drop_p = c->cfg.dropout;
  // noise-level embedding
  NoiseEmb ne;
  ne.N = N, ne.dim = dim;
  ne.pe = cx.scratch((int64_t)N * dim), ne.h1 = cx.scratch((int64_t)N * 4 * dim), ne.a1 = cx.scratch((int64_t)N * 4 * dim);
  ne.temb = cx.scratch((int64_t)N * dim);
  {
    const int64_t off = t->grad.take((int64_t)N * dim * 4);
    ne.gtemb = cx.dry() ? nullptr : reinterpret_cast<float*>(t->grad.base + off);   // in the gradient arena: zeroed per step
  }
  float* ga1 = cx.scratch((int64_t)N * 4 * dim);
  const float *w1 = c->ps.dev(c->mlp1_w), *b1 = c->ps.dev(c->mlp1_b), *w3 = c->ps.dev(c->mlp3_w), *b3 = c->ps.dev(c->mlp3_b);
  cx.run([&] {
    noise_mlp_fwd_kernel<<<N, 128, sizeof(float) * 5 * dim, cx.st()>>>(level_dev, dim, w1, b1, w3, b3, ne.pe, ne.h1, ne.a1, ne.temb);
    return after_launch("noise_mlp_fwd_kernel");
  });
  {
    float *gw1 = cx.dry() ? nullptr : cx.grad_of(c->mlp1_w), *gb1 = cx.dry() ? nullptr : cx.grad_of(c->mlp1_b);
    float *gw3 = cx.dry() ? nullptr : cx.grad_of(c->mlp3_w), *gb3 = cx.dry() ? nullptr : cx.grad_of(c->mlp3_b);
    const NoiseEmb e = ne;
    cx.back([=]() {   // recorded first, so it runs last in the backward: every layer has accumulated into gtemb by then
      cudaStream_t st = t->stream;
      auto chk = [&](int s) {
        if (s != HSIDM_OK && t->status == HSIDM_OK) t->status = s;
      };
      const int d4 = 4 * dim;
      linear_bwd_w_kernel<<<(unsigned)ceil_div((int64_t)dim * (d4 + 1), 128), 128, 0, st>>>(e.gtemb, e.a1, N, d4, dim, gw3, gb3);
      chk(after_launch("linear_bwd_w_kernel"));
      cudaMemsetAsync(ga1, 0, sizeof(float) * (int64_t)N * d4, st);
      linear_bwd_x_kernel<<<(unsigned)ceil_div((int64_t)N * d4, 128), 128, 0, st>>>(e.gtemb, w3, N, d4, dim, ga1);
      chk(after_launch("linear_bwd_x_kernel"));
      swish_bwd_kernel<<<grid1((int64_t)N * d4), 256, 0, st>>>(e.h1, ga1, (int64_t)N * d4);
      chk(after_launch("swish_bwd_kernel"));
      linear_bwd_w_kernel<<<(unsigned)ceil_div((int64_t)d4 * (dim + 1), 128), 128, 0, st>>>(ga1, e.pe, N, dim, d4, gw1, gb1);
      chk(after_launch("linear_bwd_w_kernel"));
    });
  }
  // input: cat([SR, x_noisy]) as NHWC
  T4 xin = cx.alloc(N, H, W, 2 * ch);
  xin.g = nullptr;   // no gradient needed for the network input
  cx.run([&] {
    qsample_cat_kernel<<<grid1(xin.numel()), 256, 0, cx.st()>>>(hr, sr, noise, level_dev, ch, H * W, xin.numel(), xin.p);
    return after_launch("qsample_cat_kernel");
  });
  std::vector<T4> feats;
  T4 x;
  for (const LayerW& L : c->downs) {
    if (L.kind == LayerW::CONV) x = conv_layer(cx, xin, L.conv);
    else if (L.kind == LayerW::RES) x = res_layer(cx, L.rb, x, ne, drop_p);
    else x = conv_layer(cx, x, L.conv, /*stride=*/2);
    feats.push_back(x);
  }
  for (const LayerW& L : c->mid) x = res_layer(cx, L.rb, x, ne, drop_p);
  for (const LayerW& L : c->ups) {
    if (L.kind == LayerW::RES) {
      T4 cat = concat_layer(cx, x, feats.back());
      feats.pop_back();
      x = res_layer(cx, L.rb, cat, ne, drop_p);
    } else {
      x = conv_layer(cx, x, L.conv, 1, /*up=*/1);
    }
  }
  T4 a = gn_layer(cx, x, c->fin_gn_w, c->fin_gn_b, true, 0.f);   // final_conv is a Block without dropout (unet.py:236)
  return conv_layer(cx, a, c->fin_conv);
}

int reserve(Bump& b, int64_t bytes) {
  if (bytes <= b.cap) return HSIDM_OK;
  if (b.base) {
    HSIDM_CUDA(cudaDeviceSynchronize());
    cudaFree(b.base);
    b.base = nullptr, b.cap = 0;
  }
  bytes = round_up(bytes, 1 << 20);
  cudaError_t e = cudaMalloc(&b.base, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    HSIDM_FAIL(HSIDM_OOM_WORKSPACE, "training workspace of %lld bytes could not be allocated: %s", (long long)bytes, cudaGetErrorString(e));
  }
  b.cap = bytes;
  return HSIDM_OK;
}

}  // namespace
}  // namespace hsidm

using namespace hsidm;

extern "C" {

int64_t hsidm_train_grad_numel(const hsidm_ctx* c) { return c ? c->ps.bytes() / (int64_t)sizeof(float) : 0; }

int64_t hsidm_train_grad_offset(const hsidm_ctx* c, int index) {
  if (!c || index < 0 || index >= c->ps.size()) return -1;
  return c->ps.dev(index) - c->ps.dev(0);
}

int hsidm_train_forward(hsidm_ctx* c, const float* hr, const float* sr, const float* noise, const float* levels, int B, int H, int W,
                        int loss_type, uint64_t dropout_seed, float* grad_slab, float* loss_sum, hsidm_stream stream_) {
  if (!c || !hr || !sr || !noise || !levels || !loss_sum || !grad_slab) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_train_forward: null argument");
  if (loss_type != 0 && loss_type != 1) HSIDM_FAIL(HSIDM_BAD_ARG, "loss_type must be 0 (l1) or 1 (l2)");
  if (!c->committed) HSIDM_FAIL(HSIDM_BAD_STATE, "hsidm_unet_commit has not been called");
  if (c->cfg.in_channel != 2 * c->cfg.out_channel)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "the conditional training step needs in_channel == 2*out_channel (diffusion.py:246-247)");
  const int div = 1 << (c->cfg.n_mults - 1);
  if (B <= 0 || H <= 0 || W <= 0 || H % div || W % div) HSIDM_FAIL(HSIDM_BAD_SHAPE, "bad training batch shape %dx%dx%d", B, H, W);
  if (c->cfg.dropout < 0.f || c->cfg.dropout >= 1.f) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "dropout %f outside [0,1)", c->cfg.dropout);
  {
    int max_mult = 1;
    for (int i = 0; i < c->cfg.n_mults; ++i) max_mult = std::max(max_mult, c->cfg.channel_mults[i]);
    if (2 * c->cfg.inner_channel * max_mult / c->cfg.norm_groups > 256)
      HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "training: more than 256 channels per GroupNorm group");
  }
  HSIDM_DEVICE(c->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!c->train) {
    c->train = new TrainState();
    c->train_free = free_train;
  }
  TrainState* t = static_cast<TrainState*>(c->train);
  Ctx cx{c, t};
  t->stream = stream, t->status = HSIDM_OK, t->have_forward = false;
  t->tape.clear();
  t->grads = grad_slab;
  if (t->N != B || t->H != H || t->W != W || !t->act.base) {
    // dry pass: size the two arenas
    t->act.dry = t->grad.dry = true;
    t->act.top = t->grad.top = t->act.peak = t->grad.peak = 0;
    t->dropout_layer = 0;
    cx.scratch((int64_t)B);
    const T4 e = unet_train_forward(cx, nullptr, nullptr, nullptr, nullptr, B, H, W);
    cx.scratch(e.numel());   // the loss gradient
    t->act.dry = t->grad.dry = false;
    HSIDM_TRY(reserve(t->act, t->act.peak));
    HSIDM_TRY(reserve(t->grad, t->grad.peak));
    t->N = B, t->H = H, t->W = W;
    if (!t->loss_parts) {
      t->loss_blocks = 1024;
      HSIDM_CUDA(cudaMalloc(&t->loss_parts, sizeof(float) * t->loss_blocks));
      HSIDM_CUDA(cudaMalloc(&t->loss_sum_dev, sizeof(float)));
    }
  }
  HSIDM_CUDA(cudaStreamWaitEvent(stream, c->ev_arena, 0));
  t->act.top = t->grad.top = 0;
  t->dropout_layer = 0;
  t->dropout_seed = dropout_seed;
  t->loss_type = loss_type;
  t->noise = noise;
  t->inv_count = 1.0f / ((float)B * c->cfg.out_channel * H * W);
  float* level_dev = cx.scratch((int64_t)B);
  HSIDM_CUDA(cudaMemcpyAsync(level_dev, levels, sizeof(float) * B, cudaMemcpyDefault, stream));   // host or device levels
  t->eps = unet_train_forward(cx, hr, sr, noise, level_dev, B, H, W);
  HSIDM_TRY(t->status);
  // loss and its gradient w.r.t. eps (the gradient arena is zeroed by the backward before anything accumulates, so the
  // eps gradient is produced there, not here); here only the value
  const int64_t total = t->eps.numel();
  const int blocks = (int)std::min<int64_t>(t->loss_blocks, ceil_div(total, 256));
  // geps goes to a scratch region of the ACTIVATION arena and is copied in by the backward
  float* geps = cx.scratch(total);
  if (t->act.top > t->act.cap || t->grad.top > t->grad.cap)
    HSIDM_FAIL(HSIDM_OOM_WORKSPACE, "training workspace overrun (%lld of %lld bytes): dry pass and real pass disagree", (long long)t->act.top,
               (long long)t->act.cap);
  loss_kernel<<<blocks, 256, 0, stream>>>(t->eps.p, noise, c->cfg.out_channel, H * W, total, loss_type, t->inv_count, t->loss_parts, geps);
  HSIDM_TRY(after_launch("loss_kernel"));
  loss_fold_kernel<<<1, 32, 0, stream>>>(t->loss_parts, blocks, t->loss_sum_dev);
  HSIDM_TRY(after_launch("loss_fold_kernel"));
  HSIDM_CUDA(cudaMemcpyAsync(loss_sum, t->loss_sum_dev, sizeof(float), cudaMemcpyDefault, stream));
  {
    const T4 eps = t->eps;
    TrainState* tt = t;
    t->tape.emplace_back([=]() {   // recorded last: the first thing the backward does after zeroing the gradient arena
      if (cudaMemcpyAsync(eps.g, geps, sizeof(float) * total, cudaMemcpyDeviceToDevice, tt->stream) != cudaSuccess) {
        set_last_error("copy of the loss gradient failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (tt->status == HSIDM_OK) tt->status = HSIDM_CUDA_ERROR;
      }
    });
  }
  t->have_forward = true;
  HSIDM_CUDA(cudaEventRecord(c->ev_arena, stream));
  return HSIDM_OK;
}

int hsidm_train_backward(hsidm_ctx* c, float upstream, hsidm_stream stream_) {
  if (!c || !c->train) HSIDM_FAIL(HSIDM_BAD_STATE, "hsidm_train_backward: no forward pass to differentiate");
  TrainState* t = static_cast<TrainState*>(c->train);
  if (!t->have_forward) HSIDM_FAIL(HSIDM_BAD_STATE, "hsidm_train_backward: the saved forward pass was already consumed");
  HSIDM_DEVICE(c->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  t->stream = stream, t->status = HSIDM_OK;
  HSIDM_CUDA(cudaStreamWaitEvent(stream, c->ev_arena, 0));
  HSIDM_CUDA(cudaMemsetAsync(t->grad.base, 0, (size_t)t->grad.top, stream));
  HSIDM_CUDA(cudaMemsetAsync(t->grads, 0, (size_t)c->ps.bytes(), stream));
  for (auto it = t->tape.rbegin(); it != t->tape.rend(); ++it) {
    (*it)();
    if (t->status != HSIDM_OK) break;
  }
  t->have_forward = false;
  t->tape.clear();
  HSIDM_TRY(t->status);
  if (upstream != 1.0f) {   // d(loss_normalised * upstream): the saved gradient is for upstream = 1
    const int64_t n = c->ps.bytes() / (int64_t)sizeof(float);
    scale_kernel<<<grid1(n), 256, 0, stream>>>(t->grads, upstream, n);
    HSIDM_TRY(after_launch("scale_kernel"));
  }
  HSIDM_CUDA(cudaEventRecord(c->ev_arena, stream));
  return HSIDM_OK;
}

}  // extern "C"
