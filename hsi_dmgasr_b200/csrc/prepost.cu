// The steps either side of the sampling path, on the GPU:
//   * bicubic x`scale` pre-upsampling of the low-resolution cube (reference: torch.nn.functional.interpolate(...,
//     scale_factor=4, mode='bicubic'), sr_gae.py:72) - PyTorch's convention: align_corners=False, A = -0.75, source
//     index (o + 0.5)/scale - 0.5 without clamping, neighbour indices clamped to the border;
//   * the two validation metrics the parity gates are stated in, per cube, after the driver's clamp to [0,1]
//     (sr_gae.py:474-475): MPSNR (eval_hsi.py:110-121) and SAM (eval_hsi.py:47-65).
// Both are HBM-bound: 4 B read per low-res element + 4 B written per output element; 8 B read per element pair.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "kernels.cuh"

namespace hsidm {
namespace {

__device__ __forceinline__ void cubic_weights(float t, float (&w)[4]) {
  constexpr float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// One thread per output pixel of one (n, c) plane; planes along blockIdx.y.
__global__ void __launch_bounds__(256) bicubic_kernel(const float* __restrict__ src, float* __restrict__ dst, int h, int w, int H, int W,
                                                      float inv_scale_y, float inv_scale_x, int clamp01) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= H * W) return;
  const int oy = o / W, ox = o - oy * W;
  const float* plane = src + (long long)blockIdx.y * h * w;
  const float ry = inv_scale_y * (oy + 0.5f) - 0.5f, rx = inv_scale_x * (ox + 0.5f) - 0.5f;
  const float fy = floorf(ry), fx = floorf(rx);
  const int iy = (int)fy, ix = (int)fx;
  float wy[4], wx[4];
  cubic_weights(ry - fy, wy);
  cubic_weights(rx - fx, wx);
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = min(max(iy - 1 + j, 0), h - 1);
    float row = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = min(max(ix - 1 + i, 0), w - 1);
      row = fmaf(__ldg(plane + (long long)y * w + x), wx[i], row);
    }
    acc = fmaf(row, wy[j], acc);
  }
  if (clamp01) acc = fminf(fmaxf(acc, 0.f), 1.f);
  dst[(long long)blockIdx.y * H * W + o] = acc;
}

// Per-cube partial sums: grid = (slabs, N).  Every block covers a pixel range of one cube and all its bands:
//   sq[c]   += (clamp(a) - clamp(b))^2 per band (MPSNR),
//   angle   += arccos(<p,t> / (|p||t|)), count += 1 over pixels whose two spectra are non-zero (SAM).
// Partials are written per block and folded in fixed order by metrics_finalize (no floating-point atomics).
__global__ void __launch_bounds__(256) metrics_partial_kernel(const float* __restrict__ truth, const float* __restrict__ pred, int C, int HW,
                                                              int pix_per_block, double* __restrict__ part /*[N][slabs][C+2]*/) {
  extern __shared__ double sh[];   // [C + 2]
  const int n = blockIdx.y, slab = blockIdx.x, slabs = gridDim.x;
  for (int i = threadIdx.x; i < C + 2; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const float* t0 = truth + (long long)n * C * HW;
  const float* p0 = pred + (long long)n * C * HW;
  const int px0 = slab * pix_per_block, px1 = min(HW, px0 + pix_per_block);
  // SAM: one thread per pixel walks the bands (coalesced across the warp: NCHW planes)
  double ang = 0.0, cnt = 0.0;
  for (int px = px0 + threadIdx.x; px < px1; px += blockDim.x) {
    float dot = 0.f, nt = 0.f, np = 0.f;
    for (int c = 0; c < C; ++c) {
      const float a = fminf(fmaxf(__ldg(t0 + (long long)c * HW + px), 0.f), 1.f);
      const float b = fminf(fmaxf(__ldg(p0 + (long long)c * HW + px), 0.f), 1.f);
      dot = fmaf(a, b, dot), nt = fmaf(a, a, nt), np = fmaf(b, b, np);
    }
    if (nt != 0.f && np != 0.f) {
      const float cosv = dot / (sqrtf(nt) * sqrtf(np));
      ang += (double)acosf(fminf(fmaxf(cosv, -1.f), 1.f));
      cnt += 1.0;
    }
  }
  // MPSNR: warp w takes bands w, w + warps, ...; lanes stride the pixel range
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  for (int c = warp; c < C; c += warps) {
    double s = 0.0;
    for (int px = px0 + lane; px < px1; px += 32) {
      const float a = fminf(fmaxf(__ldg(t0 + (long long)c * HW + px), 0.f), 1.f);
      const float b = fminf(fmaxf(__ldg(p0 + (long long)c * HW + px), 0.f), 1.f);
      const double d = (double)a - (double)b;
      s += d * d;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sh[c] = s;
  }
  // block fold of the SAM terms in lane order
  for (int o = 16; o > 0; o >>= 1) ang += __shfl_xor_sync(0xffffffffu, ang, o), cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  __shared__ double wa[8], wc[8];
  if (lane == 0) wa[warp] = ang, wc[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < warps; ++i) a += wa[i], c += wc[i];
    sh[C] = a, sh[C + 1] = c;
  }
  __syncthreads();
  double* out = part + ((long long)n * slabs + slab) * (C + 2);
  for (int i = threadIdx.x; i < C + 2; i += blockDim.x) out[i] = sh[i];
}

__global__ void metrics_finalize_kernel(const double* __restrict__ part, int C, int HW, int slabs, float data_range,
                                        float* __restrict__ out /*[N][2] = (mpsnr dB, sam degrees)*/) {
  const int n = blockIdx.x;
  // bands in index order by one thread each, folded by thread 0 afterwards through shared memory
  extern __shared__ double band[];   // [C]
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0;
    for (int sl = 0; sl < slabs; ++sl) s += part[((long long)n * slabs + sl) * (C + 2) + c];
    band[c] = 10.0 * log10((double)data_range * data_range / (s / HW));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0, a = 0.0, k = 0.0;
    for (int c = 0; c < C; ++c) m += band[c];
    for (int sl = 0; sl < slabs; ++sl) {
      a += part[((long long)n * slabs + sl) * (C + 2) + C];
      k += part[((long long)n * slabs + sl) * (C + 2) + C + 1];
    }
    out[2 * n] = (float)(m / C);
    out[2 * n + 1] = (float)(a / k * 180.0 / 3.14159265358979323846);
  }
}


// Feathered overlap-add, gather form: one thread per scene pixel, all bands.  ramp(d) = (d+1)/(ov+1) over the first / last
// `ov` pixels of a tile, 1 inside; weight = ramp_y * ramp_x (fp32 product, as pipeline.feather_window builds it); each band
// accumulates tile * weight in ascending (iy, ix) order and divides by the weight sum: the same operations in the same
// order as the host reference in tests/, so the result is bit-identical to it.
__global__ void __launch_bounds__(256) blend_kernel(const float* __restrict__ tiles, const int* __restrict__ ys, int ny,
                                                    const int* __restrict__ xs, int nx, int C, int t, int ov, int H, int W,
                                                    float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  auto ramp = [&](int d) {
    if (d < ov) return __fdiv_rn((float)d + 1.0f, (float)ov + 1.0f);
    if (d >= t - ov) return __fdiv_rn((float)(t - 1 - d) + 1.0f, (float)ov + 1.0f);
    return 1.0f;
  };
  // covering tiles along each axis (at most a handful)
  int iy[8], ix[8], cy = 0, cx = 0;
  for (int i = 0; i < ny && cy < 8; ++i)
    if (y >= ys[i] && y < ys[i] + t) iy[cy++] = i;
  for (int i = 0; i < nx && cx < 8; ++i)
    if (x >= xs[i] && x < xs[i] + t) ix[cx++] = i;
  float wsum = 0.f;
  for (int a = 0; a < cy; ++a)
    for (int b = 0; b < cx; ++b) wsum = __fadd_rn(wsum, __fmul_rn(ramp(y - ys[iy[a]]), ramp(x - xs[ix[b]])));
  const long long tt = (long long)t * t;
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
    for (int a = 0; a < cy; ++a)
      for (int b = 0; b < cx; ++b) {
        const int dy = y - ys[iy[a]], dx = x - xs[ix[b]];
        const float w = __fmul_rn(ramp(dy), ramp(dx));
        const float v = __ldg(tiles + (((long long)iy[a] * nx + ix[b]) * C + c) * tt + (long long)dy * t + dx);
        acc = __fadd_rn(acc, __fmul_rn(v, w));
      }
    out[((long long)c * H + y) * W + x] = __fdiv_rn(acc, wsum);
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// MATLAB-style imresize (imsize.py:116-158; the dataset code's degradation and pre-upsampling, HStest.py:44-45,
// HStrain.py:61-63): separable resampling with an antialiasing kernel when shrinking.  Per output index o of one axis
// (imsize.py:35-59, `contributions`): u = (o+1)/scale + 0.5 (1 - 1/scale), P = ceil(width) + 2 taps starting at
// floor(u - width/2) - 1 (0-based), weight h(u - tap - 1) with h = kernel (scale >= 1) or scale*kernel(scale*.)
// (scale < 1, width = 4/scale), normalised to sum 1; taps outside the axis are mirrored (period 2*length).
// All of it in float64 like the reference; the result is rounded to fp32 once.
__device__ __forceinline__ double imresize_kernel_fn(double x, int method) {
  const double ax = fabs(x);
  if (method == 1) return ax <= 1.0 ? 1.0 - ax : 0.0;   // triangle (imsize.py:18-23): (x+1)[-1<=x<0] + (1-x)[0<=x<=1]
  const double ax2 = ax * ax, ax3 = ax2 * ax;             // cubic, a = -0.5 (imsize.py:26-32)
  if (ax <= 1.0) return 1.5 * ax3 - 2.5 * ax2 + 1.0;
  if (ax <= 2.0) return -0.5 * ax3 + 2.5 * ax2 - 4.0 * ax + 2.0;
  return 0.0;
}

__global__ void __launch_bounds__(128) imresize_table_kernel(double* __restrict__ wt, int* __restrict__ idx, int in_len, int out_len,
                                                             double scale, int P, int method) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= out_len) return;
  const double width = scale < 1.0 ? 4.0 / scale : 4.0;
  const double u = (double)(o + 1) / scale + 0.5 * (1.0 - 1.0 / scale);
  const double left = floor(u - width / 2.0);
  double sum = 0.0;
  for (int p = 0; p < P; ++p) {
    const double ind = left + (double)p - 1.0;
    const double arg = u - ind - 1.0;
    const double w = scale < 1.0 ? scale * imresize_kernel_fn(scale * arg, method) : imresize_kernel_fn(arg, method);
    wt[(long long)o * P + p] = w;
    sum += w;
    long long m = (long long)ind % (2LL * in_len);
    if (m < 0) m += 2LL * in_len;
    idx[(long long)o * P + p] = (int)(m < in_len ? m : 2LL * in_len - 1 - m);
  }
  for (int p = 0; p < P; ++p) wt[(long long)o * P + p] /= sum;
}

// One thread per output pixel of one plane; planes along blockIdx.y.  Rows first, then columns - in float64 the order of
// the two axes (the reference takes the axis with the smaller scale first, imsize.py:141-152) changes the last bit of a
// double at most, far below the fp32 rounding of the result.
__global__ void __launch_bounds__(256) imresize_kernel(const float* __restrict__ src, float* __restrict__ dst, int h, int w, int H, int W,
                                                       const double* __restrict__ wy, const int* __restrict__ iy, int Py,
                                                       const double* __restrict__ wx, const int* __restrict__ ix, int Px) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= H * W) return;
  const int oy = o / W, ox = o - oy * W;
  const float* plane = src + (long long)blockIdx.y * h * w;
  const double* wyr = wy + (long long)oy * Py;
  const int* iyr = iy + (long long)oy * Py;
  const double* wxr = wx + (long long)ox * Px;
  const int* ixr = ix + (long long)ox * Px;
  double acc = 0.0;
  for (int j = 0; j < Py; ++j) {
    const double wj = wyr[j];
    if (wj == 0.0) continue;
    const float* row = plane + (long long)iyr[j] * w;
    double r = 0.0;
    for (int i = 0; i < Px; ++i) r += wxr[i] * (double)__ldg(row + ixr[i]);
    acc += wj * r;
  }
  dst[(long long)blockIdx.y * H * W + o] = (float)acc;
}

// ------------------------------------------------------------------------------------------------------------------------
// The remaining indices of quality_assessment (eval_hsi.py:217-238) per cube, after the driver's clamp to [0,1]:
//   ERGAS (eval_hsi.py:18-35), CrossCorrelation (:58-70), RMSE (:88-96) from six per-band sums,
//   MSSIM (:124-135) = skimage.metrics.structural_similarity per band with its defaults for float images: 7x7 uniform
//   window, K1 = 0.01, K2 = 0.03, sample covariance (49/48), mean over the pixels whose window lies inside the image.
// Float64 accumulation, fixed-order folds: deterministic.
__device__ __forceinline__ float clamp01f(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__device__ __forceinline__ double block_sum_256(double v, double* sh /*[8]*/, int tid) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = tid >> 5, lane = tid & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < 8; ++i) s += sh[i];
  return s;
}

// grid (C, N): sums over one band plane of a, b, a^2, b^2, a*b, (a-b)^2 -> sums[n][c][6]
__global__ void __launch_bounds__(256) band_sums_kernel(const float* __restrict__ truth, const float* __restrict__ pred, int C, int HW,
                                                        double* __restrict__ sums) {
  __shared__ double sh[8];
  const long long plane = (long long)blockIdx.y * C + blockIdx.x;
  const float* t0 = truth + plane * HW;
  const float* p0 = pred + plane * HW;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int px = threadIdx.x; px < HW; px += 256) {
    const double a = clamp01f(__ldg(t0 + px)), b = clamp01f(__ldg(p0 + px)), d = a - b;
    s[0] += a, s[1] += b, s[2] += a * a, s[3] += b * b, s[4] += a * b, s[5] += d * d;
  }
  for (int k = 0; k < 6; ++k) {
    const double v = block_sum_256(s[k], sh, threadIdx.x);
    if (threadIdx.x == 0) sums[plane * 6 + k] = v;
  }
}

// grid (tiles_x * tiles_y, C * N), block 32 x 8: every thread evaluates SSIM at one pixel whose 7x7 window is inside the
// image; the block's sum goes to part[plane][tile].
constexpr int kSsimTX = 32, kSsimTY = 8, kSsimWin = 7;
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ truth, const float* __restrict__ pred, int H, int W, int tiles_x,
                                                   float data_range, double* __restrict__ part) {
  __shared__ float sa[kSsimTY + kSsimWin - 1][kSsimTX + kSsimWin - 1];
  __shared__ float sb[kSsimTY + kSsimWin - 1][kSsimTX + kSsimWin - 1];
  __shared__ double sh[8];
  const long long plane = blockIdx.y;
  const int tile = blockIdx.x, ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int y0 = ty * kSsimTY, x0 = tx * kSsimTX;   // top-left corner of the tile's first window
  const float* t0 = truth + plane * H * W;
  const float* p0 = pred + plane * H * W;
  const int tid = threadIdx.y * kSsimTX + threadIdx.x;
  constexpr int SW = kSsimTX + kSsimWin - 1, SH = kSsimTY + kSsimWin - 1;
  for (int i = tid; i < SW * SH; i += 256) {
    const int yy = i / SW, xx = i - yy * SW;
    const int y = y0 + yy, x = x0 + xx;
    const bool in = y < H && x < W;
    sa[yy][xx] = in ? clamp01f(__ldg(t0 + (long long)y * W + x)) : 0.f;
    sb[yy][xx] = in ? clamp01f(__ldg(p0 + (long long)y * W + x)) : 0.f;
  }
  __syncthreads();
  double S = 0.0;
  if (y0 + (int)threadIdx.y + kSsimWin <= H && x0 + (int)threadIdx.x + kSsimWin <= W) {
    double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
    for (int j = 0; j < kSsimWin; ++j)
#pragma unroll
      for (int i = 0; i < kSsimWin; ++i) {
        const double a = sa[threadIdx.y + j][threadIdx.x + i], b = sb[threadIdx.y + j][threadIdx.x + i];
        sx += a, sy += b, sxx += a * a, syy += b * b, sxy += a * b;
      }
    constexpr double NP = kSsimWin * kSsimWin, cov_norm = NP / (NP - 1.0);
    const double ux = sx / NP, uy = sy / NP;
    const double vx = cov_norm * (sxx / NP - ux * ux), vy = cov_norm * (syy / NP - uy * uy), vxy = cov_norm * (sxy / NP - ux * uy);
    const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
    S = ((2.0 * ux * uy + c1) * (2.0 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
  }
  const double v = block_sum_256(S, sh, tid);
  if (tid == 0) part[plane * gridDim.x + tile] = v;
}

// One block per cube: out[n] = (MPSNR, MSSIM, ERGAS, SAM, CrossCorrelation, RMSE) - the key order of quality_assessment.
// mpsnr_sam [N][2] comes from quality_metrics above.
__global__ void __launch_bounds__(128) assess_finalize_kernel(const double* __restrict__ sums, const double* __restrict__ ssim_part,
                                                              int ssim_tiles, const float* __restrict__ mpsnr_sam, int C, int H, int W,
                                                              float ratio, float* __restrict__ out) {
  __shared__ double acc[4][128];
  const int n = blockIdx.x;
  const double HW = (double)H * W;
  double ergas = 0, cc = 0, sq = 0, ssim = 0;
  for (int c = threadIdx.x; c < C; c += 128) {
    const double* s = sums + ((long long)n * C + c) * 6;
    const double ma = s[0] / HW, mb = s[1] / HW;
    ergas += (s[5] / HW) / (ma * ma);
    cc += (s[4] - HW * ma * mb) / sqrt((s[2] - HW * ma * ma) * (s[3] - HW * mb * mb));
    sq += s[5];
    double t = 0.0;
    if (ssim_part)
      for (int k = 0; k < ssim_tiles; ++k) t += ssim_part[((long long)n * C + c) * ssim_tiles + k];
    ssim += t / ((double)(H - kSsimWin + 1) * (W - kSsimWin + 1));
  }
  acc[0][threadIdx.x] = ergas, acc[1][threadIdx.x] = cc, acc[2][threadIdx.x] = sq, acc[3][threadIdx.x] = ssim;
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0, k = 0, q = 0, m = 0;
    for (int i = 0; i < 128; ++i) e += acc[0][i], k += acc[1][i], q += acc[2][i], m += acc[3][i];
    float* o = out + 6 * n;
    o[0] = mpsnr_sam[2 * n];
    o[1] = ssim_part ? (float)(m / C) : nanf("");
    o[2] = (float)((100.0 / ratio) * sqrt(e / C));
    o[3] = mpsnr_sam[2 * n + 1];
    o[4] = (float)(k / C);
    o[5] = (float)sqrt(q / (HW * C));
  }
}
}  // namespace

int bicubic_upsample(const float* src, float* dst, int planes, int h, int w, int scale, int clamp01, cudaStream_t stream) {
  if (planes <= 0 || h <= 0 || w <= 0 || scale < 1) HSIDM_FAIL(HSIDM_BAD_SHAPE, "bicubic_upsample: bad shape %d x %dx%d, scale %d", planes, h, w, scale);
  if (planes > 65535) HSIDM_FAIL(HSIDM_BAD_SHAPE, "bicubic_upsample: at most 65535 (image, band) planes per call (got %d)", planes);
  const int H = h * scale, W = w * scale;
  ProfScope prof(PROF_OTHER, 4.0 * planes * ((double)h * w + (double)H * W), stream, "bicubic");
  dim3 grid((unsigned)ceil_div((int64_t)H * W, 256), (unsigned)planes);
  bicubic_kernel<<<grid, 256, 0, stream>>>(src, dst, h, w, H, W, 1.0f / scale, 1.0f / scale, clamp01);
  return after_launch("bicubic_kernel");
}

int64_t quality_metrics_scratch_bytes(int N, int C, int HW) {
  const int slabs = (int)std::min<int64_t>(ceil_div(HW, 1024), 64);
  return (int64_t)N * slabs * (C + 2) * (int64_t)sizeof(double);
}

int quality_metrics(const float* truth, const float* pred, int N, int C, int HW, float data_range, void* scratch, float* out,
                    cudaStream_t stream) {
  if (N <= 0 || C <= 0 || HW <= 0 || C > 4096) HSIDM_FAIL(HSIDM_BAD_SHAPE, "quality_metrics: bad shape N=%d C=%d HW=%d", N, C, HW);
  const int slabs = (int)std::min<int64_t>(ceil_div(HW, 1024), 64);
  const int ppb = (int)ceil_div(HW, slabs);
  ProfScope prof(PROF_OTHER, 8.0 * N * C * (double)HW, stream, "metrics");
  metrics_partial_kernel<<<dim3(slabs, N), 256, sizeof(double) * (C + 2), stream>>>(truth, pred, C, HW, ppb, static_cast<double*>(scratch));
  HSIDM_TRY(after_launch("metrics_partial_kernel"));
  metrics_finalize_kernel<<<N, 128, sizeof(double) * C, stream>>>(static_cast<const double*>(scratch), C, HW, slabs, data_range, out);
  return after_launch("metrics_finalize_kernel");
}


static int imresize_taps(double scale) { return (int)std::ceil(scale < 1.0 ? 4.0 / scale : 4.0) + 2; }

int64_t imresize_scratch_bytes(int H, int W, double scale_h, double scale_w) {
  // per axis: out_len x P doubles + out_len x P ints, each table rounded up to 16 bytes
  auto axis = [](int out_len, double scale) {
    const int64_t n = (int64_t)out_len * imresize_taps(scale);
    return ((n * 8 + 15) / 16 + (n * 4 + 15) / 16) * 16;
  };
  return axis(H, scale_h) + axis(W, scale_w);
}

int imresize(const float* src, float* dst, int planes, int h, int w, int H, int W, double scale_h, double scale_w, int method,
             void* scratch, cudaStream_t stream) {
  if (planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) HSIDM_FAIL(HSIDM_BAD_SHAPE, "imresize: bad shape %d x %dx%d -> %dx%d", planes, h, w, H, W);
  if (planes > 65535) HSIDM_FAIL(HSIDM_BAD_SHAPE, "imresize: at most 65535 planes per call (got %d)", planes);
  if (method != 0 && method != 1) HSIDM_FAIL(HSIDM_BAD_ARG, "imresize: method %d (0 = bicubic, 1 = bilinear)", method);
  if (!(scale_h > 0.0) || !(scale_w > 0.0) || imresize_taps(scale_h) > 4096 || imresize_taps(scale_w) > 4096)
    HSIDM_FAIL(HSIDM_BAD_ARG, "imresize: bad scale %g x %g", scale_h, scale_w);
  const int Py = imresize_taps(scale_h), Px = imresize_taps(scale_w);
  char* sp = static_cast<char*>(scratch);
  auto carve = [&](int64_t bytes) { char* q = sp; sp += (bytes + 15) / 16 * 16; return q; };
  double* wy = reinterpret_cast<double*>(carve((int64_t)H * Py * 8));
  int* iy = reinterpret_cast<int*>(carve((int64_t)H * Py * 4));
  double* wx = reinterpret_cast<double*>(carve((int64_t)W * Px * 8));
  int* ix = reinterpret_cast<int*>(carve((int64_t)W * Px * 4));
  ProfScope prof(PROF_OTHER, 4.0 * planes * ((double)h * w + (double)H * W), stream, "imresize");
  imresize_table_kernel<<<(unsigned)ceil_div(H, 128), 128, 0, stream>>>(wy, iy, h, H, scale_h, Py, method);
  HSIDM_TRY(after_launch("imresize_table_kernel"));
  imresize_table_kernel<<<(unsigned)ceil_div(W, 128), 128, 0, stream>>>(wx, ix, w, W, scale_w, Px, method);
  HSIDM_TRY(after_launch("imresize_table_kernel"));
  dim3 grid((unsigned)ceil_div((int64_t)H * W, 256), (unsigned)planes);
  imresize_kernel<<<grid, 256, 0, stream>>>(src, dst, h, w, H, W, wy, iy, Py, wx, ix, Px);
  return after_launch("imresize_kernel");
}

static int ssim_tiles(int H, int W, int* tiles_x) {
  if (H < kSsimWin || W < kSsimWin) return 0;
  const int tx = (int)ceil_div(W - kSsimWin + 1, kSsimTX), ty = (int)ceil_div(H - kSsimWin + 1, kSsimTY);
  if (tiles_x) *tiles_x = tx;
  return tx * ty;
}

int64_t quality_assessment_scratch_bytes(int N, int C, int H, int W) {
  const int64_t metrics = (quality_metrics_scratch_bytes(N, C, H * W) + 15) / 16 * 16;
  return metrics + 48 + (int64_t)N * 2 * 4 + (int64_t)N * C * 6 * 8 + (int64_t)N * C * ssim_tiles(H, W, nullptr) * 8;
}

int quality_assessment(const float* truth, const float* pred, int N, int C, int H, int W, float ratio, void* scratch, float* out,
                       cudaStream_t stream) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || C > 4096 || C > 65535 || (int64_t)N * C > 65535)
    HSIDM_FAIL(HSIDM_BAD_SHAPE, "quality_assessment: bad shape N=%d C=%d %dx%d (N*C <= 65535)", N, C, H, W);
  if (!(ratio > 0.f)) HSIDM_FAIL(HSIDM_BAD_ARG, "quality_assessment: ratio must be positive");
  char* sp = static_cast<char*>(scratch);
  auto carve = [&](int64_t bytes) { char* q = sp; sp += (bytes + 15) / 16 * 16; return q; };
  void* metrics_scratch = carve(quality_metrics_scratch_bytes(N, C, H * W));
  float* mpsnr_sam = reinterpret_cast<float*>(carve((int64_t)N * 2 * 4));
  double* sums = reinterpret_cast<double*>(carve((int64_t)N * C * 6 * 8));
  int tiles_x = 0;
  const int tiles = ssim_tiles(H, W, &tiles_x);
  double* ssim_part = tiles ? reinterpret_cast<double*>(carve((int64_t)N * C * tiles * 8)) : nullptr;
  HSIDM_TRY(quality_metrics(truth, pred, N, C, H * W, 1.0f, metrics_scratch, mpsnr_sam, stream));
  ProfScope prof(PROF_OTHER, 2 * 8.0 * N * C * (double)H * W, stream, "assessment");
  band_sums_kernel<<<dim3(C, N), 256, 0, stream>>>(truth, pred, C, H * W, sums);
  HSIDM_TRY(after_launch("band_sums_kernel"));
  if (tiles) {
    ssim_kernel<<<dim3(tiles, N * C), dim3(kSsimTX, kSsimTY), 0, stream>>>(truth, pred, H, W, tiles_x, 1.0f, ssim_part);
    HSIDM_TRY(after_launch("ssim_kernel"));
  }
  assess_finalize_kernel<<<N, 128, 0, stream>>>(sums, ssim_part, tiles, mpsnr_sam, C, H, W, ratio, out);
  return after_launch("assess_finalize_kernel");
}

}  // namespace hsidm

namespace hsidm {
int blend_tiles(const float* tiles, const int* ys, int ny, const int* xs, int nx, int C, int tile, int overlap, int H, int W,
                float* out, cudaStream_t stream) {
  if (ny <= 0 || nx <= 0 || C <= 0 || tile <= 0 || overlap < 0 || 2 * overlap > tile || H < tile || W < tile)
    HSIDM_FAIL(HSIDM_BAD_SHAPE, "blend_tiles: bad geometry (%dx%d tiles of %d, overlap %d, scene %dx%d)", ny, nx, tile, overlap, H, W);
  if (H > 65535) HSIDM_FAIL(HSIDM_BAD_SHAPE, "blend_tiles: scene height %d exceeds 65535 rows", H);
  ProfScope prof(PROF_OTHER, 4.0 * ((double)ny * nx * C * tile * tile + (double)C * H * W), stream, "blend_tiles");
  blend_kernel<<<dim3((unsigned)ceil_div(W, 256), (unsigned)H), 256, 0, stream>>>(tiles, ys, ny, xs, nx, C, tile, overlap, H, W, out);
  return after_launch("blend_kernel");
}
}  // namespace hsidm
