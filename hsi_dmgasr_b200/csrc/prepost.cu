// The steps either side of the sampling path, on the GPU:
//   * bicubic x`scale` pre-upsampling of the low-resolution cube (reference: torch.nn.functional.interpolate(...,
//     scale_factor=4, mode='bicubic'), sr_gae.py:72) - PyTorch's convention: align_corners=False, A = -0.75, source
//     index (o + 0.5)/scale - 0.5 without clamping, neighbour indices clamped to the border;
//   * the two validation metrics the parity gates are stated in, per cube, after the driver's clamp to [0,1]
//     (sr_gae.py:474-475): MPSNR (eval_hsi.py:110-121) and SAM (eval_hsi.py:47-65).
// Both are HBM-bound: 4 B read per low-res element + 4 B written per output element; 8 B read per element pair.
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
