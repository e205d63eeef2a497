// Library-level C ABI entry points and the test hooks of include/hsidm_debug.h.
#include "../../include/hsidm_debug.h"
#include <algorithm>

#include "net.cuh"

using namespace hsidm;

extern "C" {

int hsidm_version(void) { return HSIDM_VERSION; }
const char* hsidm_last_error(void) { return g_last_error.c_str(); }
int64_t hsidm_launch_count(void) { return g_launches; }

int hsidm_prof_enable(int on) {
  prof_reset();
  g_prof_on = on != 0;
  return HSIDM_OK;
}
int hsidm_prof_dump(const char* path) {
  if (!path) HSIDM_FAIL(HSIDM_BAD_ARG, "null path");
  return prof_dump(path);
}
int hsidm_prof_read(int kind, double* ms, double* work, int64_t* launches) {
  if (!ms || !work || !launches || kind < 0 || kind >= PROF_KINDS) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_prof_read: bad argument");
  return prof_read(kind, ms, work, launches);
}

int hsidm_debug_conv2d(int backend, int precision, const void* src0, int c0, const void* src1, int c1, int src_layout,
                       int N, int H, int W, int up, int stride, const float* weight, const float* bias, int Cout,
                       int ksize, const float* nbias, int64_t nbias_stride, int act, float scale, const void* resid,
                       void* out, int out_layout) {
  if (!src0 || !weight || !out) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_debug_conv2d: null argument");
  ParamStore ps;
  ConvW w = make_conv(ps, "w", c0 + c1, Cout, ksize, bias != nullptr);
  HSIDM_TRY(ps.alloc_all());
  int64_t wshape[4] = {Cout, c0 + c1, ksize, ksize};
  HSIDM_TRY(ps.set("w.weight", weight, wshape, 4));
  int64_t bshape[1] = {Cout};
  if (bias) HSIDM_TRY(ps.set("w.bias", bias, bshape, 1));
  if (precision == HSIDM_BF16) HSIDM_TRY(conv_tc_init());
  int s = pack_conv(ps, w, precision == HSIDM_BF16);
  if (s == HSIDM_OK && up && precision == HSIDM_BF16) s = pack_conv_up(ps, w);
  if (s == HSIDM_OK && stride == 2 && precision == HSIDM_BF16) s = pack_conv_s2(ps, w);
  if (s != HSIDM_OK) {
    free_conv(w);
    return s;
  }
  ConvOp op;
  op.src[0].p = src0, op.src[0].C = c0, op.src[0].layout = src_layout;
  if (c1) op.src[1].p = src1, op.src[1].C = c1, op.src[1].layout = src_layout;
  op.N = N, op.Hin = H, op.Win = W, op.up = up, op.stride = stride, op.ksize = ksize;
  const int He = up ? 2 * H : H, We = up ? 2 * W : W;
  op.Hout = stride == 2 ? (He + 1) / 2 : He, op.Wout = stride == 2 ? (We + 1) / 2 : We;
  op.Cout = Cout;
  op.nbias = nbias, op.nbias_stride = nbias_stride, op.act = act, op.scale = scale, op.resid = resid;
  op.out = out, op.out_layout = out_layout;
  if (backend == 2) {
    Exec ex;
    ex.prec = precision;
    ex.arena.begin(true), ex.dry = true;
    run_conv(ex, op, w, ps);
    s = ex.arena.reserve(ex.arena.peak() + 1024);
    if (s == HSIDM_OK) {
      ex.arena.begin(false), ex.dry = false, ex.status = HSIDM_OK;
      run_conv(ex, op, w, ps);
      s = ex.status;
    }
    if (cudaDeviceSynchronize() != cudaSuccess && s == HSIDM_OK) {
      set_last_error("kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
      s = HSIDM_CUDA_ERROR;
    }
    free_conv(w);
    return s;
  }
  op.w_f32 = w.w_f32, op.w_bf16 = w.w_bf16, op.bias = ps.dev(w.pb);
  s = backend == 1 ? conv_tc(op, nullptr) : conv_simt(op, precision, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess && s == HSIDM_OK) {
    set_last_error("kernel failed: %s", cudaGetErrorString(e));
    s = HSIDM_CUDA_ERROR;
  }
  free_conv(w);
  return s;
}

int hsidm_debug_groupnorm(int precision, const void* x0, int c0, const void* x1, int c1, int N, int HW, int groups,
                          const float* gamma, const float* beta, float eps, int swish, void* out) {
  void* scratch = nullptr;
  unsigned* tickets = nullptr;
  float* stats = nullptr;
  HSIDM_CUDA(cudaMalloc(&scratch, gn_scratch_bytes(c0, c1, N, HW, groups) + 16));
  HSIDM_CUDA(cudaMalloc(&tickets, sizeof(unsigned) * N));
  HSIDM_CUDA(cudaMalloc(&stats, sizeof(float) * 2 * N * groups));
  HSIDM_CUDA(cudaMemset(tickets, 0, sizeof(unsigned) * N));
  int s = gn_stats(x0, c0, x1, c1, N, HW, groups, eps, scratch, tickets, stats, precision, nullptr);
  if (s == HSIDM_OK) s = gn_apply(x0, c0, x1, c1, N, HW, groups, stats, gamma, beta, swish, out, precision, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(scratch);
  cudaFree(tickets);
  cudaFree(stats);
  if (e != cudaSuccess && s == HSIDM_OK) {
    set_last_error("kernel failed: %s", cudaGetErrorString(e));
    s = HSIDM_CUDA_ERROR;
  }
  return s;
}

int hsidm_bicubic_upsample(const float* lr, float* sr, int N, int C, int h, int w, int scale, int clamp01, hsidm_stream stream) {
  if (!lr || !sr) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_bicubic_upsample: null argument");
  if (N <= 0 || C <= 0) HSIDM_FAIL(HSIDM_BAD_SHAPE, "hsidm_bicubic_upsample: bad batch %d x %d", N, C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // (image, band) planes go along gridDim.y: at most 65535 per launch
  const long long planes = (long long)N * C;
  for (long long p0 = 0; p0 < planes; p0 += 65535) {
    const int cnt = (int)std::min<long long>(65535, planes - p0);
    HSIDM_TRY(bicubic_upsample(lr + p0 * h * w, sr + p0 * (long long)h * scale * w * scale, cnt, h, w, scale, clamp01, st));
  }
  return HSIDM_OK;
}

int hsidm_quality_metrics(const float* truth, const float* pred, int N, int C, int H, int W, float* out, hsidm_stream stream) {
  if (!truth || !pred || !out) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_quality_metrics: null argument");
  if (H <= 0 || W <= 0) HSIDM_FAIL(HSIDM_BAD_SHAPE, "hsidm_quality_metrics: bad image size %dx%d", H, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* scratch = nullptr;
  HSIDM_CUDA(cudaMallocAsync(&scratch, (size_t)quality_metrics_scratch_bytes(N, C, H * W), st));
  const int s = quality_metrics(truth, pred, N, C, H * W, 1.0f, scratch, out, st);
  HSIDM_CUDA(cudaFreeAsync(scratch, st));
  return s;
}

int hsidm_imresize(const float* in, float* out, int planes, int h, int w, int out_h, int out_w, double scale_h, double scale_w,
                   int method, hsidm_stream stream) {
  if (!in || !out) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_imresize: null argument");
  if (planes <= 0 || h <= 0 || w <= 0 || out_h <= 0 || out_w <= 0)
    HSIDM_FAIL(HSIDM_BAD_SHAPE, "hsidm_imresize: bad shape %d x %dx%d -> %dx%d", planes, h, w, out_h, out_w);
  if (scale_h < 0.0 || scale_w < 0.0) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_imresize: negative scale");
  if (scale_h == 0.0) scale_h = (double)out_h / h;
  if (scale_w == 0.0) scale_w = (double)out_w / w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* scratch = nullptr;
  HSIDM_CUDA(cudaMallocAsync(&scratch, (size_t)imresize_scratch_bytes(out_h, out_w, scale_h, scale_w), st));
  int s = HSIDM_OK;
  for (long long p0 = 0; p0 < planes && s == HSIDM_OK; p0 += 65535) {   // planes go along gridDim.y
    const int cnt = (int)std::min<long long>(65535, planes - p0);
    s = imresize(in + p0 * h * w, out + p0 * (long long)out_h * out_w, cnt, h, w, out_h, out_w, scale_h, scale_w, method, scratch, st);
  }
  HSIDM_CUDA(cudaFreeAsync(scratch, st));
  return s;
}

int hsidm_quality_assessment(const float* truth, const float* pred, int N, int C, int H, int W, float ratio, float* out,
                             hsidm_stream stream) {
  if (!truth || !pred || !out) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_quality_assessment: null argument");
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) HSIDM_FAIL(HSIDM_BAD_SHAPE, "hsidm_quality_assessment: bad shape %d x %d x %dx%d", N, C, H, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* scratch = nullptr;
  HSIDM_CUDA(cudaMallocAsync(&scratch, (size_t)quality_assessment_scratch_bytes(N, C, H, W), st));
  const int s = quality_assessment(truth, pred, N, C, H, W, ratio, scratch, out, st);
  HSIDM_CUDA(cudaFreeAsync(scratch, st));
  return s;
}

int hsidm_randn(float* out, int64_t n, uint64_t seed, int64_t first_element, hsidm_stream stream) {
  if (!out || n < 0 || first_element < 0 || first_element % 4) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_randn: bad argument");
  // step 0xFFFFFFFF: a Philox stream no loop index of hsidm_sample uses (those are t = 1 .. T-1)
  return randn_fill(out, n, seed, 0xFFFFFFFFu, (uint64_t)first_element / 4, static_cast<cudaStream_t>(stream));
}

int hsidm_blend_tiles(const float* tiles, const int32_t* ys, int ny, const int32_t* xs, int nx, int C, int tile, int overlap,
                      int H, int W, float* out, hsidm_stream stream) {
  if (!tiles || !ys || !xs || !out) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_blend_tiles: null argument");
  return blend_tiles(tiles, ys, ny, xs, nx, C, tile, overlap, H, W, out, static_cast<cudaStream_t>(stream));
}

int hsidm_check_health(void) {
  int v = 0;
  HSIDM_TRY(conv_tc_error_flag(&v));
  if (v != 0)
    HSIDM_FAIL(HSIDM_CUDA_ERROR, "a tensor-core kernel gave up waiting on a pipeline barrier (code %d): results of the calls since the "
               "last health check are invalid", v);
  return HSIDM_OK;
}

int hsidm_debug_conv_mode(int no_halo, int variant) {
  conv_tc_set_mode(no_halo, variant);
  return HSIDM_OK;
}

int hsidm_debug_halo_timing(long long* device_counters) {
  conv_halo_set_timing(device_counters);
  return HSIDM_OK;
}

int hsidm_debug_tc_error_flag(int* value) {
  if (!value) HSIDM_FAIL(HSIDM_BAD_ARG, "null argument");
  return conv_tc_error_flag(value);
}

}  // extern "C"
