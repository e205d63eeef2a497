// Shared host/device helpers for the hsidm kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/hsidm.h"

namespace hsidm {

// ---- error plumbing -------------------------------------------------------------------------------
struct Error {
  int code;
  std::string msg;
};
void set_last_error(const char* fmt, ...);
extern thread_local std::string g_last_error;
extern int64_t g_launches;  // kernels launched by this library (bench.py reports it)

#define HSIDM_FAIL(code, ...)            \
  do {                                   \
    ::hsidm::set_last_error(__VA_ARGS__); \
    return (code);                       \
  } while (0)

#define HSIDM_CUDA(expr)                                                                       \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::hsidm::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                              __LINE__);                                                       \
      return HSIDM_CUDA_ERROR;                                                                 \
    }                                                                                          \
  } while (0)

#define HSIDM_TRY(expr)        \
  do {                         \
    int _s = (expr);           \
    if (_s != HSIDM_OK) return _s; \
  } while (0)

// Makes `device` current for the duration of an entry point and restores the caller's device afterwards (a library
// call must not change torch.cuda.current_device() behind the caller's back).
class DeviceGuard {
 public:
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev_) != cudaSuccess) prev_ = -1;
    err_ = prev_ == device ? cudaSuccess : cudaSetDevice(device);
    switched_ = err_ == cudaSuccess && prev_ != device && prev_ >= 0;
  }
  ~DeviceGuard() {
    if (switched_) cudaSetDevice(prev_);
  }
  cudaError_t error() const { return err_; }

 private:
  int prev_ = -1;
  bool switched_ = false;
  cudaError_t err_ = cudaSuccess;
};
#define HSIDM_DEVICE(dev)                                                                                   \
  ::hsidm::DeviceGuard _device_guard(dev);                                                                  \
  if (_device_guard.error() != cudaSuccess) {                                                               \
    ::hsidm::set_last_error("cudaSetDevice(%d) failed: %s", (dev), cudaGetErrorString(_device_guard.error())); \
    return HSIDM_CUDA_ERROR;                                                                                \
  }

// ---- opt-in per-kernel-class timing (bench.py's roofline leg): CUDA events around each launch on its own stream ----
enum ProfKind { PROF_CONV_TC = 0, PROF_CONV_SIMT = 1, PROF_GN_STATS = 2, PROF_GN_APPLY = 3, PROF_GEMM = 4,
                PROF_POSTERIOR = 5, PROF_OTHER = 6, PROF_KINDS = 7 };
extern bool g_prof_on;
int prof_start(int kind, double work, cudaStream_t stream, const char* tag = nullptr);  // token (<0: not recording)
void prof_stop(int token, cudaStream_t stream);
void prof_reset();
int prof_read(int kind, double* ms, double* work, int64_t* launches);
int prof_dump(const char* path);  // one CSV line per recorded launch: kind,tag,work,ms
struct ProfScope {
  int token;
  cudaStream_t stream;
  ProfScope(int kind, double work, cudaStream_t s, const char* tag = nullptr)
      : token(g_prof_on ? prof_start(kind, work, s, tag) : -1), stream(s) {}
  ~ProfScope() {
    if (token >= 0) prof_stop(token, stream);
  }
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// Every kernel of the denoise step is launched with programmatic stream serialization: its CTAs may be scheduled while the
// previous kernel of the stream is still draining (as SMs free up), run their prologue (barrier init, tensor-memory
// allocation, descriptor prefetch, weight-tile prefetch: nothing that depends on the predecessor) and then block in
// griddep_wait() until the predecessor grid has completed and its writes are visible.  Rule for kernels launched through
// launch_pdl(): every thread executes griddep_wait() before its first access to global memory that is not constant for
// the lifetime of the launch (activations, statistics, per-step tables); packed weights may be read before it.
extern bool g_pdl_on;   // HSIDM_NO_PDL=1 switches the attribute off (A/B runs); the device-side calls are then no-ops
// Set per pass by the executors: PDL pays when kernels are short (small batches: -5 % per step at 1-11 latents) and
// measured neutral to slightly negative at 176 latents, where each kernel runs for tens of microseconds.
extern bool g_pdl_pass;
#if defined(__CUDACC__)
static __device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
static __device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_pdl_on && g_pdl_pass) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster_x, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr, cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

inline int after_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_last_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return HSIDM_CUDA_ERROR;
  }
  return HSIDM_OK;
}

// ---- activation element types ------------------------------------------------------------------------
using bf16 = __nv_bfloat16;

template <typename T>
struct ActTraits;
template <>
struct ActTraits<float> {
  static constexpr int kPrecision = HSIDM_F32;
};
template <>
struct ActTraits<bf16> {
  static constexpr int kPrecision = HSIDM_BF16;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ bf16 from_f32<bf16>(float v) {
  return __float2bfloat16_rn(v);
}

// 8 consecutive activations <-> 8 floats (16 B for bf16, 32 B for fp32); pointers must be so aligned.
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}

__device__ __forceinline__ float swish_f(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- activation tensor handle (NHWC, element type decided by the context precision) ----------------------
struct Act {
  void* p = nullptr;
  int N = 0, H = 0, W = 0, C = 0;
  float* stats = nullptr;  // [N][slots][C][2] partial (sum, sumsq) left by the producing convolution, or null
  int slots = 0;
  int64_t numel() const { return (int64_t)N * H * W * C; }
};

// ---- stream-ordered workspace arena --------------------------------------------------------------------
// All kernels of one context run on one stream in program order, so a block can be handed out again as soon
// as the host code that consumed it has *enqueued* its last reader.  Dry mode only measures the peak.
class Arena {
 public:
  ~Arena();
  int reserve(int64_t bytes);  // (re)allocate backing store; synchronises when it has to grow
  void begin(bool dry);        // start a pass; every block must have been freed
  void* alloc(int64_t bytes);
  void free(void* p);
  int64_t peak() const { return peak_; }
  int64_t capacity() const { return cap_; }
  bool dry() const { return dry_; }
  bool failed() const { return failed_; }

 private:
  struct Block {
    int64_t off, size;
    bool used;
  };
  char* base_ = nullptr;
  int64_t cap_ = 0, peak_ = 0, top_ = 0;
  bool dry_ = false, failed_ = false;
  static constexpr int kMaxBlocks = 512;
  Block blocks_[kMaxBlocks];
  int nblocks_ = 0;
};

}  // namespace hsidm
