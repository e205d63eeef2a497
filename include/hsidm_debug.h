/*
 * hsidm_debug.h - test hooks into single kernels of libhsidm_b200 (used by tests/, never by the product path).
 * All pointers are device pointers; activations are NHWC in the element type selected by `precision`
 * (fp32 or bf16) unless the layout argument says NCHW fp32. Each call synchronises the device before returning.
 */
#ifndef HSIDM_DEBUG_H_
#define HSIDM_DEBUG_H_
#include "hsidm.h"
#ifdef __cplusplus
extern "C" {
#endif

/* One convolution as the executors issue it (nn.Conv2d of unet.py:62,71,87,102,121,122 / AE.py / common.py).
 * backend: 0 = CUDA-core kernel, 1 = tcgen05 tensor-core kernel (fails if the shape does not fit),
 *          2 = the executors' dispatcher (tensor core when possible, lowering stride-2 / upsample).
 * weight: fp32 [Cout, c0+c1, k, k] (reference layout); bias/nbias/resid may be NULL.
 * src_layout / out_layout: 0 = NHWC activation type, 1 = NCHW fp32. */
HSIDM_API int hsidm_debug_conv2d(int backend, int precision, const void* src0, int c0, const void* src1, int c1, int src_layout,
                       int N, int H, int W, int up, int stride, const float* weight, const float* bias, int Cout,
                       int ksize, const float* nbias, int64_t nbias_stride, int act, float scale, const void* resid,
                       void* out, int out_layout);

/* GroupNorm(+Swish) over the concatenation of two NHWC tensors (unet.py:84). */
HSIDM_API int hsidm_debug_groupnorm(int precision, const void* x0, int c0, const void* x1, int c1, int N, int HW, int groups,
                          const float* gamma, const float* beta, float eps, int swish, void* out);

/* Kernel-selection knobs for A/B tests: no_halo = 1 forces the per-tap tcgen05 kernel for every 3x3 conv;
 * variant is a bit mask of developer switches (0 = the production routing): 1, 2, 4 = timing experiments that skip the
 * epilogue / weight loads / halo loads (results invalid); 8 = GroupNorm through a normalised copy instead of fused into
 * the consuming convolution's load path; 16 = 64-output-channel convs on the narrow <2,64> halo tile instead of <4,64>;
 * 32 = no CTA-pair (cta_group::2) tile for Cout % 256 == 0; 64 / 128 = CTA-pair tiles also for the 128- / 64-channel tiles;
 * 256 = GroupNorm statistics through gn_finalize launches instead of the producing convolution's tail;
 * 512 = self-attention with q, k, v materialised instead of the folded projections (Wk^T Wq, Wout Wv);
 * 1024 = no occupancy-based narrowing of the halo tile at small batches (always the full-batch tile shape);
 * 4096 = attention scores + softmax and P.V as two GEMM launches instead of the fused attn_flash kernel;
 * 16384 = the attention's transposing GroupNorm with one block per 64x64 tile instead of blocks walking several tiles. */
HSIDM_API int hsidm_debug_conv_mode(int no_halo, int variant);

/* Developer probe: when device_counters is non-null every halo-kernel launch writes 8 int64 cycle counters per CTA
 * ([a_empty wait, b_empty wait, tmem_empty wait, a_full wait, b_full wait, MMA-issuer total, tmem_full wait,
 * epilogue total]); null switches the probe off.  The buffer must hold 8 * (number of SMs) values. */
HSIDM_API int hsidm_debug_halo_timing(long long* device_counters);

/* Reads and clears the tensor-core kernel's barrier-timeout flag (0 = healthy). */
HSIDM_API int hsidm_debug_tc_error_flag(int* value);

#ifdef __cplusplus
}
#endif
#endif
