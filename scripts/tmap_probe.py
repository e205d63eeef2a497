"""Probe: does cuTensorMapEncodeTiled accept NON-monotonic strides (dims {C, W, N, H} of an NHWC tensor: the image index
before the row index)?  The halo kernel's two-images-per-tile form for 8x8 images would need it."""
import ctypes as C
import torch

lib = C.CDLL("libcuda.so.1")
x = torch.zeros(4, 8, 8, 512, dtype=torch.bfloat16, device="cuda")
tmap = (C.c_uint64 * 16)()
N, H, W, Cc = 4, 8, 8, 512


def enc(dims, strides, box):
    d = (C.c_uint64 * 4)(*dims)
    s = (C.c_uint64 * 3)(*strides)
    b = (C.c_uint32 * 4)(*box)
    e = (C.c_uint32 * 4)(1, 1, 1, 1)
    # CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 = 9, INTERLEAVE_NONE = 0, SWIZZLE_128B = 3, L2_PROMOTION_L2_256B = 3, OOB_FILL_NONE = 0
    return lib.cuTensorMapEncodeTiled(C.byref(tmap), 9, 4, C.c_void_p(x.data_ptr()), d, s, b, e, 0, 3, 3, 0)


print("monotonic  {C,W,H,N}:", enc((Cc, W, H, N), (Cc * 2, W * Cc * 2, H * W * Cc * 2), (64, 10, 10, 1)))
print("non-monot. {C,W,N,H}:", enc((Cc, W, N, H), (Cc * 2, H * W * Cc * 2, W * Cc * 2), (64, 10, 2, 10)))
