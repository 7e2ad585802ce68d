"""GPU probe: does cuTensorMapEncodeTiled accept a tensor map whose dim-1 stride (48 B) is SMALLER than the dim-0
extent (64 floats = 256 B), i.e. overlapping rows -- the space-to-depth stem's sliding window?  Diagnostic only."""
import torch
from cuda.bindings import driver as cu

torch.zeros(1, device="cuda")
buf = torch.zeros(2 * 115 * 115 * 12 + 64, device="cuda")
for dims, strides in (((64, 112, 115, 2), (48, 115 * 48, 115 * 115 * 48)),
                      ((48, 112, 115, 2), (48, 115 * 48, 115 * 115 * 48)),
                      ((12, 115, 115, 2), (48, 115 * 48, 115 * 115 * 48))):
    for swz in (cu.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B, cu.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B):
        box = (32 if dims[0] >= 32 else dims[0], 16, 8, 1)
        r = cu.cuTensorMapEncodeTiled(cu.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, buf.data_ptr(),
                                      [cu.cuuint64_t(d) for d in dims], [cu.cuuint64_t(s) for s in strides],
                                      [cu.cuuint32_t(b) for b in box], [cu.cuuint32_t(1)] * 4,
                                      cu.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                      cu.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      cu.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
        print(dims, strides, str(swz).split(".")[-1], "->", r[0])
