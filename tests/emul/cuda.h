// cuda.h — HOST STAND-IN (tests/emul only): the driver-API types the product's headers and tensor-map setup use.
#pragma once
#include <cstdint>
typedef uint64_t cuuint64_t;
typedef uint32_t cuuint32_t;
typedef enum { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 } CUresult;
typedef enum { CU_TENSOR_MAP_DATA_TYPE_FLOAT64 = 10 } CUtensorMapDataType;
typedef enum { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 } CUtensorMapInterleave;
typedef enum { CU_TENSOR_MAP_SWIZZLE_NONE = 0 } CUtensorMapSwizzle;
typedef enum { CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3 } CUtensorMapL2promotion;
typedef enum { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 } CUtensorMapFloatOOBfill;
// what cuTensorMapEncodeTiled records for a 3-D float64 tensor (the real one is 128 opaque bytes)
struct CUtensorMap
{
    void* base;
    uint64_t dim[3];    // elements, x fastest
    uint64_t stride[2]; // bytes: row, plane
    uint32_t box[3];
    char pad[128 - 8 - 24 - 16 - 12];
};
