// Host-side helpers shared by every translation unit of libb200lp.so: error reporting across the C ABI and
// TMA tensor-map construction through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200lp.h"

namespace b200lp {

void set_error(const char* fmt, ...);
// number of kernels this library has launched (bench.py's `gpu_launches`)
void count_launch(int n = 1);

#define B200LP_CHECK_CUDA(expr)                                                               \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::b200lp::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return B200LP_ECUDA;                                                              \
        }                                                                                     \
    } while (0)

#define B200LP_REQUIRE(cond, ...)             \
    do {                                      \
        if (!(cond)) {                        \
            ::b200lp::set_error(__VA_ARGS__); \
            return B200LP_EINVAL;             \
        }                                     \
    } while (0)

// cuTensorMapEncodeTiled for an fp32 tensor of rank `rank` (dims innermost first), 128-byte swizzle,
// zero fill for out-of-bounds elements.  strides_bytes has rank-1 entries (dim 1..rank-1).
// swizzle_base32 = false: CU_TENSOR_MAP_SWIZZLE_128B (16-byte chunks, K-major UMMA operands);
// swizzle_base32 = true : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (32-byte chunks; the only layout the tensor core accepts
//                         for MN-major 32-bit operands, UMMA layout type SWIZZLE_128B_BASE32B).
enum TmapDtype { kTmapF32 = 0, kTmapBF16 = 1 };
int encode_tmap(CUtensorMap* out, const void* base, TmapDtype dtype, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, bool swizzle_base32);
inline int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, bool swizzle_base32 = false) {
    return encode_tmap(out, base, kTmapF32, rank, dims, strides_bytes, box, swizzle_base32);
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int ilog2_exact(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return ((1 << l) == v) ? l : -1;
}

}  // namespace b200lp
