// Host-side creation of TMA tensor maps (cuTensorMapEncodeTiled resolved through the runtime, so the
// library does not link libcuda).  Maps are cached by their full description.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace lsvs {

// 2-D bf16 tensor, inner dimension contiguous: dims {inner, outer}, row pitch in bytes, 128-byte swizzle.
// box_inner must be 64 elements (= 128 bytes = one swizzle row).  Out-of-bounds reads return zero.
// Returns nullptr on failure (lsvs::g_err set).  The returned pointer stays valid for the process lifetime.
const CUtensorMap* tmap_2d_bf16(const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                                uint32_t box_inner, uint32_t box_outer);

// 2-D fp32 tensor (row-major, `inner` columns contiguous), box 32 columns x `box_outer` rows, 128-byte swizzle:
// destination map of the TMA reduce-add residual epilogue (csrc/gemm.cu).
const CUtensorMap* tmap_2d_f32_box32(const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_outer);

}  // namespace lsvs
