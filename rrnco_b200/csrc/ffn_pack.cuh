// Packed FFN weight stream shared by the standalone tcgen05 FFN kernel and the fused rollout kernel.
//   slice s in [0, 64): chunk c = s >> 4 (128 hidden units), half = (s >> 3) & 1 (0: W1 rows of the chunk,
//   1: W2 columns of the chunk), ks = s & 7 (16 k values).  Each slice is one contiguous 16 KB block
//   [hi | lo][16-byte K chunk c4 (4)][row (128)][4 floats]  ==  the shared-memory core-matrix layout (tc05.cuh),
//   so a single cp.async.bulk moves it and two UMMA descriptors (hi, lo) address it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rrnco {
constexpr int kFRows = 128;
constexpr int kFSliceK = 16;
constexpr int kFStages = 4;
constexpr int kFSlices = 64;
constexpr int kFSliceFloats = 2 * kFRows * kFSliceK;  // hi | lo
constexpr uint32_t kFSliceBytes = kFSliceFloats * 4;  // 16 KB
constexpr uint32_t kLboTile = kFRows * 16;            // bytes between the 16-byte K chunks of a 128-row tile
constexpr uint32_t kSbo = 128;                        // bytes between 8-row groups
constexpr int64_t kFfnPackedFloats = (int64_t)kFSlices * kFSliceFloats;  // 1 MB

// launches the packing kernel on `st` (defined in ffn_tc_kernel.cu)
int pack_ffn_weights(const float* w1, const float* w2, float* packed, cudaStream_t st);
}  // namespace rrnco
