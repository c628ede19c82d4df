// Packed FFN weight stream shared by the standalone tcgen05 FFN kernel and the fused rollout kernel, and the
// two-term fp16 operand split ("f16s") both of them compute in.
//
// f16s split of an fp32 value x, pre-scaled by a power of two S:  x' = S x,  hi = x' rounded to 11 significant
// bits (= fp16(x')),  lo = x' - hi  (exact in fp32, |lo| <= 2^-12 |x'|);  both parts are stored as fp16.
//   activations (A operand): S = kAScale;   weights / logit keys (B operand): S = kWScale / kLkScale
//   S_a S_w a w ~= A_hi B_hi + A_lo B_hi + A_hi B_lo        (dropped term a_lo w_lo <= 2^-24 |a w|)
// i.e. three kind::f16 tcgen05 MMAs (K = 16 each, twice the TF32 rate) into ONE fp32 accumulator give the same
// fp32-faithful product as 3xTF32 at half the tensor time and half the operand bytes (the weight stream, 512 KB per
// decode step, is what bounds the FFN: one SM ingests ~27 B/clk through TMA).  The scales keep the lo parts in the
// normal fp16 range for every value that matters (|a| >= 2^-6, |w| >= 2^-10; smaller ones carry an absolute error
// <= 2^-25 / S) and are undone exactly in the epilogues.  |S x| >= 65520 (|a| >= 4094, |w| >= 255) overflows fp16
// and surfaces as non-finite logits (RRNCO_DEV_NAN_LOGITS), never silently.
//
// Stream layout: 16 slices in the order the tensor pipe consumes them.  job j = s >> 1 in
//   G1(0) G1(1) G2(0) G1(2) G2(1) G1(3) G2(2) G2(3)      (G1(c): W1 rows of hidden chunk c, G2(c): W2 columns of c)
// so that epilogue 1 of chunk c overlaps GEMM1 of chunk c+1; s & 1 selects 64 k values.  Each slice is one
// contiguous 32 KB block [B_hi | B_lo][16-byte K chunk (8)][row (128)][8 halves] == the shared-memory core-matrix
// layout (tc05.cuh), so one cp.async.bulk moves it and UMMA descriptors address its four K steps.  Large slices
// matter: a bulk copy costs ~250 cycles of fixed overhead plus ~1 cycle per 66 bytes, and the tensor pipe needs a
// K step (two variants, 8 KB) every ~290 cycles.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rrnco {
constexpr int kFRows = 128;
constexpr int kFSliceK = 64;                                     // k values per slice
constexpr int kFKSteps = kFSliceK / 16;                          // kind::f16 MMAs (K = 16) per slice and split term
constexpr int kFSlicesPerJob = kFRows / kFSliceK;                // a job = one 128 x 128 x 128 GEMM
constexpr int kFStages = 4;                                      // ring stages (128 KB in flight)
constexpr int kFSlices = 8 * kFSlicesPerJob;
constexpr int kFVariantHalves = kFRows * kFSliceK;               // B_hi or B_lo: 16 KB
constexpr int kFSliceHalves = 2 * kFVariantHalves;
constexpr uint32_t kFSliceBytes = kFSliceHalves * 2;             // 32 KB
constexpr uint32_t kLboTile = kFRows * 16;                       // bytes between the 16-byte K chunks of a 128-row tile
constexpr uint32_t kSbo = 128;                                   // bytes between 8-row groups
constexpr int64_t kFfnPackedBytes = (int64_t)kFSlices * kFSliceBytes;  // 512 KB
constexpr float kWScale = 256.0f;                                // weights; |w| < 255 representable
constexpr float kLkScale = 16.0f;                                // logit keys; |Lk| < 4094 representable
constexpr float kAScale = 16.0f;                                 // activations; |a| < 4094 representable
constexpr float kKvScale = 16.0f;                                // attention keys / values; |x| < 4094 representable
constexpr uint32_t kKvOffV = 65536, kKvOffLk = 131072;           // byte offsets of the V and logit-key tiles in a slot
constexpr uint32_t kKvSlotBytes = 196608;                        // packed K | V | Lk tiles of one instance (per-SM workspace slot)
constexpr int kKvSlots = 256;                                    // >= %nsmid

// chunk / half of job j (see above)
__host__ __device__ constexpr int ffn_job_chunk(int j) { return j == 0 ? 0 : j == 1 ? 1 : j == 2 ? 0 : j == 3 ? 2 : j == 4 ? 1 : j == 5 ? 3 : j == 6 ? 2 : 3; }
__host__ __device__ constexpr int ffn_job_half(int j) { return j == 0 ? 0 : j == 1 ? 0 : j == 2 ? 1 : j == 3 ? 0 : j == 4 ? 1 : j == 5 ? 0 : j == 6 ? 1 : 1; }

#ifdef __CUDACC__
// Packed fp32 pairs (sm_100: FFMA2 / FADD2, two IEEE operations per issue slot; bit-identical to the scalar forms).  The
// element-wise passes of the rollout kernels are bound by instruction issue, so halving their FMA / ADD count is a direct gain.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
// pair (x0, x1) (already scaled) -> packed hi / lo fp16 words: F2FP, unpack, one FADD2, F2FP
__device__ __forceinline__ void f16s_split_pair(float2 x, uint32_t& hi, uint32_t& lo) {
  const __half2 hh = __floats2half2_rn(x.x, x.y);
  const float2 hf = __half22float2(hh);
  const float2 r = fadd2(x, make_float2(-hf.x, -hf.y));
  const __half2 ll = __floats2half2_rn(r.x, r.y);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}
// pair (x0, x1), pre-scaled by `scale` -> packed hi / lo fp16 words (low half = x0): hi = fp16(x) (round to nearest, 11
// significant bits; one F2FP per pair), lo = fp16(x - hi) with the subtraction exact in fp32.  6 instructions per pair.
__device__ __forceinline__ void f16s_split2(float x0, float x1, float scale, uint32_t& hi, uint32_t& lo) {
  x0 *= scale;
  x1 *= scale;
  const __half2 hh = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(hh);
  const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}
#endif

// launches the packing kernel on `st` (defined in ffn_tc_kernel.cu); `packed` holds kFfnPackedBytes
int pack_ffn_weights(const float* w1, const float* w2, void* packed, cudaStream_t st);
}  // namespace rrnco
