// Training hand-off, residual FFN of RRNet_PointerAttention (rrnco/models/decoder.py:272-277, 296) over the rows of the
// batched replay (one row = one decode step of one rollout; rrnco_b200/training.py):
//
//   ffn_train_kernel<0>   forward        hidden = relu(x W1^T + b1)        y = hidden W2^T + b2 + x     (+ relu bit mask, + hidden)
//   ffn_train_kernel<1>   backward-data  dhid   = (dy W2) . mask           dx = dhid W1 + dy            (+ dhid)
//                         = the forward chain on the transposed weights with the activation replaced by the saved mask
//   xty_kernel            weight gradients  C[512,128] += X^T Y  (dW1 = dhid^T x, dW2^T = hidden^T dy) + column sums (db1, db2)
//
// All contractions run on tcgen05 (kind::f16) in the fp32-faithful fp16 hi|lo operand split of ffn_pack.cuh: three MMAs per
// product into one fp32 TMEM accumulator.  Gradients are small (1e-6 .. 1e-9): the caller passes a power-of-two scale that
// is applied before the split and undone exactly in the epilogue.
//
// ffn_train_kernel: persistent, one CTA per SM (512 TMEM columns: two ping-pong hidden accumulators, the output accumulator,
// the fp16 hi|lo hidden operand of GEMM2), 10 warps: 0-7 operand conversion + epilogues (thread per row, two column halves),
// 8 = TMA producer of the 16 packed weight slices per tile (32 KB each, 4-stage ring), 9 = one elected MMA issuer.
// Tensor-pipe order per tile  G1(0) G1(1) G2(0) G1(2) G2(1) G1(3) G2(2) G2(3): epilogue 1 of chunk c runs under GEMM1 of chunk
// c + 1, and the conversion of the next tile's x under GEMM2(3).
#include "../csrc/common.cuh"
#include "../csrc/tc05.cuh"
#include "../csrc/ffn_pack.cuh"
#include "../../include/rrnco_b200_train.h"

namespace rrnco {

constexpr int kTThreads = 320;
constexpr float kF16Max = 65504.f;

struct FfnTrainSmem {
  uint16_t a_hi[kFRows * kE];  // x tile, [16-byte K chunk (16)][row (128)][8 halves]
  uint16_t a_lo[kFRows * kE];
  uint16_t w[kFStages][kFSliceHalves];
  float b1[kF];
  float b2[kE];
  uint64_t bar_full[kFStages];
  uint64_t bar_empty[kFStages];
  uint64_t bar_a;      // x tile converted (256 arrivals)
  uint64_t bar_h[2];   // GEMM1 into hidden accumulator b complete
  uint64_t bar_epi;    // epilogue 1 of a chunk complete: GEMM2 operand written, accumulator free (256 arrivals)
  uint64_t bar_g2;     // GEMM2 of a chunk complete: its TMEM operand may be overwritten / output complete
  uint32_t tmem_base;
};
static_assert(sizeof(FfnTrainSmem) <= 227 * 1024, "shared memory");

// Same stream layout as ffn_pack.cuh describes: slice s = job * 2 + (64-k half), [B_hi | B_lo][K chunk (8)][row (128)][8 halves]
__global__ void pack_ffn_train_kernel(const float* __restrict__ wa, const float* __restrict__ wb, uint32_t* __restrict__ packed,
                                      uint32_t* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int kPairs = kFSliceK / 2;
  if (i >= kFSlices * kFRows * kPairs) return;
  const int kp = i % kPairs, row = (i / kPairs) & 127, s = i / (kPairs * kFRows);
  const int j = s / kFSlicesPerJob, ks = s % kFSlicesPerJob, c = ffn_job_chunk(j), half = ffn_job_half(j);
  const int k = ks * kFSliceK + kp * 2;
  float v0, v1;
  if (half == 0) {
    v0 = wa[(size_t)(c * kFRows + row) * kE + k];
    v1 = wa[(size_t)(c * kFRows + row) * kE + k + 1];
  } else {
    v0 = wb[(size_t)row * kF + c * kFRows + k];
    v1 = wb[(size_t)row * kF + c * kFRows + k + 1];
  }
  if (!(fabsf(v0) * kWScale < kF16Max) || !(fabsf(v1) * kWScale < kF16Max)) atomicOr(status, RRNCO_DEV_NAN_LOGITS);
  uint32_t hi, lo;
  f16s_split2(v0, v1, kWScale, hi, lo);
  uint32_t* dst = packed + (size_t)s * (kFSliceHalves / 2) + (kp >> 2) * (kFRows * 4) + row * 4 + (kp & 3);
  dst[0] = hi;
  dst[kFVariantHalves / 2] = lo;
}

__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

template <int kMode>
__global__ void __launch_bounds__(kTThreads, 1) ffn_train_kernel(int64_t M, const float* __restrict__ x,
                                                                const uint16_t* __restrict__ wpacked,
                                                                const float* __restrict__ b1, const float* __restrict__ b2,
                                                                const float* __restrict__ a_scale_ptr, uint32_t* __restrict__ mask,
                                                                float* __restrict__ hid_out, float* __restrict__ y,
                                                                uint32_t* __restrict__ status) {
  extern __shared__ __align__(128) unsigned char ffn_train_smem_raw[];
  FfnTrainSmem& sm = *reinterpret_cast<FfnTrainSmem*>(ffn_train_smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (M + kFRows - 1) / kFRows;

  if (warp == 0) tc05::tmem_alloc(&sm.tmem_base, 512);
  if (tid == 32) {
    for (int i = 0; i < kFStages; ++i) {
      tc05::mbar_init(&sm.bar_full[i], 1);
      tc05::mbar_init(&sm.bar_empty[i], 1);
    }
    tc05::mbar_init(&sm.bar_a, 256);
    tc05::mbar_init(&sm.bar_h[0], 1);
    tc05::mbar_init(&sm.bar_h[1], 1);
    tc05::mbar_init(&sm.bar_epi, 256);
    tc05::mbar_init(&sm.bar_g2, 1);
    tc05::fence_mbar_init();
  }
  if (kMode == 0) {
    for (int i = tid; i < kF; i += kTThreads) sm.b1[i] = b1[i];
    for (int i = tid; i < kE; i += kTThreads) sm.b2[i] = b2[i];
  }
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tbase = sm.tmem_base;
  const uint32_t t_hacc0 = tbase, t_oacc = tbase + 256, t_hhi = tbase + 384, t_hlo = tbase + 448;

  const int uwarp = __shfl_sync(0xffffffffu, warp, 0);
  if (uwarp == 8) {
    // ===== TMA producer =====
    if (tc05::elect_one()) {
      uint32_t st = 0, round = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
#pragma unroll 1
        for (int s = 0; s < kFSlices; ++s) {
          if (round > 0) tc05::mbar_wait(&sm.bar_empty[st], (round - 1) & 1);
          tc05::mbar_arrive_expect_tx(&sm.bar_full[st], kFSliceBytes);
          tc05::bulk_g2s(sm.w[st], wpacked + (size_t)s * kFSliceHalves, kFSliceBytes, &sm.bar_full[st]);
          if (++st == (uint32_t)kFStages) { st = 0; ++round; }
        }
      }
    }
    return;
  }
  if (uwarp == 9) {
    // ===== MMA issue: one elected thread, fixed order =====
    if (tc05::elect_one()) {
      const uint32_t idesc = tc05::make_idesc_f16(128, 128);
      const uint32_t a_addr = tc05::smem_u32(sm.a_hi), a_lo_off = kFRows * kE * 2;
      uint32_t st = 0, round = 0, n_a = 0, n_epi = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        tc05::mbar_wait(&sm.bar_a, n_a & 1u);
        ++n_a;
        tc05::fence_after_sync();
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          const int c = ffn_job_chunk(j), half = ffn_job_half(j);
          if (half == 1) {  // GEMM2 operand of chunk c written (and the accumulator GEMM1(c + 2) overwrites has been read)
            tc05::mbar_wait(&sm.bar_epi, n_epi & 1u);
            ++n_epi;
            tc05::fence_after_sync();
          }
#pragma unroll 1
          for (int sj = 0; sj < kFSlicesPerJob; ++sj) {
            tc05::mbar_wait(&sm.bar_full[st], round & 1u);
            tc05::fence_after_sync();
            const uint32_t b_addr = tc05::smem_u32(sm.w[st]);
#pragma unroll
            for (int kk = 0; kk < kFKSteps; ++kk) {
              const int ks = sj * kFKSteps + kk;  // K step (16 values) of the job
              const uint64_t b_hi = tc05::make_desc(b_addr + kk * 2 * kLboTile, kLboTile, kSbo);
              const uint64_t b_lo = tc05::make_desc(b_addr + kFVariantHalves * 2 + kk * 2 * kLboTile, kLboTile, kSbo);
              if (half == 0) {
                const uint32_t d = t_hacc0 + (c & 1) * 128;
                const uint64_t a_hi = tc05::make_desc(a_addr + ks * 2 * kLboTile, kLboTile, kSbo);
                const uint64_t a_lo = tc05::make_desc(a_addr + a_lo_off + ks * 2 * kLboTile, kLboTile, kSbo);
                tc05::mma_ss_f16(d, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
                tc05::mma_ss_f16(d, a_lo, b_hi, idesc, 1u);
                tc05::mma_ss_f16(d, a_hi, b_lo, idesc, 1u);
              } else {
                tc05::mma_ts_f16(t_oacc, t_hhi + ks * 8, b_hi, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                tc05::mma_ts_f16(t_oacc, t_hlo + ks * 8, b_hi, idesc, 1u);
                tc05::mma_ts_f16(t_oacc, t_hhi + ks * 8, b_lo, idesc, 1u);
              }
            }
            tc05::commit(&sm.bar_empty[st]);
            if (++st == (uint32_t)kFStages) { st = 0; ++round; }
          }
          tc05::commit(half == 0 ? &sm.bar_h[c & 1] : &sm.bar_g2);
        }
      }
    }
    return;
  }

  // ================= compute warps 0-7 =================
  const int row = (warp & 3) * 32 + lane;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int colhalf = warp >> 2;  // warps 0-3: columns [0, 64) of a 128-wide block, warps 4-7: [64, 128)
  const float a_scale = a_scale_ptr ? __ldg(a_scale_ptr) : kAScale;
  const float h_scale = kMode == 0 ? kAScale : a_scale;
  const float unscale1 = 1.0f / (a_scale * kWScale), unscale2 = 1.0f / (h_scale * kWScale);
  float amax = 0.f;  // largest scaled operand magnitude seen by this thread (fp16 overflow check)

  float4 buf[16];   // this thread's 8 (row, 8-column chunk) items of the NEXT tile, loaded while the current tile is in its epilogues
  auto load_tile = [&](int64_t m0) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = tid + it * 256, r = idx & 127, c8 = idx >> 7;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (m0 + r < M) {
        const float4* src = reinterpret_cast<const float4*>(x + (m0 + r) * kE) + c8 * 2;
        v0 = __ldg(src);
        v1 = __ldg(src + 1);
      }
      buf[2 * it] = v0;
      buf[2 * it + 1] = v1;
    }
  };
  auto convert_tile = [&]() {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = tid + it * 256, r = idx & 127, c8 = idx >> 7;
      const float4 v0 = buf[2 * it], v1 = buf[2 * it + 1];
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))) * a_scale);
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w))) * a_scale);
      uint32_t h[4], l[4];
      f16s_split2(v0.x, v0.y, a_scale, h[0], l[0]);
      f16s_split2(v0.z, v0.w, a_scale, h[1], l[1]);
      f16s_split2(v1.x, v1.y, a_scale, h[2], l[2]);
      f16s_split2(v1.z, v1.w, a_scale, h[3], l[3]);
      const int dst = c8 * (kFRows * 8) + r * 8;
      *reinterpret_cast<uint4*>(&sm.a_hi[dst]) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(&sm.a_lo[dst]) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    tc05::fence_proxy_async();
    tc05::mbar_arrive(&sm.bar_a);
  };

  uint32_t n_h[2] = {0u, 0u}, n_g2 = 0u;
  int64_t t = blockIdx.x;
  if (t < n_tiles) {
    load_tile(t * kFRows);
    convert_tile();
  }
  for (; t < n_tiles; t += gridDim.x) {
    const int64_t m0 = t * kFRows;
    const bool valid = m0 + row < M;
    if (t + gridDim.x < n_tiles) load_tile((t + gridDim.x) * kFRows);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      tc05::mbar_wait(&sm.bar_h[c & 1], n_h[c & 1] & 1u);
      ++n_h[c & 1];
      if (c > 0) {  // GEMM2(c - 1) has consumed the previous hidden operand
        tc05::mbar_wait(&sm.bar_g2, n_g2 & 1u);
        ++n_g2;
      }
      tc05::fence_after_sync();
      const uint32_t t_h = t_hacc0 + (c & 1) * 128;
      uint32_t word = 0u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col0 = colhalf * 64 + q * 16;
        uint32_t v[16], hi[8], lo[8];
        tc05::tmem_ld16(t_h + lane_base + col0, v);
        if (kMode == 1 && (q & 1) == 0) word = valid ? __ldg(mask + (m0 + row) * 16 + c * 4 + colhalf * 2 + (q >> 1)) : 0u;
        tc05::tmem_wait_ld();
        float h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (kMode == 0) {
            h[i] = fmaxf(fmaf(__uint_as_float(v[i]), unscale1, sm.b1[c * kFRows + col0 + i]), 0.f);
            word |= (h[i] > 0.f ? 1u : 0u) << ((q & 1) * 16 + i);
          } else {
            h[i] = ((word >> ((q & 1) * 16 + i)) & 1u) ? __uint_as_float(v[i]) * unscale1 : 0.f;
          }
          amax = fmaxf(amax, fabsf(h[i]) * h_scale);
        }
        if (hid_out != nullptr && valid) {
          // tile-blocked, transposed: [tile][hidden unit (512)][row (128)] - the 32 lanes of a store are 32 consecutive rows of one
          // unit (one 128-byte wavefront) instead of 32 half-filled sectors 2 KB apart; xty_kernel is the only reader
          float* dst = hid_out + t * (int64_t)(kF * kFRows) + (int64_t)(c * kFRows + col0) * kFRows + row;
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[i * kFRows] = h[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) f16s_split2(h[2 * i], h[2 * i + 1], h_scale, hi[i], lo[i]);
        tc05::tmem_st8(t_hhi + lane_base + (col0 >> 1), hi);
        tc05::tmem_st8(t_hlo + lane_base + (col0 >> 1), lo);
        if (kMode == 0 && (q & 1) == 1) {
          if (mask != nullptr && valid) mask[(m0 + row) * 16 + c * 4 + colhalf * 2 + (q >> 1)] = word;
          word = 0u;
        }
      }
      tc05::tmem_wait_st();
      tc05::fence_before_sync();
      tc05::mbar_arrive(&sm.bar_epi);
    }
    // every GEMM1 of this tile has completed (bar_h of chunk 3): the x tile may be replaced while GEMM2(3) runs
    if (t + gridDim.x < n_tiles) convert_tile();
    tc05::mbar_wait(&sm.bar_g2, n_g2 & 1u);  // GEMM2(3): output accumulator complete
    ++n_g2;
    tc05::fence_after_sync();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col0 = colhalf * 64 + q * 16;
      uint32_t v[16];
      tc05::tmem_ld16(t_oacc + lane_base + col0, v);
      tc05::tmem_wait_ld();
      if (y != nullptr && valid) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(x + (m0 + row) * kE + col0 + i));
          float4 o;
          o.x = fmaf(__uint_as_float(v[i]), unscale2, r.x);
          o.y = fmaf(__uint_as_float(v[i + 1]), unscale2, r.y);
          o.z = fmaf(__uint_as_float(v[i + 2]), unscale2, r.z);
          o.w = fmaf(__uint_as_float(v[i + 3]), unscale2, r.w);
          if (kMode == 0) {
            o.x += sm.b2[col0 + i];
            o.y += sm.b2[col0 + i + 1];
            o.z += sm.b2[col0 + i + 2];
            o.w += sm.b2[col0 + i + 3];
          }
          *reinterpret_cast<float4*>(y + (m0 + row) * kE + col0 + i) = o;
        }
      }
    }
    tc05::fence_before_sync();
  }
  if (!(amax < kF16Max)) atomicOr(status, RRNCO_DEV_NAN_LOGITS);  // fp16 operand overflow or NaN input: loud
  tc05::fence_before_sync();
  compute_sync();
  if (warp == 0) tc05::tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// C[512, 128] += X^T Y over `M` rows.  K = rows: both operands are transposed on the way into shared memory (lane = feature:
// coalesced 128-byte row segments from global, one 16-byte K chunk of 8 rows per store), so the MMAs see the same K-major
// core-matrix layout as everywhere else.  One CTA = one half of X's 512 features (256 TMEM columns, 96 KB of shared memory:
// two CTAs per SM), persistent over 32-row tiles with a two-stage operand buffer; 8 conversion warps + one issuing warp.
constexpr int kXRows = 32;
constexpr int kXFeat = 256;
constexpr int kXThreads = 288;
struct XtySmem {
  uint16_t x_hi[2][kXFeat * kXRows];  // [128-feature block (2)][8-row K chunk (4)][feature (128)][8 halves]
  uint16_t x_lo[2][kXFeat * kXRows];
  uint16_t y_hi[2][kE * kXRows];      // [K chunk (4)][column (128)][8 halves]
  uint16_t y_lo[2][kE * kXRows];
  uint64_t bar_full[2];
  uint64_t bar_free[2];
  uint64_t bar_done;
  uint32_t tmem_base;
};
static_assert(sizeof(XtySmem) <= 113 * 1024, "two CTAs per SM");

__global__ void __launch_bounds__(kXThreads, 2) xty_kernel(int64_t M, const float* __restrict__ X, const float* __restrict__ Y,
                                                          const float* __restrict__ sx_ptr, const float* __restrict__ sy_ptr,
                                                          float* __restrict__ C, float* __restrict__ xsum, float* __restrict__ ysum,
                                                          uint32_t* __restrict__ status, int x_tiled) {
  extern __shared__ __align__(128) unsigned char xty_smem_raw[];
  XtySmem& sm = *reinterpret_cast<XtySmem*>(xty_smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = blockIdx.y;
  const int64_t n_tiles = (M + kXRows - 1) / kXRows;
  if (warp == 0) tc05::tmem_alloc(&sm.tmem_base, 256);
  if (tid == 32) {
    for (int i = 0; i < 2; ++i) {
      tc05::mbar_init(&sm.bar_full[i], 256);
      tc05::mbar_init(&sm.bar_free[i], 1);
    }
    tc05::mbar_init(&sm.bar_done, 1);
    tc05::fence_mbar_init();
  }
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tb = sm.tmem_base;

  const int uwarp = __shfl_sync(0xffffffffu, warp, 0);
  if (uwarp == 8) {
    if (tc05::elect_one()) {
      const uint32_t idesc = tc05::make_idesc_f16(128, 128);
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const uint32_t st = it & 1u;
        tc05::mbar_wait(&sm.bar_full[st], (it >> 1) & 1u);
        tc05::fence_after_sync();
        const uint32_t xh = tc05::smem_u32(sm.x_hi[st]), xl = tc05::smem_u32(sm.x_lo[st]);
        const uint32_t yh = tc05::smem_u32(sm.y_hi[st]), yl = tc05::smem_u32(sm.y_lo[st]);
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
#pragma unroll
          for (int ks = 0; ks < kXRows / 16; ++ks) {
            const uint32_t ao = mb * (4 * kFRows * 16) + ks * 2 * kLboTile, bo = ks * 2 * kLboTile;
            const uint64_t a_hi = tc05::make_desc(xh + ao, kLboTile, kSbo), a_lo = tc05::make_desc(xl + ao, kLboTile, kSbo);
            const uint64_t b_hi = tc05::make_desc(yh + bo, kLboTile, kSbo), b_lo = tc05::make_desc(yl + bo, kLboTile, kSbo);
            tc05::mma_ss_f16(tb + mb * 128, a_hi, b_hi, idesc, (it > 0 || ks > 0) ? 1u : 0u);
            tc05::mma_ss_f16(tb + mb * 128, a_lo, b_hi, idesc, 1u);
            tc05::mma_ss_f16(tb + mb * 128, a_hi, b_lo, idesc, 1u);
          }
        }
        tc05::commit(&sm.bar_free[st]);
      }
      tc05::commit(&sm.bar_done);
    }
    return;
  }

  const float sx = sx_ptr ? __ldg(sx_ptr) : kAScale, sy = sy_ptr ? __ldg(sy_ptr) : kAScale;
  const int f = tid;                         // X feature of this thread within the CTA's half
  const int fy = tid & 127, ry = tid >> 7;   // Y column, first 8-row group (second: ry + 2)
  const float* xcol = X + half * kXFeat + f;
  float csum = 0.f, ysm = 0.f, amax = 0.f;
  float csum4[4] = {0.f, 0.f, 0.f, 0.f};   // x_tiled: the four features of this thread's items
  uint32_t it = 0;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
    const uint32_t st = it & 1u;
    const int64_t r0 = t * kXRows;
    float xv[kXRows], yv[16];
    if (x_tiled) {
      // X = [tile of 128 rows][feature (512)][row (128)] (ffn_train_kernel's hidden_out): item = (feature, 8-row group), 32 contiguous
      // bytes; lane & 3 = group, so four lanes read one 128-byte line; item u of this thread: feature u * 64 + tid / 4
      const float* xt = X + (r0 >> 7) * (int64_t)(kF * kFRows) + (int64_t)(half * kXFeat) * kFRows + (r0 & 127);
      const int kc = tid & 3;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int fu = u * 64 + (tid >> 2);
        const float4* src = reinterpret_cast<const float4*>(xt + (int64_t)fu * kFRows + kc * 8);
        const int64_t rr = r0 + kc * 8;
        float4 a = __ldg(src), c4 = __ldg(src + 1);
        if (rr + 7 >= M) {  // ragged tail: rows beyond M hold nothing
          if (rr + 0 >= M) a.x = 0.f; if (rr + 1 >= M) a.y = 0.f; if (rr + 2 >= M) a.z = 0.f; if (rr + 3 >= M) a.w = 0.f;
          if (rr + 4 >= M) c4.x = 0.f; if (rr + 5 >= M) c4.y = 0.f; if (rr + 6 >= M) c4.z = 0.f; if (rr + 7 >= M) c4.w = 0.f;
        }
        xv[u * 8] = a.x; xv[u * 8 + 1] = a.y; xv[u * 8 + 2] = a.z; xv[u * 8 + 3] = a.w;
        xv[u * 8 + 4] = c4.x; xv[u * 8 + 5] = c4.y; xv[u * 8 + 6] = c4.z; xv[u * 8 + 7] = c4.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < kXRows; ++r) xv[r] = (r0 + r < M) ? __ldg(xcol + (r0 + r) * kF) : 0.f;
    }
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int64_t rr = r0 + (ry + 2 * g) * 8 + r;
        yv[g * 8 + r] = rr < M ? __ldg(Y + rr * kE + fy) : 0.f;
      }
    if (it >= 2) tc05::mbar_wait(&sm.bar_free[st], ((it >> 1) - 1) & 1u);  // the MMAs that read this stage have completed
#pragma unroll
    for (int u = 0; u < 4; ++u) {   // row-major X: u = 8-row group of feature f; tiled X: u = item (feature u * 64 + tid / 4, group tid & 3)
      uint32_t h[4], l[4];
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = xv[u * 8 + 2 * i], b = xv[u * 8 + 2 * i + 1];
        part += a + b;
        amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)) * sx);
        f16s_split2(a, b, sx, h[i], l[i]);
      }
      csum += part;
      csum4[u] += part;
      const int fe = x_tiled ? u * 64 + (tid >> 2) : f, kc = x_tiled ? (tid & 3) : u;
      const int dst = (fe >> 7) * (4 * kFRows * 8) + kc * (kFRows * 8) + (fe & 127) * 8;
      *reinterpret_cast<uint4*>(&sm.x_hi[st][dst]) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(&sm.x_lo[st][dst]) = make_uint4(l[0], l[1], l[2], l[3]);
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = yv[g * 8 + 2 * i], b = yv[g * 8 + 2 * i + 1];
        ysm += a + b;
        amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)) * sy);
        f16s_split2(a, b, sy, h[i], l[i]);
      }
      const int dst = (ry + 2 * g) * (kFRows * 8) + fy * 8;
      *reinterpret_cast<uint4*>(&sm.y_hi[st][dst]) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(&sm.y_lo[st][dst]) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    tc05::fence_proxy_async();
    tc05::mbar_arrive(&sm.bar_full[st]);
  }
  // ---- epilogue: every MMA has completed ----
  tc05::mbar_wait(&sm.bar_done, 0u);
  tc05::fence_after_sync();
  {
    const int mb = warp >> 2, m = (warp & 3) * 32 + lane;
    const uint32_t lane_b = (uint32_t)((warp & 3) * 32) << 16;
    const float unscale = 1.0f / (sx * sy);
    float* crow = C + (size_t)(half * kXFeat + mb * 128 + m) * kE;
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
      uint32_t v[16];
      tc05::tmem_ld16(tb + lane_b + mb * 128 + q * 16, v);
      tc05::tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) atomicAdd(crow + q * 16 + i, __uint_as_float(v[i]) * unscale);
    }
  }
  if (xsum != nullptr) {
    if (x_tiled) {
#pragma unroll
      for (int u = 0; u < 4; ++u) atomicAdd(xsum + half * kXFeat + u * 64 + (tid >> 2), csum4[u]);
    } else {
      atomicAdd(xsum + half * kXFeat + f, csum);
    }
  }
  if (ysum != nullptr && half == 0) atomicAdd(ysum + fy, ysm);
  if (!(amax < kF16Max)) atomicOr(status, RRNCO_DEV_NAN_LOGITS);
  tc05::fence_before_sync();
  asm volatile("bar.sync 1, 256;\n" ::: "memory");
  if (warp == 0) tc05::tmem_dealloc(tb, 256);
}

}  // namespace rrnco

using namespace rrnco;

extern "C" {

int64_t rrnco_train_ffn_packed_bytes(void) { return kFfnPackedBytes; }

int rrnco_train_ffn_pack(const float* wa, const float* wb, void* packed, uint32_t* status, void* stream) {
  RRNCO_CHECK_ARG(wa && wb && packed && status && (reinterpret_cast<uintptr_t>(packed) & 15u) == 0);
  const int n = kFSlices * kFRows * (kFSliceK / 2);
  pack_ffn_train_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(wa, wb, reinterpret_cast<uint32_t*>(packed), status);
  return rrnco_launch_status();
}

int rrnco_train_ffn(int32_t mode, int64_t rows, const float* x, const void* packed, const float* b1, const float* b2,
                    const float* a_scale, uint32_t* mask, float* hidden_out, float* y, uint32_t* status, void* stream) {
  if (rows == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG((mode == 0 || mode == 1) && rows > 0 && x && packed && status);
  RRNCO_CHECK_ARG(mode == 1 ? mask != nullptr : (b1 && b2));
  RRNCO_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(hidden_out) |
                    reinterpret_cast<uintptr_t>(y)) & 15u) == 0);
  static PerDeviceOnce once;
  if (once.first()) {
    if (cudaFuncSetAttribute(ffn_train_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FfnTrainSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(ffn_train_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FfnTrainSmem)) != cudaSuccess) {
      once.undo();
      return RRNCO_ERR_CUDA;
    }
  }
  const int64_t n_tiles = (rows + kFRows - 1) / kFRows;
  const int sms = device_sm_count();
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  const uint16_t* wp = reinterpret_cast<const uint16_t*>(packed);
  if (mode == 0)
    ffn_train_kernel<0><<<grid, kTThreads, sizeof(FfnTrainSmem), (cudaStream_t)stream>>>(rows, x, wp, b1, b2, a_scale, mask,
                                                                                      hidden_out, y, status);
  else
    ffn_train_kernel<1><<<grid, kTThreads, sizeof(FfnTrainSmem), (cudaStream_t)stream>>>(rows, x, wp, b1, b2, a_scale, mask,
                                                                                      hidden_out, y, status);
  return rrnco_launch_status();
}

int rrnco_train_xty(int64_t rows, const float* x, int32_t x_tiled, const float* y, const float* sx, const float* sy, float* c,
                    float* xsum, float* ysum, uint32_t* status, void* stream) {
  if (rows == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(rows > 0 && x && y && c && status && (reinterpret_cast<uintptr_t>(x) & 15u) == 0);
  static PerDeviceOnce once;
  if (once.first()) {
    if (cudaFuncSetAttribute(xty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(XtySmem)) != cudaSuccess ||
        cudaFuncSetAttribute(xty_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) {
      once.undo();
      return RRNCO_ERR_CUDA;
    }
  }
  const int64_t n_tiles = (rows + kXRows - 1) / kXRows;
  const int sms = device_sm_count();
  dim3 grid((unsigned)(n_tiles < sms ? n_tiles : sms), 2);
  xty_kernel<<<grid, kXThreads, sizeof(XtySmem), (cudaStream_t)stream>>>(rows, x, y, sx, sy, c, xsum, ysum, status, x_tiled);
  return rrnco_launch_status();
}

}  // extern "C"
