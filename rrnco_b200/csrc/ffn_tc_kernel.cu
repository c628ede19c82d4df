// Residual FFN of RRNet_PointerAttention (rrnco/models/decoder.py:272-277,296) on the 5th-gen tensor cores:
//     out = W2 relu(W1 g + b1) + b2 + g          g: [M,128] fp32, W1: [512,128], W2: [128,512]
// tcgen05.mma kind::f16 on the two-term fp16 operand split of ffn_pack.cuh (three MMAs per K step into one fp32
// accumulator in TMEM = fp32-faithful, like 3xTF32 at half the tensor time).  One CTA = 128 rows, 12 warps:
//   warps 0-7  epilogues (thread per row, two column halves)
//   warp  8    TMA producer: the 64 packed weight slices (12 KB each) through an 8-stage full/empty mbarrier ring
//   warps 9-11 MMA issue, one warp per split term (a single thread sustains only ~1 tcgen05.mma per 160 cycles)
// Tensor-pipe order  G1(0) G1(1) G2(0) G1(2) G2(1) G1(3) G2(2) G2(3): GEMM1 alternates between two TMEM accumulators so
// that epilogue 1 of chunk c (+b1, relu, split -> fp16 A operand of GEMM2 in TMEM) runs under GEMM1 of chunk c+1.
// This is the standalone form of the phase the fused rollout kernel embeds; it is exported so that it can be
// validated on its own (tests/test_gpu_parity.py::test_pointer_ffn_tcgen05).
#include "common.cuh"
#include "tc05.cuh"
#include "ffn_pack.cuh"

namespace rrnco {

constexpr int kFThreads = 384;
constexpr float kUnscale = 1.0f / (kAScale * kWScale);

__device__ long long g_ffn_dbg[32];
__device__ int g_ffn_mode = 0;  // debug: 1 = producer streams nothing and the MMAs do not wait for weights (issue / tensor-rate probe)
#define FFN_STAMP(i) do { if (blockIdx.x == 0 && tid == 0) g_ffn_dbg[i] = clock64(); } while (0)

struct FfnSmem {
  uint16_t g_hi[kFRows * kE];  // [16-byte K chunk (16)][row (128)][8 halves]
  uint16_t g_lo[kFRows * kE];
  uint16_t w[kFStages][kFSliceHalves];
  float b1[kF];
  float b2[kE];
  uint64_t bar_full[kFStages];
  uint64_t bar_empty[kFStages];
  uint64_t bar_h[2];   // GEMM1 into accumulator b complete (3 issuer commits)
  uint64_t bar_epi;    // epilogue 1 of a chunk complete: A operand of GEMM2 written, accumulator re-zeroed (256 arrivals)
  uint64_t bar_g2;     // GEMM2 of a chunk complete: its A operand may be overwritten (3 issuer commits)
  uint32_t tmem_base;
};

// Pack W1 / W2 into the streaming order of the kernel (ffn_pack.cuh): one thread per (slice, row, k pair)
__global__ void pack_ffn_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                        uint32_t* __restrict__ packed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kFSlices * kFRows * (kFSliceK / 2)) return;
  constexpr int kPairs = kFSliceK / 2;  // k pairs per row of a slice
  const int kp = i % kPairs, row = (i / kPairs) & 127, s = i / (kPairs * kFRows);
  const int j = s / kFSlicesPerJob, ks = s % kFSlicesPerJob, c = ffn_job_chunk(j), half = ffn_job_half(j);
  const int k = ks * kFSliceK + kp * 2;
  float v0, v1;
  if (half == 0) {
    v0 = w1[(size_t)(c * kFRows + row) * kE + k];
    v1 = w1[(size_t)(c * kFRows + row) * kE + k + 1];
  } else {
    v0 = w2[(size_t)row * kF + c * kFRows + k];
    v1 = w2[(size_t)row * kF + c * kFRows + k + 1];
  }
  uint32_t hi, lo;
  f16s_split2(v0, v1, kWScale, hi, lo);
  uint32_t* dst = packed + (size_t)s * (kFSliceHalves / 2) + (kp >> 2) * (kFRows * 4) + row * 4 + (kp & 3);
  dst[0] = hi;
  dst[kFVariantHalves / 2] = lo;
}

__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

__global__ void __launch_bounds__(kFThreads, 1) pointer_ffn_tc_kernel(int64_t M, const float* __restrict__ g_in,
                                                                     const uint16_t* __restrict__ wpacked,
                                                                     const float* __restrict__ b1,
                                                                     const float* __restrict__ b2,
                                                                     float* __restrict__ g_out) {
  extern __shared__ __align__(128) unsigned char ffn_smem_raw[];
  FfnSmem& sm = *reinterpret_cast<FfnSmem*>(ffn_smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * kFRows;
  const int mode = g_ffn_mode;

  FFN_STAMP(0);
  if (warp == 0) tc05::tmem_alloc(&sm.tmem_base, 512);
  if (tid == 32) {
    for (int i = 0; i < kFStages; ++i) {
      tc05::mbar_init(&sm.bar_full[i], 1);
      tc05::mbar_init(&sm.bar_empty[i], 3);
    }
    tc05::mbar_init(&sm.bar_h[0], 3);
    tc05::mbar_init(&sm.bar_h[1], 3);
    tc05::mbar_init(&sm.bar_epi, 256);
    tc05::mbar_init(&sm.bar_g2, 3);
    tc05::fence_mbar_init();
  }
  if (tid < 256) {
    for (int i = tid; i < kF; i += 256) sm.b1[i] = b1[i];
    if (tid < kE) sm.b2[tid] = b2[tid];
    // g -> A_hi | A_lo fp16 tiles in the core-matrix layout: item = (row, 8-wide K chunk)
    for (int idx = tid; idx < kFRows * 16; idx += 256) {
      const int row = idx & 127, c8 = idx >> 7;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (m0 + row < M) {
        v0 = __ldg(reinterpret_cast<const float4*>(g_in + (m0 + row) * kE) + c8 * 2);
        v1 = __ldg(reinterpret_cast<const float4*>(g_in + (m0 + row) * kE) + c8 * 2 + 1);
      }
      uint32_t h[4], l[4];
      f16s_split2(v0.x, v0.y, kAScale, h[0], l[0]); f16s_split2(v0.z, v0.w, kAScale, h[1], l[1]);
      f16s_split2(v1.x, v1.y, kAScale, h[2], l[2]); f16s_split2(v1.z, v1.w, kAScale, h[3], l[3]);
      const int dst = c8 * (kFRows * 8) + row * 8;
      *reinterpret_cast<uint4*>(&sm.g_hi[dst]) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(&sm.g_lo[dst]) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    tc05::fence_proxy_async();
  }
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int colhalf = (warp >> 2) & 1;  // warps 0-3: columns [0,64), warps 4-7: [64,128)
  __syncthreads();                      // tmem_base visible
  tc05::fence_after_sync();
  const uint32_t tbase = sm.tmem_base;
  const uint32_t t_hacc0 = tbase, t_oacc = tbase + 256, t_hhi = tbase + 384, t_hlo = tbase + 448;
  if (tid < 256) {  // zero the three accumulators: every MMA accumulates
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      tc05::tmem_st16(t_hacc0 + lane_base + colhalf * 64 + q * 16, z);
      tc05::tmem_st16(t_hacc0 + 128 + lane_base + colhalf * 64 + q * 16, z);
      tc05::tmem_st16(t_oacc + lane_base + colhalf * 64 + q * 16, z);
    }
    tc05::tmem_wait_st();
  }
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  FFN_STAMP(1);

  if (warp == 8) {
    // ===== TMA producer =====
    if (lane == 0 && mode != 1) {
      for (int s = 0; s < kFSlices; ++s) {
        const int st = s & (kFStages - 1);
        if (s >= kFStages && mode != 3) tc05::mbar_wait(&sm.bar_empty[st], ((s / kFStages) - 1) & 1);
        tc05::mbar_arrive_expect_tx(&sm.bar_full[st], kFSliceBytes);
        tc05::bulk_g2s(sm.w[st], wpacked + (size_t)s * kFSliceHalves, kFSliceBytes, &sm.bar_full[st]);
        if (blockIdx.x == 0 && (s & 3) == 3) g_ffn_dbg[12 + (s >> 2)] = clock64();
      }
    }
  } else if (warp >= 9) {
    // ===== MMA issue: warp 9 + p issues split term p of every slice (0: A_hi B_hi, 1: A_lo B_hi, 2: A_hi B_lo); all
    // terms accumulate into the same pre-zeroed TMEM tile, so no issue order between the warps is needed =====
    if (lane == 0) {
      const int term = warp - 9;
      const uint32_t idesc = tc05::make_idesc_f16(128, 128);
      const uint32_t a_addr = tc05::smem_u32(term == 1 ? sm.g_lo : sm.g_hi);
      const uint32_t t_a = term == 1 ? t_hlo : t_hhi;
      uint32_t epi_phase = 0;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        const int c = ffn_job_chunk(j), half = ffn_job_half(j);
        if (half == 1) {  // A operand of GEMM2(c) written, and (implied, one job later) accumulator c & 1 re-zeroed
          tc05::mbar_wait(&sm.bar_epi, epi_phase);
          epi_phase ^= 1u;
          tc05::fence_after_sync();
        }
#pragma unroll 1
        for (int sj = 0; sj < kFSlicesPerJob; ++sj) {
          const int s = j * kFSlicesPerJob + sj, st = s & (kFStages - 1);
          if (mode == 0 || mode == 3) tc05::mbar_wait(&sm.bar_full[st], (s / kFStages) & 1);
          tc05::fence_after_sync();
          const uint32_t b_addr = tc05::smem_u32(sm.w[st]) + (term == 2 ? kFVariantHalves * 2 : 0);
#pragma unroll
          for (int kk = 0; kk < kFKSteps; ++kk) {
            const int ks = sj * kFKSteps + kk;  // K step (16 values) of the job
            const uint64_t bdesc = tc05::make_desc(b_addr + kk * 2 * kLboTile, kLboTile, kSbo);
            if (half == 0) {
              const uint64_t adesc = tc05::make_desc(a_addr + ks * 2 * kLboTile, kLboTile, kSbo);
              tc05::mma_ss_f16(t_hacc0 + (c & 1) * 128, adesc, bdesc, idesc, 1u);
            } else {
              tc05::mma_ts_f16(t_oacc, t_a + ks * 8, bdesc, idesc, 1u);
            }
          }
          tc05::commit(&sm.bar_empty[st]);
        }
        tc05::commit(half == 0 ? &sm.bar_h[c & 1] : &sm.bar_g2);
      }
    }
  } else {
    // ===== epilogues =====
    const int row = (warp & 3) * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      tc05::mbar_wait(&sm.bar_h[c & 1], (c >> 1) & 1, 32);
      if (c > 0) tc05::mbar_wait(&sm.bar_g2, (c - 1) & 1, 32);  // GEMM2(c-1) has consumed the previous A operand
      tc05::fence_after_sync();
      if (warp == 0) FFN_STAMP(2 + 2 * c);
      const uint32_t t_h = t_hacc0 + (c & 1) * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col0 = colhalf * 64 + q * 16;
        uint32_t v[16], hi[8], lo[8];
        tc05::tmem_ld16(t_h + lane_base + col0, v);
        tc05::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float h0 = fmaxf(fmaf(__uint_as_float(v[i]), kUnscale, sm.b1[c * kFRows + col0 + i]), 0.f);
          const float h1 = fmaxf(fmaf(__uint_as_float(v[i + 1]), kUnscale, sm.b1[c * kFRows + col0 + i + 1]), 0.f);
          f16s_split2(h0, h1, kAScale, hi[i >> 1], lo[i >> 1]);
        }
        tc05::tmem_st8(t_hhi + lane_base + (col0 >> 1), hi);
        tc05::tmem_st8(t_hlo + lane_base + (col0 >> 1), lo);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
        tc05::tmem_st16(t_h + lane_base + col0, v);  // re-zero the chunk accumulator for GEMM1(c + 2)
      }
      tc05::tmem_wait_st();
      tc05::fence_before_sync();
      tc05::mbar_arrive(&sm.bar_epi);
      if (warp == 0) FFN_STAMP(3 + 2 * c);
    }
    tc05::mbar_wait(&sm.bar_g2, 1, 32);  // GEMM2(3): output complete
    tc05::fence_after_sync();
    FFN_STAMP(10);
    // output epilogue: out = acc / (kAScale kWScale) + b2 + g
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col0 = colhalf * 64 + q * 16;
      uint32_t v[16];
      tc05::tmem_ld16(t_oacc + lane_base + col0, v);
      tc05::tmem_wait_ld();
      if (m0 + row < M) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(g_in + (m0 + row) * kE + col0 + i));
          float4 o;
          o.x = fmaf(__uint_as_float(v[i]), kUnscale, sm.b2[col0 + i]) + g.x;
          o.y = fmaf(__uint_as_float(v[i + 1]), kUnscale, sm.b2[col0 + i + 1]) + g.y;
          o.z = fmaf(__uint_as_float(v[i + 2]), kUnscale, sm.b2[col0 + i + 2]) + g.z;
          o.w = fmaf(__uint_as_float(v[i + 3]), kUnscale, sm.b2[col0 + i + 3]) + g.w;
          *reinterpret_cast<float4*>(g_out + (m0 + row) * kE + col0 + i) = o;
        }
      }
    }
    tc05::fence_before_sync();
  }
  __syncthreads();
  FFN_STAMP(26);
  if (warp == 0) tc05::tmem_dealloc(tbase, 512);
}

int pack_ffn_weights(const float* w1, const float* w2, void* packed, cudaStream_t st) {
  const int n = kFSlices * kFRows * (kFSliceK / 2);
  pack_ffn_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(w1, w2, reinterpret_cast<uint32_t*>(packed));
  return cudaGetLastError() == cudaSuccess ? RRNCO_OK : RRNCO_ERR_CUDA;
}

}  // namespace rrnco

using namespace rrnco;

extern "C" {

int rrnco_debug_ffn_mode(int mode) {
  return cudaMemcpyToSymbol(g_ffn_mode, &mode, sizeof(int)) == cudaSuccess ? RRNCO_OK : RRNCO_ERR_CUDA;
}

// debug: clock64 stamps of CTA 0 of the last pointer_ffn_tc_kernel launch (host buffer of 32 int64)
int rrnco_debug_ffn_stamps(long long* h_out) {
  return cudaMemcpyFromSymbol(h_out, g_ffn_dbg, sizeof(long long) * 32) == cudaSuccess ? RRNCO_OK : RRNCO_ERR_CUDA;
}

int64_t rrnco_pointer_ffn_workspace_bytes(void) { return kFfnPackedBytes; }

int rrnco_pointer_ffn(int64_t n_rows, const float* g_in, const float* w1, const float* b1, const float* w2,
                      const float* b2, float* g_out, void* workspace, void* stream) {
  if (n_rows == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(n_rows > 0 && g_in && w1 && b1 && w2 && b2 && g_out && workspace);
  RRNCO_CHECK_ARG(((uintptr_t)g_in & 15) == 0 && ((uintptr_t)g_out & 15) == 0 && ((uintptr_t)workspace & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = pack_ffn_weights(w1, w2, workspace, st);
  if (rc != RRNCO_OK) return rc;
  static PerDeviceOnce once;
  if (once.first()) {
    if (cudaFuncSetAttribute(pointer_ffn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FfnSmem)) !=
        cudaSuccess) {
      once.undo();
      return RRNCO_ERR_CUDA;
    }
  }
  pointer_ffn_tc_kernel<<<(unsigned)((n_rows + kFRows - 1) / kFRows), kFThreads, sizeof(FfnSmem), st>>>(
      n_rows, g_in, reinterpret_cast<const uint16_t*>(workspace), b1, b2, g_out);
  return rrnco_launch_status();
}

}  // extern "C"
