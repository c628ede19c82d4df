// Residual FFN of RRNet_PointerAttention (rrnco/models/decoder.py:272-277,296) on the 5th-gen tensor cores:
//     out = W2 relu(W1 g + b1) + b2 + g          g: [M,128] fp32, W1: [512,128], W2: [128,512]
// tcgen05.mma kind::tf32 with the 3xTF32 error-compensated split (a*b ~ ah*bh + al*bh + ah*bl), fp32
// accumulators in TMEM.  One CTA = 128 rows:
//   GEMM1 (per 128-wide hidden chunk): A = g (hi | lo) in shared memory, B = pre-packed W1 slices streamed from L2
//          by a TMA producer warp (cp.async.bulk + mbarrier full/empty ring, 4 stages x 16 KB)
//   epilogue 1: TMEM -> registers (thread per row), + b1, relu, split -> TMEM as the A operand of GEMM2
//   GEMM2: A = hidden chunk (hi | lo) in TMEM, B = W2 slices, accumulating the 128-wide output across chunks
// This is the standalone form of the phase the fused rollout kernel embeds; it is exported so that it can be
// validated on its own (tests/test_gpu_parity.py::test_pointer_ffn_tcgen05).
#include "common.cuh"
#include "tc05.cuh"
#include "ffn_pack.cuh"

namespace rrnco {

constexpr int kFThreads = 288;
constexpr int kFIssuers = 3;                        // MMA-issuing warps (one per 3xTF32 pass): a single thread sustains
                                                    // only ~1 tcgen05.mma per 160 cycles, the tensor pipe needs one per 64
constexpr int kFReplicas = 1;                       // copies of the packed weights: CTAs spread over them so that the
                                                    // whole chip does not hammer the same L2 lines in lock-step                      // 8 compute warps + 1 TMA producer warp

__device__ long long g_ffn_dbg[32];
__device__ int g_ffn_mode = 0;  // debug: 1 = MMAs do not wait for weights, 2 = weights streamed but no MMAs, 3 = bf16 probe,
                                // 4 = no per-slice fence+commit, 5 = no per-slice commit, 6 = no per-slice fence
#define FFN_STAMP(i) do { if (blockIdx.x == 0 && tid == 0) g_ffn_dbg[i] = clock64(); } while (0)

struct FfnSmem {
  float g_hi[kFRows * kE];
  float g_lo[kFRows * kE];
  float w[kFStages][kFSliceFloats];  // [stage][hi | lo] packed slices
  float b1[kF];
  float b2[kE];
  uint64_t bar_full[kFStages];
  uint64_t bar_empty[kFStages];
  uint64_t bar_acc;
  uint32_t tmem_base;
};

// Pack W1 / W2 into the streaming order of the kernel: slice s = (chunk c, half, ks), each slice a contiguous
// 16 KB block [hi | lo][16-byte K chunk c4][row][4 floats] = the shared-memory core-matrix layout, so that one
// cp.async.bulk moves a whole slice.
__global__ void pack_ffn_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                        float* __restrict__ packed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (slice, c4, row, e)
  if (i >= kFSlices * kFRows * kFSliceK) return;
  const int e = i & 3, row = (i >> 2) & 127, c4 = (i >> 9) & 3, s = i >> 11;
  const int c = s >> 4, half = (s >> 3) & 1, ks = s & 7;
  const int k = ks * kFSliceK + c4 * 4 + e;
  const float v = half == 0 ? w1[(size_t)(c * kFRows + row) * kE + k] : w2[(size_t)row * kF + c * kFRows + k];
  uint32_t h, l;
  split_tf32(v, h, l);
  float* dst = packed + (size_t)s * kFSliceFloats + c4 * (kFRows * 4) + row * 4 + e;
  dst[0] = __uint_as_float(h);
  dst[kFRows * kFSliceK] = __uint_as_float(l);
}

__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

__global__ void __launch_bounds__(kFThreads, 1) pointer_ffn_tc_kernel(int64_t M, const float* __restrict__ g_in,
                                                                     const float* __restrict__ wpacked,
                                                                     const float* __restrict__ b1,
                                                                     const float* __restrict__ b2,
                                                                     float* __restrict__ g_out) {
  extern __shared__ __align__(128) unsigned char ffn_smem_raw[];
  FfnSmem& sm = *reinterpret_cast<FfnSmem*>(ffn_smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * kFRows;

  FFN_STAMP(0);
  if (warp == 0) tc05::tmem_alloc(&sm.tmem_base, 512);
  if (tid == 32) {
    for (int i = 0; i < kFStages; ++i) {
      tc05::mbar_init(&sm.bar_full[i], 1);
      tc05::mbar_init(&sm.bar_empty[i], kFIssuers);
    }
    tc05::mbar_init(&sm.bar_acc, kFIssuers);
    tc05::fence_mbar_init();
  }
  if (tid < 256) {
    for (int i = tid; i < kF; i += 256) sm.b1[i] = b1[i];
    if (tid < kE) sm.b2[tid] = b2[tid];
    // g -> hi | lo in the core-matrix layout (row r, 16-byte chunk c4)
    for (int idx = tid; idx < kFRows * 32; idx += 256) {
      const int row = idx >> 5, c4 = idx & 31;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + row < M) v = __ldg(reinterpret_cast<const float4*>(g_in + (m0 + row) * kE) + c4);
      uint32_t h[4], l[4];
      split_tf32(v.x, h[0], l[0]); split_tf32(v.y, h[1], l[1]); split_tf32(v.z, h[2], l[2]); split_tf32(v.w, h[3], l[3]);
      const int dst = c4 * (kFRows * 4) + row * 4;
      *reinterpret_cast<uint4*>(&sm.g_hi[dst]) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(&sm.g_lo[dst]) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    tc05::fence_proxy_async();
  }
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  FFN_STAMP(1);

  if (warp == 8) {
    // ===== TMA producer: streams the 64 packed weight slices through the 4-stage ring =====
    if (lane == 0 && (g_ffn_mode == 0 || g_ffn_mode == 2)) {  // (timing probes 1,3-9 do not stream weights)
      for (int s = 0; s < kFSlices; ++s) {
        const int st = s & (kFStages - 1);
        if (s >= kFStages) tc05::mbar_wait(&sm.bar_empty[st], ((s / kFStages) - 1) & 1);
        tc05::mbar_arrive_expect_tx(&sm.bar_full[st], kFSliceBytes);
        const float* src = wpacked + (size_t)(blockIdx.x % kFReplicas) * kFSlices * kFSliceFloats + (size_t)s * kFSliceFloats;
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // four 4 KB pieces, start piece rotated per CTA
          const int piece = (q + blockIdx.x / kFReplicas) & 3;
          tc05::bulk_g2s(sm.w[st] + piece * (kFSliceFloats / 4), src + piece * (kFSliceFloats / 4), kFSliceBytes / 4,
                         &sm.bar_full[st]);
        }
      }
    }
    return;
  }

  const uint32_t tbase = sm.tmem_base;
  const uint32_t t_hacc = tbase, t_ahi = tbase + 128, t_alo = tbase + 256, t_out = tbase + 384;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int colhalf = warp >> 2;  // warps 0-3: columns [0,64), warps 4-7: [64,128)
  const uint32_t idesc = tc05::make_idesc_tf32(128, 128);
  const uint32_t g_hi_addr = tc05::smem_u32(sm.g_hi), g_lo_addr = tc05::smem_u32(sm.g_lo);

  {  // zero both accumulators: every MMA accumulates
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      tc05::tmem_st16(t_hacc + lane_base + colhalf * 64 + q * 16, z);
      tc05::tmem_st16(t_out + lane_base + colhalf * 64 + q * 16, z);
    }
    tc05::tmem_wait_st();
    tc05::fence_before_sync();
    compute_bar_sync();
    tc05::fence_after_sync();
  }
  int s = 0;
  uint32_t acc_phase = 0;
  const int mode = g_ffn_mode;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      if (mode == 9) {
        if (warp < 2) {
          for (int ks = 0; ks < 8; ++ks) {
            if (lane == 0) {
              const uint32_t whi = tc05::smem_u32(sm.w[ks & 3]), wlo = whi + kFRows * kFSliceK * 4;
              const int kk = warp;
              const uint64_t bh = tc05::make_desc(whi + kk * 2 * kLboTile, kLboTile, kSbo);
              const uint64_t bl = tc05::make_desc(wlo + kk * 2 * kLboTile, kLboTile, kSbo);
              const uint32_t koff = (ks * 4 + kk * 2) * kLboTile;
              const uint64_t ah = tc05::make_desc(g_hi_addr + koff, kLboTile, kSbo);
              const uint32_t d = warp == 0 ? t_hacc : t_out;
              tc05::mma_ss(d, ah, bh, idesc, 1u); tc05::mma_ss(d, ah, bl, idesc, 1u); tc05::mma_ss(d, ah, bh, idesc, 1u);
              if (ks == 7 && warp == 1) tc05::commit(&sm.bar_acc);
            }
            __syncwarp();
          }
        }
      } else if (warp < kFIssuers) {
        // ===== MMA issue: warp w issues pass w (0: lo*hi, 1: hi*lo, 2: hi*hi) of every K step; all passes
        // accumulate into the same (pre-zeroed) TMEM tile, so no issue order between the warps is needed =====
#pragma unroll 1
        for (int ks = 0; ks < 8; ++ks) {
          const int sg = s + ks;
          const int st = sg & (kFStages - 1);
          if (mode == 0) tc05::mbar_wait(&sm.bar_full[st], (sg / kFStages) & 1);
          tc05::fence_after_sync();
          if (lane == 0) {
            const uint32_t whi = tc05::smem_u32(sm.w[st]), wlo = whi + kFRows * kFSliceK * 4;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t bdesc = tc05::make_desc((warp == 1 ? wlo : whi) + kk * 2 * kLboTile, kLboTile, kSbo);
              if (half == 0) {
                const uint32_t koff = (ks * 4 + kk * 2) * kLboTile;  // 16-byte K chunk index * LBO
                const uint64_t adesc = tc05::make_desc((warp == 0 ? g_lo_addr : g_hi_addr) + koff, kLboTile, kSbo);
                tc05::mma_ss(t_hacc, adesc, bdesc, idesc, 1u);
              } else {
                const uint32_t kcol = ks * kFSliceK + kk * 8;  // hidden unit (TMEM column) of this K step
                tc05::mma_ts(t_out, (warp == 0 ? t_alo : t_ahi) + kcol, bdesc, idesc, 1u);
              }
            }
            tc05::commit(&sm.bar_empty[st]);
            if (ks == 7) tc05::commit(&sm.bar_acc);
          }
          __syncwarp();
        }
      }
      s += 8;
      if (warp == 0) FFN_STAMP(2 + 3 * (c * 2 + half));
      tc05::mbar_wait(&sm.bar_acc, acc_phase);
      acc_phase ^= 1u;
      tc05::fence_after_sync();
      FFN_STAMP(3 + 3 * (c * 2 + half));
      if (half == 0) {
        // hidden chunk ready in TMEM: + b1, relu, split -> A operand (hi | lo) of GEMM2
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col0 = colhalf * 64 + q * 16;
          uint32_t v[16], hi[16], lo[16];
          tc05::tmem_ld16(t_hacc + lane_base + col0, v);
          tc05::tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float h = fmaxf(__uint_as_float(v[i]) + sm.b1[c * kFRows + col0 + i], 0.f);
            split_tf32(h, hi[i], lo[i]);
          }
          tc05::tmem_st16(t_ahi + lane_base + col0, hi);
          tc05::tmem_st16(t_alo + lane_base + col0, lo);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
          tc05::tmem_st16(t_hacc + lane_base + col0, v);  // re-zero the chunk accumulator for the next GEMM1
        }
        tc05::tmem_wait_st();
        tc05::fence_before_sync();
        compute_bar_sync();
        FFN_STAMP(4 + 3 * (c * 2 + half));
      }
      // half == 1: GEMM2 of chunk c is complete here, so epilogue 1 of chunk c+1 may overwrite its A operand
    }
  }
  // output epilogue: out = acc + b2 + g
  const int row = (warp & 3) * 32 + lane;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col0 = colhalf * 64 + q * 16;
    uint32_t v[16];
    tc05::tmem_ld16(t_out + lane_base + col0, v);
    tc05::tmem_wait_ld();
    if (m0 + row < M) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(g_in + (m0 + row) * kE + col0 + i));
        float4 o;
        o.x = __uint_as_float(v[i]) + sm.b2[col0 + i] + g.x;
        o.y = __uint_as_float(v[i + 1]) + sm.b2[col0 + i + 1] + g.y;
        o.z = __uint_as_float(v[i + 2]) + sm.b2[col0 + i + 2] + g.z;
        o.w = __uint_as_float(v[i + 3]) + sm.b2[col0 + i + 3] + g.w;
        *reinterpret_cast<float4*>(g_out + (m0 + row) * kE + col0 + i) = o;
      }
    }
  }
  tc05::fence_before_sync();
  compute_bar_sync();
  FFN_STAMP(26);
  if (warp == 0) tc05::tmem_dealloc(tbase, 512);
}

int pack_ffn_weights(const float* w1, const float* w2, float* packed, cudaStream_t st) {
  const int n = kFSlices * kFRows * kFSliceK;
  pack_ffn_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(w1, w2, packed);
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

int64_t rrnco_pointer_ffn_workspace_bytes(void) {
  return (int64_t)kFReplicas * kFSlices * kFSliceFloats * (int64_t)sizeof(float);
}

int rrnco_pointer_ffn(int64_t n_rows, const float* g_in, const float* w1, const float* b1, const float* w2,
                      const float* b2, float* g_out, void* workspace, void* stream) {
  if (n_rows == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(n_rows > 0 && g_in && w1 && b1 && w2 && b2 && g_out && workspace);
  RRNCO_CHECK_ARG(((uintptr_t)g_in & 15) == 0 && ((uintptr_t)g_out & 15) == 0 && ((uintptr_t)workspace & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  float* packed = reinterpret_cast<float*>(workspace);
  for (int r = 0; r < kFReplicas; ++r) {
    int rc = pack_ffn_weights(w1, w2, packed + (size_t)r * kFSlices * kFSliceFloats, st);
    if (rc != RRNCO_OK) return rc;
  }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(pointer_ffn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FfnSmem)) !=
        cudaSuccess)
      return RRNCO_ERR_CUDA;
    configured = true;
  }
  pointer_ffn_tc_kernel<<<(unsigned)((n_rows + kFRows - 1) / kFRows), kFThreads, sizeof(FfnSmem), st>>>(
      n_rows, g_in, packed, b1, b2, g_out);
  return rrnco_launch_status();
}

}  // extern "C"
