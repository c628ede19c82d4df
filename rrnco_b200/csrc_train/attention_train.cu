// Training hand-off, masked 8-head attention of RRNet_PointerAttention (rrnco/models/decoder.py:281-293) over the rows of the
// batched replay: one query row per (rollout, decode step), keys / values of the row's instance (rrnco_b200/training.py).
//
// Head dim 16 and ~100 keys are far from the tile shapes of library attention kernels (fp32 SDPA: 131 ms of the 390 ms
// training step, profiles/r1_train_step_probe.txt); the arithmetic itself is small (~3 kFMA per row, head and pass), so
// both directions run on the CUDA cores with the instance's K / V tiles resident in shared memory:
//   warp = head, lane = row of a 32-row group: every K / V read is one broadcast LDS.128, the action mask is 4 bit words.
//   forward:  online softmax in the log2 domain with a lazy rescale (only when a score exceeds the running max by 2^8).
//   backward: p and ds are recomputed from the saved log-sum-exp (flash-attention identity D = dO . O); dq stays in
//             registers; dK / dV need sums over ROWS: each warp hands its 32 p / ds values of a key to itself through shared
//             memory, lane (d, half) multiplies them with the 16 dO / q values of its half it keeps in registers, one
//             shuffle joins the halves, and the result is added to the CTA's dK / dV tile in shared memory (the head's
//             16 columns belong to this warp alone: no atomics); one atomic pass per CTA adds the tile to global memory.
#include "../csrc/common.cuh"
#include "../../include/rrnco_b200_train.h"

namespace rrnco {

constexpr int kAtThreads = 256;
constexpr int kAtGroup = 32;      // rows per group (lane = row)
constexpr int kMbStride = 5;      // mask words per row in shared memory (4 + 1 pad)
constexpr float kQScale = 0.25f * 1.4426950408889634f;  // 1 / sqrt(16) and log2(e): scores in the log2 domain

// mask bytes of the 32 rows [g0, g0 + 32) -> 4 bit words per row in shared memory (thread = (row, 16-key segment))
__device__ __forceinline__ void stage_mask_bits(uint32_t* mb, const uint8_t* __restrict__ mask, int64_t g0, int64_t row_hi, int N,
                                                int tid) {
  const int r = tid >> 3, p = tid & 7;
  const int64_t row = g0 + r;
  uint32_t bits = 0u;
  if (row < row_hi) {
    const uint8_t* src = mask + row * N + p * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (p * 16 + i < N && __ldg(src + i) != 0) bits |= 1u << i;
  }
  const uint32_t other = __shfl_xor_sync(0xffffffffu, bits, 1);
  if ((p & 1) == 0) mb[r * kMbStride + (p >> 1)] = bits | (other << 16);
}

__global__ void __launch_bounds__(kAtThreads, 2) attn_fwd_kernel(int64_t L, int rows_per_cta, int N, const float* __restrict__ q,
                                                                const float* __restrict__ k, const float* __restrict__ v,
                                                                const uint8_t* __restrict__ mask, int add_res,
                                                                float* __restrict__ out, float* __restrict__ lse2) {
  extern __shared__ __align__(16) float at_smem[];
  float* Ks = at_smem;
  float* Vs = Ks + N * kE;
  uint32_t* mb = reinterpret_cast<uint32_t*>(Vs + N * kE);
  const int tid = threadIdx.x, h = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.y;
  {
    const float4* k4 = reinterpret_cast<const float4*>(k + b * N * kE);
    const float4* v4 = reinterpret_cast<const float4*>(v + b * N * kE);
    for (int i = tid; i < N * (kE / 4); i += kAtThreads) {
      reinterpret_cast<float4*>(Ks)[i] = __ldg(k4 + i);
      reinterpret_cast<float4*>(Vs)[i] = __ldg(v4 + i);
    }
  }
  const int64_t row_lo = b * L + (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row_hi = min(row_lo + rows_per_cta, (b + 1) * L);
  for (int64_t g0 = row_lo; g0 < row_hi; g0 += kAtGroup) {
    __syncthreads();
    stage_mask_bits(mb, mask, g0, row_hi, N, tid);
    __syncthreads();
    const int64_t row = g0 + lane;
    const bool valid = row < row_hi;
    float qr[16], qs[16], o[16];
    {
      const float4* src = reinterpret_cast<const float4*>(q + row * kE + h * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = valid ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        qr[4 * i] = t.x; qr[4 * i + 1] = t.y; qr[4 * i + 2] = t.z; qr[4 * i + 3] = t.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      qs[i] = qr[i] * kQScale;
      o[i] = 0.f;
    }
    float m = -1e30f, l = 0.f;
#pragma unroll 1
    for (int wi = 0; wi < 4; ++wi) {
      const uint32_t word = mb[lane * kMbStride + wi];
      const int jn = min(32, N - wi * 32);
#pragma unroll 1
      for (int jj = 0; jj < jn; jj += 2) {  // two keys per iteration (two independent chains)
        const bool bit0 = (word >> jj) & 1u, bit1 = jj + 1 < jn && ((word >> (jj + 1)) & 1u);
        if (!__any_sync(0xffffffffu, bit0 || bit1)) continue;
        const int j0 = wi * 32 + jj, j1 = min(j0 + 1, N - 1);
        const float4* kp0 = reinterpret_cast<const float4*>(Ks + j0 * kE + h * 16);
        const float4* kp1 = reinterpret_cast<const float4*>(Ks + j1 * kE + h * 16);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = kp0[i], t1 = kp1[i];
          s0 = fmaf(qs[4 * i], t.x, s0); s0 = fmaf(qs[4 * i + 1], t.y, s0); s0 = fmaf(qs[4 * i + 2], t.z, s0); s0 = fmaf(qs[4 * i + 3], t.w, s0);
          s1 = fmaf(qs[4 * i], t1.x, s1); s1 = fmaf(qs[4 * i + 1], t1.y, s1); s1 = fmaf(qs[4 * i + 2], t1.z, s1); s1 = fmaf(qs[4 * i + 3], t1.w, s1);
        }
        const float smax = fmaxf(bit0 ? s0 : -INFINITY, bit1 ? s1 : -INFINITY);
        if (smax > m + 8.f) {  // lazy rescale: rare after the first feasible key
          const float corr = exp2f(m - smax);
          l *= corr;
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] *= corr;
          m = smax;
        }
        const float p0 = bit0 ? exp2f(s0 - m) : 0.f, p1 = bit1 ? exp2f(s1 - m) : 0.f;
        l += p0 + p1;
        const float4* vp0 = reinterpret_cast<const float4*>(Vs + j0 * kE + h * 16);
        const float4* vp1 = reinterpret_cast<const float4*>(Vs + j1 * kE + h * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = vp0[i], t1 = vp1[i];
          o[4 * i] = fmaf(p0, t.x, o[4 * i]); o[4 * i + 1] = fmaf(p0, t.y, o[4 * i + 1]);
          o[4 * i + 2] = fmaf(p0, t.z, o[4 * i + 2]); o[4 * i + 3] = fmaf(p0, t.w, o[4 * i + 3]);
          o[4 * i] = fmaf(p1, t1.x, o[4 * i]); o[4 * i + 1] = fmaf(p1, t1.y, o[4 * i + 1]);
          o[4 * i + 2] = fmaf(p1, t1.z, o[4 * i + 2]); o[4 * i + 3] = fmaf(p1, t1.w, o[4 * i + 3]);
        }
      }
    }
    if (valid) {
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      float4* dst = reinterpret_cast<float4*>(out + row * kE + h * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 t = make_float4(o[4 * i] * inv, o[4 * i + 1] * inv, o[4 * i + 2] * inv, o[4 * i + 3] * inv);
        if (add_res) { t.x += qr[4 * i]; t.y += qr[4 * i + 1]; t.z += qr[4 * i + 2]; t.w += qr[4 * i + 3]; }
        dst[i] = t;
      }
      lse2[row * kH + h] = l > 0.f ? m + log2f(l) : 1e30f;
    }
  }
}

__global__ void __launch_bounds__(kAtThreads, 1) attn_bwd_kernel(int64_t L, int rows_per_cta, int N, const float* __restrict__ q,
                                                                const float* __restrict__ k, const float* __restrict__ v,
                                                                const uint8_t* __restrict__ mask, const float* __restrict__ out,
                                                                int add_res, const float* __restrict__ lse2,
                                                                const float* __restrict__ d_out, float* __restrict__ dq,
                                                                float* __restrict__ dk, float* __restrict__ dv) {
  extern __shared__ __align__(16) float at_smem[];
  float* Ks = at_smem;
  float* Vs = Ks + N * kE;
  float* dKs = Vs + N * kE;
  float* dVs = dKs + N * kE;
  float* ps = dVs + N * kE;            // [8 warps][2 buffers][2 keys][32 rows]
  float* dss = ps + 8 * 2 * 64;
  uint32_t* mb = reinterpret_cast<uint32_t*>(dss + 8 * 2 * 64);
  const int tid = threadIdx.x, h = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.y;
  {
    const float4* k4 = reinterpret_cast<const float4*>(k + b * N * kE);
    const float4* v4 = reinterpret_cast<const float4*>(v + b * N * kE);
    for (int i = tid; i < N * (kE / 4); i += kAtThreads) {
      reinterpret_cast<float4*>(Ks)[i] = __ldg(k4 + i);
      reinterpret_cast<float4*>(Vs)[i] = __ldg(v4 + i);
      reinterpret_cast<float4*>(dKs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(dVs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float* pw = ps + h * 128;
  float* dw = dss + h * 128;
  const int d = lane & 15, half = lane >> 4;
  const int64_t row_lo = b * L + (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row_hi = min(row_lo + rows_per_cta, (b + 1) * L);
  for (int64_t g0 = row_lo; g0 < row_hi; g0 += kAtGroup) {
    __syncthreads();
    stage_mask_bits(mb, mask, g0, row_hi, N, tid);
    __syncthreads();
    const int64_t row = g0 + lane;
    const bool valid = row < row_hi;
    // ---- this lane's row: scaled q, dO, D = dO . attn, log-sum-exp ----
    float qs[16], go[16], gq[16];
    float D = 0.f;
    {
      const float4* qsrc = reinterpret_cast<const float4*>(q + row * kE + h * 16);
      const float4* gsrc = reinterpret_cast<const float4*>(d_out + row * kE + h * 16);
      const float4* osrc = reinterpret_cast<const float4*>(out + row * kE + h * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 tq = valid ? __ldg(qsrc + i) : z, tg = valid ? __ldg(gsrc + i) : z;
        float4 to = valid ? __ldg(osrc + i) : z;
        if (add_res) { to.x -= tq.x; to.y -= tq.y; to.z -= tq.z; to.w -= tq.w; }
        qs[4 * i] = tq.x * kQScale; qs[4 * i + 1] = tq.y * kQScale; qs[4 * i + 2] = tq.z * kQScale; qs[4 * i + 3] = tq.w * kQScale;
        go[4 * i] = tg.x; go[4 * i + 1] = tg.y; go[4 * i + 2] = tg.z; go[4 * i + 3] = tg.w;
        D = fmaf(tg.x, to.x, D); D = fmaf(tg.y, to.y, D); D = fmaf(tg.z, to.z, D); D = fmaf(tg.w, to.w, D);
      }
    }
    const float lse = valid ? __ldg(lse2 + row * kH + h) : 1e30f;
#pragma unroll
    for (int i = 0; i < 16; ++i) gq[i] = 0.f;
    // ---- the 16 rows of this lane's half, column d: dO and q (for the sums over rows) ----
    float cdo[16], cq[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int64_t r = g0 + half * 16 + i;
      const bool ok = r < row_hi;
      cdo[i] = ok ? __ldg(d_out + r * kE + h * 16 + d) : 0.f;
      cq[i] = ok ? __ldg(q + r * kE + h * 16 + d) * 0.25f : 0.f;
    }
    int buf = 0;
#pragma unroll 1
    for (int wi = 0; wi < 4; ++wi) {
      const uint32_t word = mb[lane * kMbStride + wi];
      const int jn = min(32, N - wi * 32);
      // two keys per iteration: two independent dependency chains per warp (the kernel runs two warps per scheduler)
#pragma unroll 1
      for (int jj = 0; jj < jn; jj += 2) {
        const bool bit0 = (word >> jj) & 1u, bit1 = jj + 1 < jn && ((word >> (jj + 1)) & 1u);
        if (!__any_sync(0xffffffffu, bit0 || bit1)) continue;
        const int j0 = wi * 32 + jj, j1 = min(j0 + 1, N - 1);  // j1 clamped: its p / ds are zero when jj + 1 == jn
        const float4* kp0 = reinterpret_cast<const float4*>(Ks + j0 * kE + h * 16);
        const float4* vp0 = reinterpret_cast<const float4*>(Vs + j0 * kE + h * 16);
        const float4* kp1 = reinterpret_cast<const float4*>(Ks + j1 * kE + h * 16);
        const float4* vp1 = reinterpret_cast<const float4*>(Vs + j1 * kE + h * 16);
        float s0 = 0.f, dp0 = 0.f, s1 = 0.f, dp1 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = kp0[i], u = vp0[i], t1 = kp1[i], u1 = vp1[i];
          s0 = fmaf(qs[4 * i], t.x, s0); s0 = fmaf(qs[4 * i + 1], t.y, s0); s0 = fmaf(qs[4 * i + 2], t.z, s0); s0 = fmaf(qs[4 * i + 3], t.w, s0);
          dp0 = fmaf(go[4 * i], u.x, dp0); dp0 = fmaf(go[4 * i + 1], u.y, dp0); dp0 = fmaf(go[4 * i + 2], u.z, dp0); dp0 = fmaf(go[4 * i + 3], u.w, dp0);
          s1 = fmaf(qs[4 * i], t1.x, s1); s1 = fmaf(qs[4 * i + 1], t1.y, s1); s1 = fmaf(qs[4 * i + 2], t1.z, s1); s1 = fmaf(qs[4 * i + 3], t1.w, s1);
          dp1 = fmaf(go[4 * i], u1.x, dp1); dp1 = fmaf(go[4 * i + 1], u1.y, dp1); dp1 = fmaf(go[4 * i + 2], u1.z, dp1); dp1 = fmaf(go[4 * i + 3], u1.w, dp1);
        }
        const float p0 = bit0 ? exp2f(s0 - lse) : 0.f, p1 = bit1 ? exp2f(s1 - lse) : 0.f;
        const float ds0 = p0 * (dp0 - D), ds1 = p1 * (dp1 - D);  // gradients of the natural-log-domain scores q . k / 4
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = kp0[i], t1 = kp1[i];
          gq[4 * i] = fmaf(ds0, t.x, gq[4 * i]); gq[4 * i + 1] = fmaf(ds0, t.y, gq[4 * i + 1]);
          gq[4 * i + 2] = fmaf(ds0, t.z, gq[4 * i + 2]); gq[4 * i + 3] = fmaf(ds0, t.w, gq[4 * i + 3]);
          gq[4 * i] = fmaf(ds1, t1.x, gq[4 * i]); gq[4 * i + 1] = fmaf(ds1, t1.y, gq[4 * i + 1]);
          gq[4 * i + 2] = fmaf(ds1, t1.z, gq[4 * i + 2]); gq[4 * i + 3] = fmaf(ds1, t1.w, gq[4 * i + 3]);
        }
        float* pb = pw + buf * 64;
        float* db = dw + buf * 64;
        pb[lane] = p0; pb[32 + lane] = p1;
        db[lane] = ds0; db[32 + lane] = ds1;
        __syncwarp();
        // sums over the 32 rows: lanes 0-15 finish dV[j][d], lanes 16-31 dK[j][d]
        float av0 = 0.f, ak0 = 0.f, av1 = 0.f, ak1 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = reinterpret_cast<const float4*>(pb + half * 16)[i], u = reinterpret_cast<const float4*>(db + half * 16)[i];
          const float4 t1 = reinterpret_cast<const float4*>(pb + 32 + half * 16)[i], u1 = reinterpret_cast<const float4*>(db + 32 + half * 16)[i];
          av0 = fmaf(t.x, cdo[4 * i], av0); av0 = fmaf(t.y, cdo[4 * i + 1], av0); av0 = fmaf(t.z, cdo[4 * i + 2], av0); av0 = fmaf(t.w, cdo[4 * i + 3], av0);
          ak0 = fmaf(u.x, cq[4 * i], ak0); ak0 = fmaf(u.y, cq[4 * i + 1], ak0); ak0 = fmaf(u.z, cq[4 * i + 2], ak0); ak0 = fmaf(u.w, cq[4 * i + 3], ak0);
          av1 = fmaf(t1.x, cdo[4 * i], av1); av1 = fmaf(t1.y, cdo[4 * i + 1], av1); av1 = fmaf(t1.z, cdo[4 * i + 2], av1); av1 = fmaf(t1.w, cdo[4 * i + 3], av1);
          ak1 = fmaf(u1.x, cq[4 * i], ak1); ak1 = fmaf(u1.y, cq[4 * i + 1], ak1); ak1 = fmaf(u1.z, cq[4 * i + 2], ak1); ak1 = fmaf(u1.w, cq[4 * i + 3], ak1);
        }
        av0 += __shfl_xor_sync(0xffffffffu, av0, 16);
        ak0 += __shfl_xor_sync(0xffffffffu, ak0, 16);
        av1 += __shfl_xor_sync(0xffffffffu, av1, 16);
        ak1 += __shfl_xor_sync(0xffffffffu, ak1, 16);
        float* acc = (half == 0 ? dVs : dKs) + h * 16 + d;
        acc[j0 * kE] += half == 0 ? av0 : ak0;
        if (jj + 1 < jn) acc[j1 * kE] += half == 0 ? av1 : ak1;
        buf ^= 1;
      }
    }
    if (valid) {
      float4* dst = reinterpret_cast<float4*>(dq + row * kE + h * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        dst[i] = add_res ? make_float4(fmaf(gq[4 * i], 0.25f, go[4 * i]), fmaf(gq[4 * i + 1], 0.25f, go[4 * i + 1]),
                                       fmaf(gq[4 * i + 2], 0.25f, go[4 * i + 2]), fmaf(gq[4 * i + 3], 0.25f, go[4 * i + 3]))
                         : make_float4(gq[4 * i] * 0.25f, gq[4 * i + 1] * 0.25f, gq[4 * i + 2] * 0.25f, gq[4 * i + 3] * 0.25f);
    }
  }
  __syncthreads();
  float* gk = dk + b * N * kE;
  float* gv = dv + b * N * kE;
  for (int i = tid; i < N * kE; i += kAtThreads) {
    atomicAdd(gk + i, dKs[i]);
    atomicAdd(gv + i, dVs[i]);
  }
}

static int attn_rows_per_cta(int64_t L, int64_t n_inst, int* ctas_per_inst) {
  // ~1024 rows per CTA (the K / V tile load and the dK / dV flush are amortised), at least ~2 waves of CTAs on the device
  int64_t per = (L + 1023) / 1024;
  const int sms = device_sm_count();
  while (per * n_inst < 2LL * sms && per * 64 < L) ++per;
  int64_t rows = (L + per - 1) / per;
  rows = (rows + kAtGroup - 1) / kAtGroup * kAtGroup;
  *ctas_per_inst = (int)((L + rows - 1) / rows);
  return (int)rows;
}

}  // namespace rrnco

using namespace rrnco;

extern "C" {

int rrnco_train_attention_fwd(int64_t n_inst, int64_t rows_per_inst, int32_t n_nodes, const float* q, const float* k, const float* v,
                              const uint8_t* mask, int32_t add_residual, float* out, float* lse, void* stream) {
  if (n_inst == 0 || rows_per_inst == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(n_inst > 0 && rows_per_inst > 0 && n_nodes > 0 && q && k && v && mask && out && lse);
  RRNCO_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                    reinterpret_cast<uintptr_t>(out)) & 15u) == 0);
  if (n_nodes > 128 || n_inst > 65535) return RRNCO_ERR_UNSUPPORTED;
  const size_t smem = (size_t)2 * n_nodes * kE * sizeof(float) + kAtGroup * kMbStride * sizeof(uint32_t);
  static PerDeviceOnce once;
  if (once.first()) {
    if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * kE * 4 + 1024) != cudaSuccess ||
        cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) {
      once.undo();
      return RRNCO_ERR_CUDA;
    }
  }
  int per = 1;
  const int rows = attn_rows_per_cta(rows_per_inst, n_inst, &per);
  attn_fwd_kernel<<<dim3((unsigned)per, (unsigned)n_inst), kAtThreads, smem, (cudaStream_t)stream>>>(
      rows_per_inst, rows, n_nodes, q, k, v, mask, add_residual, out, lse);
  return rrnco_launch_status();
}

int rrnco_train_attention_bwd(int64_t n_inst, int64_t rows_per_inst, int32_t n_nodes, const float* q, const float* k, const float* v,
                              const uint8_t* mask, const float* out, int32_t add_residual, const float* lse, const float* d_out,
                              float* dq, float* dk, float* dv, void* stream) {
  if (n_inst == 0 || rows_per_inst == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(n_inst > 0 && rows_per_inst > 0 && n_nodes > 0 && q && k && v && mask && out && lse && d_out && dq && dk && dv);
  RRNCO_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                    reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(dq)) & 15u) == 0);
  const size_t smem = (size_t)4 * n_nodes * kE * sizeof(float) + 2 * 8 * 2 * 64 * sizeof(float) + kAtGroup * kMbStride * sizeof(uint32_t);
  if (smem > 227 * 1024 || n_inst > 65535) return RRNCO_ERR_UNSUPPORTED;  // n_nodes <= 108
  static PerDeviceOnce once;
  if (once.first()) {
    if (cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      once.undo();
      return RRNCO_ERR_CUDA;
    }
  }
  if (cudaMemsetAsync(dk, 0, (size_t)n_inst * n_nodes * kE * sizeof(float), (cudaStream_t)stream) != cudaSuccess ||
      cudaMemsetAsync(dv, 0, (size_t)n_inst * n_nodes * kE * sizeof(float), (cudaStream_t)stream) != cudaSuccess)
    return RRNCO_ERR_CUDA;
  int per = 1;
  const int rows = attn_rows_per_cta(rows_per_inst, n_inst, &per);
  attn_bwd_kernel<<<dim3((unsigned)per, (unsigned)n_inst), kAtThreads, smem, (cudaStream_t)stream>>>(
      rows_per_inst, rows, n_nodes, q, k, v, mask, out, add_residual, lse, d_out, dq, dk, dv);
  return rrnco_launch_status();
}

}  // extern "C"
