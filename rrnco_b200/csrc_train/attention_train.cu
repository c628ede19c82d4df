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
#include "../csrc/ffn_pack.cuh"   // ffma2 / fadd2: packed fp32 pairs (two IEEE FMAs per issue slot, same bits as the scalar forms)
#include "../../include/rrnco_b200_train.h"

namespace rrnco {

constexpr int kAtThreads = 256;
constexpr int kAtGroup = 32;      // rows per group (lane = row)
constexpr int kMbStride = 5;      // mask words per row in shared memory (4 + 1 pad)
constexpr float kQScale = 0.25f * 1.4426950408889634f;  // 1 / sqrt(16) and log2(e): scores in the log2 domain

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

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

// Forward: two rows per lane (rows g0 + lane and g0 + 32 + lane of a 64-row group), so that one broadcast K / V read serves two
// rows: the kernel is bound by shared memory -> register bandwidth (32 crossbar cycles per K + V row of a head, whatever the
// number of lanes that read it), not by the FMAs.
__global__ void __launch_bounds__(kAtThreads, 2) attn_fwd_kernel(int64_t L, int rows_per_cta, int N, const float* __restrict__ q,
                                                                const float* __restrict__ k, const float* __restrict__ v,
                                                                const uint8_t* __restrict__ mask, int add_res,
                                                                float* __restrict__ out, float* __restrict__ lse2) {
  extern __shared__ __align__(16) float at_smem[];
  float* Ks = at_smem;
  float* Vs = Ks + N * kE;
  uint32_t* mb = reinterpret_cast<uint32_t*>(Vs + N * kE);   // [64 rows][kMbStride]
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
  for (int64_t g0 = row_lo; g0 < row_hi; g0 += 2 * kAtGroup) {
    __syncthreads();
    stage_mask_bits(mb, mask, g0, row_hi, N, tid);
    stage_mask_bits(mb + kAtGroup * kMbStride, mask, g0 + kAtGroup, row_hi, N, tid);
    __syncthreads();
    const int64_t rowa = g0 + lane, rowb = g0 + kAtGroup + lane;
    const bool va = rowa < row_hi, vb = rowb < row_hi;
    if (rowa + 2 * kAtGroup < row_hi) {  // the next group's rows stream from HBM: start them now
      prefetch_l2(q + (rowa + 2 * kAtGroup) * kE + h * 16);
      prefetch_l2(q + (rowb + 2 * kAtGroup) * kE + h * 16);
      if (tid * 32 < 2 * kAtGroup * N) prefetch_l2(mask + (g0 + 2 * kAtGroup) * N + tid * 32);
    }
    float qa[16], qb[16], oa[16], ob[16];
    {
      const float4* sa = reinterpret_cast<const float4*>(q + rowa * kE + h * 16);
      const float4* sb = reinterpret_cast<const float4*>(q + rowb * kE + h * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 ta = va ? __ldg(sa + i) : z, tb = vb ? __ldg(sb + i) : z;
        qa[4 * i] = ta.x * kQScale; qa[4 * i + 1] = ta.y * kQScale; qa[4 * i + 2] = ta.z * kQScale; qa[4 * i + 3] = ta.w * kQScale;
        qb[4 * i] = tb.x * kQScale; qb[4 * i + 1] = tb.y * kQScale; qb[4 * i + 2] = tb.z * kQScale; qb[4 * i + 3] = tb.w * kQScale;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) oa[i] = ob[i] = 0.f;
    float ma = -1e30f, la = 0.f, mbx = -1e30f, lb = 0.f;
#pragma unroll 1
    for (int wi = 0; wi < 4; ++wi) {
      const uint32_t worda = mb[lane * kMbStride + wi], wordb = mb[(kAtGroup + lane) * kMbStride + wi];
      const int jn = min(32, N - wi * 32);
#pragma unroll 1
      for (int jj = 0; jj < jn; ++jj) {
        const bool bita = (worda >> jj) & 1u, bitb = (wordb >> jj) & 1u;
        if (!__any_sync(0xffffffffu, bita || bitb)) continue;
        const int j = wi * 32 + jj;
        const float4* kp = reinterpret_cast<const float4*>(Ks + j * kE + h * 16);
        // packed fp32 pairs (FFMA2), two independent chains per dot product
        float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = kp[i];
          a0 = ffma2(make_float2(qa[4 * i], qa[4 * i + 1]), make_float2(t.x, t.y), a0);
          a1 = ffma2(make_float2(qa[4 * i + 2], qa[4 * i + 3]), make_float2(t.z, t.w), a1);
          b0 = ffma2(make_float2(qb[4 * i], qb[4 * i + 1]), make_float2(t.x, t.y), b0);
          b1 = ffma2(make_float2(qb[4 * i + 2], qb[4 * i + 3]), make_float2(t.z, t.w), b1);
        }
        const float sa = (a0.x + a0.y) + (a1.x + a1.y), sb = (b0.x + b0.y) + (b1.x + b1.y);
        if (bita && sa > ma + 8.f) {  // lazy rescale: rare after the first feasible key
          const float corr = exp2f(ma - sa);
          la *= corr;
#pragma unroll
          for (int i = 0; i < 16; ++i) oa[i] *= corr;
          ma = sa;
        }
        if (bitb && sb > mbx + 8.f) {
          const float corr = exp2f(mbx - sb);
          lb *= corr;
#pragma unroll
          for (int i = 0; i < 16; ++i) ob[i] *= corr;
          mbx = sb;
        }
        const float wa = bita ? exp2f(sa - ma) : 0.f, wb = bitb ? exp2f(sb - mbx) : 0.f;
        la += wa;
        lb += wb;
        const float4* vp = reinterpret_cast<const float4*>(Vs + j * kE + h * 16);
        const float2 wa2 = make_float2(wa, wa), wb2 = make_float2(wb, wb);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = vp[i];
          const float2 x0 = ffma2(wa2, make_float2(t.x, t.y), make_float2(oa[4 * i], oa[4 * i + 1]));
          const float2 x1 = ffma2(wa2, make_float2(t.z, t.w), make_float2(oa[4 * i + 2], oa[4 * i + 3]));
          const float2 y0 = ffma2(wb2, make_float2(t.x, t.y), make_float2(ob[4 * i], ob[4 * i + 1]));
          const float2 y1 = ffma2(wb2, make_float2(t.z, t.w), make_float2(ob[4 * i + 2], ob[4 * i + 3]));
          oa[4 * i] = x0.x; oa[4 * i + 1] = x0.y; oa[4 * i + 2] = x1.x; oa[4 * i + 3] = x1.y;
          ob[4 * i] = y0.x; ob[4 * i + 1] = y0.y; ob[4 * i + 2] = y1.x; ob[4 * i + 3] = y1.y;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int64_t row = r == 0 ? rowa : rowb;
      if (row < row_hi) {
        const float l = r == 0 ? la : lb, m = r == 0 ? ma : mbx;
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        float4* dst = reinterpret_cast<float4*>(out + row * kE + h * 16);
        const float4* src = reinterpret_cast<const float4*>(q + row * kE + h * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float* o = r == 0 ? oa : ob;
          float4 t = make_float4(o[4 * i] * inv, o[4 * i + 1] * inv, o[4 * i + 2] * inv, o[4 * i + 3] * inv);
          if (add_res) {
            const float4 u = __ldg(src + i);
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
          }
          dst[i] = t;
        }
        lse2[row * kH + h] = l > 0.f ? m + log2f(l) : 1e30f;
      }
    }
  }
}

// Backward.  Shared memory -> register bandwidth (128 B per clock per SM, broadcast or not) is what bounds these kernels: a
// K / V row costs 32 crossbar cycles per (32 rows, key, head).  The sums over rows (dV = P^T dO, dK = dS^T q) therefore do NOT
// go through per-lane broadcast reads of p / ds (another 32 cycles): they are 16 x 8 x 32 products on the legacy tensor path
// (mma.sync m16n8k8, 3xTF32 = fp32-faithful): A = the head's dO^T / q^T columns of the 32 rows (raw fp32 in registers, split on
// the fly), B = the p / ds values of 8 keys as each lane wrote them ([key][row], conflict-free fragment loads: 2 cycles per key).
constexpr int kPsStride = 36;   // floats per key row of the p / ds hand-over tiles: (36 g + t) % 32 distinct over a fragment load
constexpr int kAccStride = 132; // floats per key row of the dK / dV tiles: (2 t * 132 + g) % 32 distinct over a fragment update
__global__ void __launch_bounds__(kAtThreads, 1) attn_bwd_kernel(int64_t L, int rows_per_cta, int N, const float* __restrict__ q,
                                                                const float* __restrict__ k, const float* __restrict__ v,
                                                                const uint8_t* __restrict__ mask, const float* __restrict__ out,
                                                                int add_res, const float* __restrict__ lse2,
                                                                const float* __restrict__ d_out, float* __restrict__ dq,
                                                                float* __restrict__ dk, float* __restrict__ dv) {
  extern __shared__ __align__(16) float at_smem[];
  float* Ks = at_smem;
  float* Vs = Ks + N * kE;
  float* dKs = Vs + N * kE;                 // [N][kAccStride]
  float* dVs = dKs + N * kAccStride;
  float* ps = dVs + N * kAccStride;         // [8 warps][8 keys][kPsStride]
  float* dss = ps + 8 * 8 * kPsStride;
  uint32_t* mb = reinterpret_cast<uint32_t*>(dss + 8 * 8 * kPsStride);
  const int tid = threadIdx.x, h = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t b = blockIdx.y;
  {
    const float4* k4 = reinterpret_cast<const float4*>(k + b * N * kE);
    const float4* v4 = reinterpret_cast<const float4*>(v + b * N * kE);
    for (int i = tid; i < N * (kE / 4); i += kAtThreads) {
      reinterpret_cast<float4*>(Ks)[i] = __ldg(k4 + i);
      reinterpret_cast<float4*>(Vs)[i] = __ldg(v4 + i);
    }
    for (int i = tid; i < 2 * N * kAccStride; i += kAtThreads) dKs[i] = 0.f;   // dKs and dVs are adjacent
  }
  float* pw = ps + h * 8 * kPsStride;
  float* dw = dss + h * 8 * kPsStride;
  const int64_t row_lo = b * L + (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row_hi = min(row_lo + rows_per_cta, (b + 1) * L);
  for (int64_t g0 = row_lo; g0 < row_hi; g0 += kAtGroup) {
    __syncthreads();
    stage_mask_bits(mb, mask, g0, row_hi, N, tid);
    __syncthreads();
    const int64_t row = g0 + lane;
    const bool valid = row < row_hi;
    if (row + kAtGroup < row_hi) {  // the next group's rows stream from HBM: start them now (one CTA per SM leaves nothing else to hide the latency)
      prefetch_l2(q + (row + kAtGroup) * kE + h * 16);
      prefetch_l2(d_out + (row + kAtGroup) * kE + h * 16);
      prefetch_l2(out + (row + kAtGroup) * kE + h * 16);
      if (tid * 16 < kAtGroup * N) prefetch_l2(mask + (g0 + kAtGroup) * N + tid * 16);
    }
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
    // ---- A fragments of the row sums: dO^T and (q / 4)^T of this head, [16 columns x 32 rows], raw fp32 ----
    float ado[4][4], aq[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t r = g0 + ks * 8 + t + (i >> 1) * 4;       // a0: (g, t)  a1: (g + 8, t)  a2: (g, t + 4)  a3: (g + 8, t + 4)
        const int col = h * 16 + g + (i & 1) * 8;
        const bool ok = r < row_hi;
        ado[ks][i] = ok ? __ldg(d_out + r * kE + col) : 0.f;
        aq[ks][i] = ok ? __ldg(q + r * kE + col) * 0.25f : 0.f;
      }
    const uint32_t w0 = mb[lane * kMbStride], w1 = mb[lane * kMbStride + 1], w2 = mb[lane * kMbStride + 2], w3 = mb[lane * kMbStride + 3];
#pragma unroll 1
    for (int j0 = 0; j0 < N; j0 += 8) {
      const uint32_t word = (j0 >> 5) == 0 ? w0 : (j0 >> 5) == 1 ? w1 : (j0 >> 5) == 2 ? w2 : w3;
      const uint32_t bits8 = (word >> (j0 & 31)) & 0xffu;   // keys beyond N have no bits
      if (!__any_sync(0xffffffffu, bits8 != 0u)) continue;
      // ---- phase 1: p and ds of 8 keys for this lane's row; dq in registers ----
#pragma unroll 2
      for (int kk = 0; kk < 8; ++kk) {
        const int j = min(j0 + kk, N - 1);
        const bool bit = (bits8 >> kk) & 1u;
        const float4* kp = reinterpret_cast<const float4*>(Ks + j * kE + h * 16);
        const float4* vp = reinterpret_cast<const float4*>(Vs + j * kE + h * 16);
        float kr[16];
        float sp[4], dpp[4];   // four independent partial sums each: the kernel runs two warps per scheduler, a 16-long FMA chain stalls it
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = kp[i], u = vp[i];
          kr[4 * i] = a.x; kr[4 * i + 1] = a.y; kr[4 * i + 2] = a.z; kr[4 * i + 3] = a.w;
          sp[i] = fmaf(qs[4 * i + 3], a.w, fmaf(qs[4 * i + 2], a.z, fmaf(qs[4 * i + 1], a.y, qs[4 * i] * a.x)));
          dpp[i] = fmaf(go[4 * i + 3], u.w, fmaf(go[4 * i + 2], u.z, fmaf(go[4 * i + 1], u.y, go[4 * i] * u.x)));
        }
        const float s = (sp[0] + sp[1]) + (sp[2] + sp[3]), dp = (dpp[0] + dpp[1]) + (dpp[2] + dpp[3]);
        const float p = bit ? exp2f(s - lse) : 0.f;
        const float ds = p * (dp - D);  // gradient of the natural-log-domain score q . k / 4
#pragma unroll
        for (int i = 0; i < 16; ++i) gq[i] = fmaf(ds, kr[i], gq[i]);   // (packed FFMA2 pairs were measured here: 4 % slower)
        pw[kk * kPsStride + lane] = p;
        dw[kk * kPsStride + lane] = ds;
      }
      __syncwarp();
      // ---- phase 2: C[column d][key] += A[d][row] B[row][key] over the 32 rows, for dV (A = dO^T, B = p) and dK (A = q^T / 4, B = ds) ----
      float cvp[4][4], ckp[4][4];   // one accumulator per K step: four independent MMA chains of three instead of one of twelve
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int i = 0; i < 4; ++i) cvp[ks][i] = ckp[ks][i] = 0.f;
        uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(ado[ks][i], ah[i], al[i]);
        split_tf32(pw[g * kPsStride + ks * 8 + t], bh[0], bl[0]);
        split_tf32(pw[g * kPsStride + ks * 8 + t + 4], bh[1], bl[1]);
        mma_x<3>(cvp[ks], ah, al, bh, bl);
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(aq[ks][i], ah[i], al[i]);
        split_tf32(dw[g * kPsStride + ks * 8 + t], bh[0], bl[0]);
        split_tf32(dw[g * kPsStride + ks * 8 + t + 4], bh[1], bl[1]);
        mma_x<3>(ckp[ks], ah, al, bh, bl);
      }
      float cv[4], ck[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cv[i] = (cvp[0][i] + cvp[1][i]) + (cvp[2][i] + cvp[3][i]);
        ck[i] = (ckp[0][i] + ckp[1][i]) + (ckp[2][i] + ckp[3][i]);
      }
      // c0: (d = g, key 2t)  c1: (g, 2t + 1)  c2: (g + 8, 2t)  c3: (g + 8, 2t + 1); this warp owns the head's 16 columns
      {
        const int ja = j0 + 2 * t, col = h * 16 + g;
        if (ja < N) {
          dVs[ja * kAccStride + col] += cv[0]; dVs[ja * kAccStride + col + 8] += cv[2];
          dKs[ja * kAccStride + col] += ck[0]; dKs[ja * kAccStride + col + 8] += ck[2];
        }
        if (ja + 1 < N) {
          dVs[(ja + 1) * kAccStride + col] += cv[1]; dVs[(ja + 1) * kAccStride + col + 8] += cv[3];
          dKs[(ja + 1) * kAccStride + col] += ck[1]; dKs[(ja + 1) * kAccStride + col + 8] += ck[3];
        }
      }
      __syncwarp();   // the next block of keys overwrites the hand-over tiles
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
    const int j = i >> 7, c = i & 127;
    atomicAdd(gk + i, dKs[j * kAccStride + c]);
    atomicAdd(gv + i, dVs[j * kAccStride + c]);
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
  const size_t smem = (size_t)2 * n_nodes * kE * sizeof(float) + 2 * kAtGroup * kMbStride * sizeof(uint32_t);
  static PerDeviceOnce once;
  if (once.first()) {
    if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * kE * 4 + 2048) != cudaSuccess ||
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
  const size_t smem = ((size_t)2 * n_nodes * kE + (size_t)2 * n_nodes * kAccStride + 2 * 8 * 8 * kPsStride) * sizeof(float) +
                      kAtGroup * kMbStride * sizeof(uint32_t);
  if (smem > 227 * 1024 || n_inst > 65535) return RRNCO_ERR_UNSUPPORTED;  // n_nodes <= 102
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
