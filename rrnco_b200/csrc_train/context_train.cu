// Training hand-off, context query of the pointer (rrnco/models/env_embeddings/context.py:18-70 + decoder.py:151-170) for the
// rows of the batched replay:  q = project_context([emb(node_a), emb(node_b)?, state scalars]).
// The projection is linear, so the node part is a row of a per-instance table  P_t = row_emb W_t^T  ([n_inst, N, 128], a tiny
// GEMM that stays in torch with its autograd) and the query of a row is a gather plus a rank-k update:
//     q[row] = P_a[b, ia[row]] (+ P_b[b, ib[row]]) + sum_s state[row, s] Ws[s]
// instead of a [rows, 129..256] concatenation and a [rows x 129] x [129 x 128] fp32 GEMM; the backward pass scatters dq into the
// tables with vector atomics (the tables are L2-resident) and reduces dWs per CTA.
#include "../csrc/common.cuh"
#include "../../include/rrnco_b200_train.h"

namespace rrnco {

constexpr int kCtxMaxState = 8;

__global__ void __launch_bounds__(256) context_query_fwd_kernel(int64_t rows, int64_t L, int N, const float* __restrict__ pa,
                                                               const int64_t* __restrict__ ia, const float* __restrict__ pb,
                                                               const int64_t* __restrict__ ib, const float* __restrict__ state,
                                                               int n_state, const float* __restrict__ ws, float* __restrict__ q) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  float4 w[kCtxMaxState];
#pragma unroll
  for (int s = 0; s < kCtxMaxState; ++s) w[s] = s < n_state ? __ldg(reinterpret_cast<const float4*>(ws + s * kE) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const int64_t b = row / L;
    float4 v = __ldg(reinterpret_cast<const float4*>(pa + (b * N + __ldg(ia + row)) * kE) + lane);
    if (pb != nullptr) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(pb + (b * N + __ldg(ib + row)) * kE) + lane);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
#pragma unroll
    for (int s = 0; s < kCtxMaxState; ++s)
      if (s < n_state) {
        const float t = __ldg(state + row * n_state + s);
        v.x = fmaf(t, w[s].x, v.x); v.y = fmaf(t, w[s].y, v.y); v.z = fmaf(t, w[s].z, v.z); v.w = fmaf(t, w[s].w, v.w);
      }
    reinterpret_cast<float4*>(q + row * kE)[lane] = v;
  }
}

__global__ void __launch_bounds__(256) context_query_bwd_kernel(int64_t rows, int64_t L, int N, const float* __restrict__ dq,
                                                               const int64_t* __restrict__ ia, const int64_t* __restrict__ ib,
                                                               const float* __restrict__ state, int n_state, float* __restrict__ dpa,
                                                               float* __restrict__ dpb, float* __restrict__ dws) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * 8 + warp, nwarps = (int64_t)gridDim.x * 8;
  float4 acc[kCtxMaxState];
#pragma unroll
  for (int s = 0; s < kCtxMaxState; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const int64_t b = row / L;
    const float4 g = __ldg(reinterpret_cast<const float4*>(dq + row * kE) + lane);
    atomicAdd(reinterpret_cast<float4*>(dpa + (b * N + __ldg(ia + row)) * kE) + lane, g);
    if (dpb != nullptr) atomicAdd(reinterpret_cast<float4*>(dpb + (b * N + __ldg(ib + row)) * kE) + lane, g);
#pragma unroll
    for (int s = 0; s < kCtxMaxState; ++s)
      if (s < n_state) {
        const float t = __ldg(state + row * n_state + s);
        acc[s].x = fmaf(t, g.x, acc[s].x); acc[s].y = fmaf(t, g.y, acc[s].y); acc[s].z = fmaf(t, g.z, acc[s].z); acc[s].w = fmaf(t, g.w, acc[s].w);
      }
  }
#pragma unroll
  for (int s = 0; s < kCtxMaxState; ++s) {  // sum over the 8 warps, then one atomic per column
    if (s >= n_state) break;
    red[warp][lane] = acc[s];
    __syncthreads();
    if (warp == 0) {
      float4 t = red[0][lane];
#pragma unroll
      for (int i = 1; i < 8; ++i) {
        const float4 u = red[i][lane];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      atomicAdd(reinterpret_cast<float4*>(dws + s * kE) + lane, t);
    }
    __syncthreads();
  }
}

}  // namespace rrnco

using namespace rrnco;

extern "C" {

int rrnco_train_context_query_fwd(int64_t rows, int64_t rows_per_inst, int32_t n_nodes, const float* table_a, const int64_t* index_a,
                                  const float* table_b, const int64_t* index_b, const float* state, int32_t n_state,
                                  const float* state_w, float* q, void* stream) {
  if (rows == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(rows > 0 && rows_per_inst > 0 && n_nodes > 0 && table_a && index_a && q && n_state >= 0 && n_state <= kCtxMaxState);
  RRNCO_CHECK_ARG((table_b == nullptr) == (index_b == nullptr) && (n_state == 0 || (state && state_w)));
  RRNCO_CHECK_ARG(((reinterpret_cast<uintptr_t>(table_a) | reinterpret_cast<uintptr_t>(table_b) | reinterpret_cast<uintptr_t>(state_w) |
                    reinterpret_cast<uintptr_t>(q)) & 15u) == 0);
  int64_t grid = (rows + 7) / 8;
  const int64_t cap = 16LL * device_sm_count();
  if (grid > cap) grid = cap;
  context_query_fwd_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(rows, rows_per_inst, n_nodes, table_a, index_a, table_b,
                                                                          index_b, state, n_state, state_w, q);
  return rrnco_launch_status();
}

int rrnco_train_context_query_bwd(int64_t rows, int64_t rows_per_inst, int32_t n_nodes, const float* dq, const int64_t* index_a,
                                  const int64_t* index_b, const float* state, int32_t n_state, float* d_table_a, float* d_table_b,
                                  float* d_state_w, void* stream) {
  if (rows == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(rows > 0 && rows_per_inst > 0 && n_nodes > 0 && dq && index_a && d_table_a && n_state >= 0 && n_state <= kCtxMaxState);
  RRNCO_CHECK_ARG((d_table_b == nullptr) == (index_b == nullptr) && (n_state == 0 || (state && d_state_w)));
  RRNCO_CHECK_ARG(((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(d_table_a) | reinterpret_cast<uintptr_t>(d_table_b) |
                    reinterpret_cast<uintptr_t>(d_state_w)) & 15u) == 0);
  int64_t grid = (rows + 7) / 8;
  const int64_t cap = 8LL * device_sm_count();
  if (grid > cap) grid = cap;
  context_query_bwd_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(rows, rows_per_inst, n_nodes, dq, index_a, index_b, state,
                                                                          n_state, d_table_a, d_table_b, d_state_w);
  return rrnco_launch_status();
}

}  // extern "C"
