// RRNetDecoder._precompute_cache (rrnco/models/decoder.py:214-232) + the node half of the context
// projection (rrnco/models/env_embeddings/context.py:27-31), once per batch:
//   K | V | Lk = col_emb . W_node^T (three 128-wide column blocks),  P = row_emb . W_ctx[:, :E]^T
//   (ATSP: P1 = row_emb . W_ctx[:, :E]^T for the first node, P2 = row_emb . W_ctx[:, E:2E]^T for the current).
// Each CTA computes a 128 x 128 output block with K = 128 entirely from shared memory, 3xTF32 mma.
#include "common.cuh"

namespace rrnco {

constexpr int kGLd = 132;

struct GemmJob {
  const float* A;   // [M, 128] activations
  const float* W;   // weight rows (output dims), row stride ldw, first input column at off
  int ldw, off;
  float* C;         // [M, 128]
};
struct GemmJobs {
  GemmJob job[5];
  int64_t M;
};

__global__ void __launch_bounds__(256, 1) gemm128_kernel(const GemmJobs jobs) {
  extern __shared__ __align__(16) float gs[];
  float* sA = gs;
  float* sW = gs + 128 * kGLd;
  const GemmJob jb = jobs.job[blockIdx.y];
  const int64_t m0 = (int64_t)blockIdx.x * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int idx = tid; idx < 128 * 32; idx += 256) {
    const int row = idx >> 5, c4 = idx & 31;
    const bool ok = m0 + row < jobs.M;
    cp_async16_zfill(sA + row * kGLd + c4 * 4, jb.A + (ok ? (m0 + row) : 0) * kE + c4 * 4, ok);
  }
  cp_async_commit();
  for (int idx = tid; idx < 128 * 128; idx += 256) {
    const int row = idx >> 7, k = idx & 127;
    sW[row * kGLd + k] = __ldg(jb.W + (size_t)row * jb.ldw + jb.off + k);
  }
  cp_async_wait<0>();
  __syncthreads();
  const int wm = warp >> 1, wn = warp & 1;
  float acc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll 2
  for (int ks = 0; ks < 16; ++ks) {
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const float* ap = sA + (wm * 32 + mt * 16 + g) * kGLd + ks * 8 + t;
      split_tf32(ap[0], ah[mt][0], al[mt][0]);
      split_tf32(ap[8 * kGLd], ah[mt][1], al[mt][1]);
      split_tf32(ap[4], ah[mt][2], al[mt][2]);
      split_tf32(ap[8 * kGLd + 4], ah[mt][3], al[mt][3]);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float* bp = sW + (wn * 64 + nt * 8 + g) * kGLd + ks * 8 + t;
      uint32_t bh[2], bl[2];
      split_tf32(bp[0], bh[0], bl[0]);
      split_tf32(bp[4], bh[1], bl[1]);
      mma_x<3>(acc[0][nt], ah[0], al[0], bh, bl);
      mma_x<3>(acc[1][nt], ah[1], al[1], bh, bl);
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = wn * 64 + nt * 8 + 2 * t;
      const int64_t row = m0 + wm * 32 + mt * 16 + g;
      if (row < jobs.M) *reinterpret_cast<float2*>(jb.C + row * kE + col) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
      if (row + 8 < jobs.M)
        *reinterpret_cast<float2*>(jb.C + (row + 8) * kE + col) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
    }
}

}  // namespace rrnco

using namespace rrnco;

extern "C" int rrnco_precompute_cache(int32_t env, int64_t n_inst, int32_t n_nodes, const float* row_emb,
                                      const float* col_emb, const float* w_node, const float* w_ctx, int32_t ctx_in,
                                      float* glimpse_key, float* glimpse_val, float* logit_key, float* ctx_node_proj,
                                      float* ctx_node_proj2, void* stream) {
  RRNCO_CHECK_ARG(n_inst > 0 && n_nodes > 0 && row_emb && col_emb && w_node && w_ctx && glimpse_key && glimpse_val &&
                  logit_key && ctx_node_proj);
  RRNCO_CHECK_ARG(env >= 0 && env <= 2 && ctx_in >= kE);
  RRNCO_CHECK_ARG(env != RRNCO_ENV_ATSP || (ctx_node_proj2 && ctx_in == 2 * kE));
  RRNCO_CHECK_ARG(((uintptr_t)row_emb & 15) == 0 && ((uintptr_t)col_emb & 15) == 0);
  GemmJobs jobs{};
  jobs.M = n_inst * n_nodes;
  jobs.job[0] = {col_emb, w_node, kE, 0, glimpse_key};
  jobs.job[1] = {col_emb, w_node + (size_t)kE * kE, kE, 0, glimpse_val};
  jobs.job[2] = {col_emb, w_node + (size_t)2 * kE * kE, kE, 0, logit_key};
  jobs.job[3] = {row_emb, w_ctx, ctx_in, 0, ctx_node_proj};
  int n_jobs = 4;
  if (env == RRNCO_ENV_ATSP) jobs.job[n_jobs++] = {row_emb, w_ctx, ctx_in, kE, ctx_node_proj2};
  const size_t smem = 2 * 128 * kGLd * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) {
    if (cudaFuncSetAttribute(gemm128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      once.undo();
      return RRNCO_ERR_CUDA;
    }
  }
  dim3 grid((unsigned)((jobs.M + 127) / 128), n_jobs);
  gemm128_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(jobs);
  return rrnco_launch_status();
}
