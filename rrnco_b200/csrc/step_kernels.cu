// Per-step decoder kernels for ANY number of nodes (the key-streaming variant used when N exceeds the
// 128-key tile of the fused kernel, e.g. BASELINE config "ATSP n=1000 generalisation rollout"):
//   rrnco_decoder_logits_large  = RRNetDecoder.forward            rrnco/models/decoder.py:151-206
//       attention_stream_kernel : context query + 8-head masked attention, online softmax over streamed keys
//       pointer FFN             : the tcgen05 kernel of ffn_tc_kernel.cu (rows = rollouts)
//       logits_stream_kernel    : g'.Lk^T / sqrt(E), scale-adaptive bias, log(exp(.) + 1e-6)
//   rrnco_select_action         = DecodingStrategy.step           rrnco/models/decoding.py:219-298,311-361
// One warp per rollout; keys / values / logit keys are streamed from L2 with coalesced 128-bit loads (a key row is
// 512 contiguous bytes = one float4 per lane); infeasible keys are skipped.  Instance data is never replicated:
// rollout r = s * n_inst + b reads the cache / matrices of instance b.
#include "common.cuh"
#include "ffn_pack.cuh"

namespace rrnco {

constexpr int kStepWarps = 4;

__device__ __forceinline__ float sfexp(float x) { return __expf(x); }

struct StepArgs {
  int env, N, n_state, use_placeholder;
  int64_t R, n_inst, data_rows;
  const float *K, *V, *Lk, *P1, *P2;           // [n_inst, N, 128]
  const float *wstate, *placeholder;            // [n_state, 128], [128]
  const float *dist, *dur;                      // [data_rows, N, N]
  float alpha, beta;
  const int64_t *cur, *first;                   // [R]
  const uint8_t* mask;                          // [R, N] bool
  const float* state;                           // [R, n_state]
  float* G;                                     // [R, 128] glimpse (attention out) / g' (FFN out)
  float* logits;                                // [R, N]
  uint32_t* status;
};

__global__ void __launch_bounds__(kStepWarps * 32) attention_stream_kernel(const StepArgs a) {
  const int64_t r = (int64_t)blockIdx.x * kStepWarps + (threadIdx.x >> 5);
  if (r >= a.R) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = r % a.n_inst;
  const int N = a.N;
  const float4* Kb = reinterpret_cast<const float4*>(a.K + b * (int64_t)N * kE);
  const float4* Vb = reinterpret_cast<const float4*>(a.V + b * (int64_t)N * kE);
  const int cur = (int)a.cur[r];
  // query: node half of the context projection (precomputed) + state scalars . state weights
  float4 q;
  if (a.env == RRNCO_ENV_ATSP) {
    if (a.use_placeholder) {
      q = __ldg(reinterpret_cast<const float4*>(a.placeholder) + lane);
    } else {
      const float4 f4 = __ldg(reinterpret_cast<const float4*>(a.P1 + (b * N + (int)a.first[r]) * (int64_t)kE) + lane);
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(a.P2 + (b * N + cur) * (int64_t)kE) + lane);
      q = make_float4(f4.x + c4.x, f4.y + c4.y, f4.z + c4.z, f4.w + c4.w);
    }
  } else {
    q = __ldg(reinterpret_cast<const float4*>(a.P1 + (b * N + cur) * (int64_t)kE) + lane);
    for (int k = 0; k < a.n_state; ++k) {
      const float st = a.state[r * a.n_state + k];
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.wstate + k * kE) + lane);
      q.x = fmaf(st, w4.x, q.x); q.y = fmaf(st, w4.y, q.y); q.z = fmaf(st, w4.z, q.z); q.w = fmaf(st, w4.w, q.w);
    }
  }
  // lane l owns dims 4l..4l+3 of head l >> 2; online softmax state is replicated over the 4 lanes of a head
  float m = -INFINITY, l = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const uint8_t* mrow = a.mask + r * (int64_t)N;
  for (int n0 = 0; n0 < N; n0 += 32) {
    const int nn = n0 + lane;
    uint32_t feas = __ballot_sync(0xffffffffu, nn < N && mrow[nn] != 0);
    while (feas) {
      const int j = __ffs(feas) - 1;
      feas &= feas - 1;
      const int n = n0 + j;
      const float4 k4 = __ldg(Kb + (size_t)n * 32 + lane);
      float s = q.x * k4.x + q.y * k4.y + q.z * k4.z + q.w * k4.w;
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s *= 0.25f;  // 1 / sqrt(head dim 16)
      const float mn = fmaxf(m, s);
      const float corr = sfexp(m - mn), p = sfexp(s - mn);
      const float4 v4 = __ldg(Vb + (size_t)n * 32 + lane);
      l = l * corr + p;
      acc.x = acc.x * corr + p * v4.x; acc.y = acc.y * corr + p * v4.y;
      acc.z = acc.z * corr + p * v4.z; acc.w = acc.w * corr + p * v4.w;
      m = mn;
    }
  }
  const float inv = l > 0.f ? 1.0f / l : 0.f;
  if (l <= 0.f && lane == 0) atomicOr(a.status, RRNCO_DEV_NO_FEASIBLE);
  float4 g4 = make_float4(acc.x * inv + q.x, acc.y * inv + q.y, acc.z * inv + q.z, acc.w * inv + q.w);
  reinterpret_cast<float4*>(a.G + r * (int64_t)kE)[lane] = g4;  // glimpse = heads + q (decoder.py:292-293)
}

__global__ void __launch_bounds__(kStepWarps * 32) logits_stream_kernel(const StepArgs a) {
  __shared__ float4 sg[kStepWarps][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kStepWarps + w;
  if (r >= a.R) return;
  const int64_t b = r % a.n_inst, drow = b % a.data_rows;
  const int N = a.N;
  sg[w][lane] = reinterpret_cast<const float4*>(a.G + r * (int64_t)kE)[lane];
  __syncwarp();
  const int cur = (int)a.cur[r];
  const float* Lkb = a.Lk + b * (int64_t)N * kE;
  const float* Drow = a.dist + (drow * N + cur) * (int64_t)N;
  const float* Urow = a.env == RRNCO_ENV_RCVRPTW ? a.dur + (drow * N + cur) * (int64_t)N : nullptr;
  bool nan_seen = false;
  for (int n0 = 0; n0 < N; n0 += 32) {
    const int n = n0 + lane;  // one key per lane, its 512-byte row streamed as 32 float4
    if (n < N) {
      const float4* krow = reinterpret_cast<const float4*>(Lkb + (size_t)n * kE);
      float acc = 0.f;
#pragma unroll 8
      for (int d4 = 0; d4 < 32; ++d4) {
        const float4 k4 = __ldg(krow + d4);
        const float4 g4 = sg[w][d4];
        acc = fmaf(g4.x, k4.x, acc); acc = fmaf(g4.y, k4.y, acc); acc = fmaf(g4.z, k4.z, acc); acc = fmaf(g4.w, k4.w, acc);
      }
      float lg = acc * 0.08838834764831845f;  // 1 / sqrt(128)
      nan_seen |= lg != lg;
      float bias = __fmul_rn(a.alpha, Drow[n]);
      if (Urow) bias = __fadd_rn(bias, __fmul_rn(a.beta, Urow[n]));
      lg = __logf(__fadd_rn(sfexp(__fsub_rn(lg, bias)), 1e-6f));  // decoder.py:198
      a.logits[r * (int64_t)N + n] = lg;
    }
  }
  if (__any_sync(0xffffffffu, nan_seen) && lane == 0) atomicOr(a.status, RRNCO_DEV_NAN_LOGITS);
}

// DecodingStrategy.step: process_logits (tanh clip, mask, temperature, log-softmax) + greedy / Gumbel-max / forced
__global__ void __launch_bounds__(kStepWarps * 32) select_action_kernel(int64_t R, int N, const float* __restrict__ logits,
                                                                       const uint8_t* __restrict__ mask, int mode,
                                                                       float tanh_clip, float temperature, uint64_t seed,
                                                                       int step, const int64_t* __restrict__ forced,
                                                                       int64_t* __restrict__ action_out,
                                                                       float* __restrict__ logp_out, uint32_t* status) {
  const int64_t r = (int64_t)blockIdx.x * kStepWarps + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const float* lrow = logits + r * (int64_t)N;
  const uint8_t* mrow = mask + r * (int64_t)N;
  auto clipped = [&](int n) -> float {
    float l = lrow[n];
    if (tanh_clip > 0.f) l = __fmul_rn(1.0f - __fdividef(2.0f, __expf(2.0f * l) + 1.0f), tanh_clip);
    return mrow[n] ? __fdiv_rn(l, temperature) : -INFINITY;
  };
  float mx = -INFINITY;
  for (int n = lane; n < N; n += 32) mx = fmaxf(mx, clipped(n));
  mx = warp_max(mx);
  float se = 0.f;
  for (int n = lane; n < N; n += 32) se += sfexp(clipped(n) - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  se = __logf(se);
  float best = -INFINITY;
  int besti = 0x7fffffff;
  const uint2 key2 = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (int n = lane; n < N; n += 32) {
    const float lp = __fsub_rn(__fsub_rn(clipped(n), mx), se);
    float key = lp;
    if (mode == RRNCO_DECODE_SAMPLING) {  // same stream as the fused kernel: Philox(seed; (r, step, n >> 2))[n & 3]
      const uint4 rnd = philox4x32(make_uint4((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)step, (uint32_t)(n >> 2)), key2);
      const int c = n & 3;
      const uint32_t x = c == 0 ? rnd.x : c == 1 ? rnd.y : c == 2 ? rnd.z : rnd.w;
      key = lp + (-logf(-logf(u01(x))));
    }
    if (key > best) { best = key; besti = n; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  int act = besti == 0x7fffffff ? 0 : besti;
  if (mode == RRNCO_DECODE_EVALUATE) act = min(max((int)forced[r], 0), N - 1);
  if (lane == 0) {
    if (!mrow[act]) atomicOr(status, besti == 0x7fffffff ? RRNCO_DEV_NO_FEASIBLE : RRNCO_DEV_INFEASIBLE);
    action_out[r] = act;
    logp_out[r] = __fsub_rn(__fsub_rn(clipped(act), mx), se);
  }
}

}  // namespace rrnco

using namespace rrnco;

extern "C" {

int64_t rrnco_decoder_logits_large_workspace_bytes(int64_t n_rollouts) {
  if (n_rollouts <= 0) return 0;
  return 2 * n_rollouts * (int64_t)kE * (int64_t)sizeof(float) + rrnco_pointer_ffn_workspace_bytes();
}

int rrnco_decoder_logits_large(int32_t env, int32_t n_nodes, int64_t n_inst, int32_t n_starts,
                               const rrnco_decoder_weights_t* w, const rrnco_decoder_cache_t* cache,
                               const rrnco_instance_data_t* data, const int64_t* current, const int64_t* first,
                               const uint8_t* mask, const float* ctx_state, int32_t use_placeholder, float* logits_out,
                               uint32_t* status, void* workspace, void* stream) {
  RRNCO_CHECK_ARG(w && cache && data && n_inst > 0 && n_starts > 0 && n_nodes > 1 && env >= 0 && env <= 2);
  RRNCO_CHECK_ARG(w->ffn_w1 && w->ffn_b1 && w->ffn_w2 && w->ffn_b2);
  RRNCO_CHECK_ARG(cache->glimpse_key && cache->glimpse_val && cache->logit_key && cache->ctx_node_proj);
  RRNCO_CHECK_ARG(data->data_rows > 0 && data->distance && current && mask && logits_out && status && workspace);
  RRNCO_CHECK_ARG(((uintptr_t)workspace & 15) == 0);
  RRNCO_CHECK_ARG(env == RRNCO_ENV_ATSP ? (first && cache->ctx_node_proj2) : (ctx_state && w->ctx_state_w));
  RRNCO_CHECK_ARG(env != RRNCO_ENV_RCVRPTW || data->duration);
  RRNCO_CHECK_ARG(!use_placeholder || w->ctx_placeholder_q);
  cudaStream_t st = (cudaStream_t)stream;
  StepArgs a{};
  a.env = env; a.N = n_nodes; a.n_state = env == RRNCO_ENV_ATSP ? 0 : env == RRNCO_ENV_RCVRP ? 1 : 4;
  a.use_placeholder = use_placeholder;
  a.R = n_inst * n_starts; a.n_inst = n_inst; a.data_rows = data->data_rows;
  a.K = cache->glimpse_key; a.V = cache->glimpse_val; a.Lk = cache->logit_key;
  a.P1 = cache->ctx_node_proj; a.P2 = cache->ctx_node_proj2;
  a.wstate = w->ctx_state_w; a.placeholder = w->ctx_placeholder_q;
  a.dist = data->distance; a.dur = data->duration; a.alpha = w->alpha; a.beta = w->beta;
  a.cur = current; a.first = first; a.mask = mask; a.state = ctx_state;
  float* g0 = reinterpret_cast<float*>(workspace);
  float* g1 = g0 + a.R * kE;
  void* ffn_ws = g1 + a.R * kE;
  a.logits = logits_out; a.status = status;
  const unsigned grid = (unsigned)((a.R + kStepWarps - 1) / kStepWarps);
  a.G = g0;
  attention_stream_kernel<<<grid, kStepWarps * 32, 0, st>>>(a);
  int rc = rrnco_launch_status();
  if (rc != RRNCO_OK) return rc;
  rc = rrnco_pointer_ffn(a.R, g0, w->ffn_w1, w->ffn_b1, w->ffn_w2, w->ffn_b2, g1, ffn_ws, stream);
  if (rc != RRNCO_OK) return rc;
  a.G = g1;
  logits_stream_kernel<<<grid, kStepWarps * 32, 0, st>>>(a);
  return rrnco_launch_status();
}

int rrnco_select_action(int64_t n_rollouts, int32_t n_nodes, const float* logits, const uint8_t* mask, int32_t decode_mode,
                        float tanh_clipping, float temperature, uint64_t seed, int32_t step, const int64_t* forced_action,
                        int64_t* action_out, float* logprob_out, uint32_t* status, void* stream) {
  if (n_rollouts == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(n_rollouts > 0 && n_nodes > 0 && logits && mask && action_out && logprob_out && status);
  RRNCO_CHECK_ARG(decode_mode >= 0 && decode_mode <= 2 && temperature > 0.f);
  RRNCO_CHECK_ARG(decode_mode != RRNCO_DECODE_EVALUATE || forced_action);
  select_action_kernel<<<(unsigned)((n_rollouts + kStepWarps - 1) / kStepWarps), kStepWarps * 32, 0, (cudaStream_t)stream>>>(
      n_rollouts, n_nodes, logits, mask, decode_mode, tanh_clipping, temperature, seed, step, forced_action, action_out,
      logprob_out, status);
  return rrnco_launch_status();
}

}  // extern "C"
