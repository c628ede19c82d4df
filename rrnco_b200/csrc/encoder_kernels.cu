// Encoder hot spot: the gating neural adaptive bias of an attention-free block without the
// duration channel -- DistAngleFusion.forward, rrnco/models/nn/attn_freenet.py:242-289, the variant the encoder builds for
// ATSP and RCVRP (rrnco/models/encoder.py:63-66); one module per block, 2 blocks per layer, 6 layers.
//
// Upstream materialises dist_emb = MLP(cost) and angle_emb = MLP(angle) as [B, N, N, E] tensors (two Linear(E, E) per
// ordered node pair: 65.5 kFLOP and 1 KB of activations per pair and module; 43 GB per block at B' = 8192, N = 101) and
// then keeps four scalar products of them: the gate logit  wg_d . d + wg_a . a + bg  and, because the output layer is
// linear,  wo . (g d + (1 - g) a) + bo = g (wo . d) + (1 - g)(wo . a) + bo.   The second Linear of each MLP is linear too:
//     d = W2d h_d + b2d,  h_d = relu(w1d c + b1d)   =>   wg_d . d = (W2d^T wg_d) . h_d + wg_d . b2d
// so the whole module collapses, exactly in real arithmetic, to four E-vectors u and a few constants per module
// (rrnco_nab_pack, fp64 accumulation), and a pair costs 2 x E relu-FMAs + 4 x E FMAs on the CUDA cores: nothing is
// materialised, the kernel reads the cost entry (4 B) and writes the bias (4 B).  This is NOT GEMM-shaped work (four output
// columns per pair, an A operand that would have to be generated element by element): no tensor cores.
//
// Each of the four collapsed functions  f(x) = sum_k u_k relu(w_k x + b_k)  is PIECEWISE LINEAR in its scalar argument with
// at most E breakpoints t_k = -b_k / w_k.  rrnco_nab_pack therefore also sorts the breakpoints of the two MLPs and tabulates
// slope and intercept of the gate and output functions on each of the E + 1 segments (fp64 prefix sums): the default
// variant of the kernel finds the segment of the cost entry and of the angle by binary search in shared memory and
// evaluates four FMAs -- ~100 instructions per pair instead of ~1000, which moves the kernel from the fp32 pipes toward
// the HBM bound (4 B read + 4 B written per pair).  The brute-force sum over the hidden units (variant 1) is the cross-check.
//
// Layout of `packed` (fp32): [E][8] = (w1d, b1d, ug_d, uo_d, w1a, b1a, ug_a, uo_a) per hidden unit k, 4 constants
// (gate constant, wo . b2d, wo . b2a, bo); then per MLP (dist, angle): E sorted breakpoints, (E + 1) x 4 segment
// coefficients (gate slope, gate intercept, out slope, out intercept).
#include "common.cuh"

namespace rrnco {

constexpr int kNabThreads = 256;
constexpr int kNabPairs = 4;                                  // pairs per thread: every parameter load serves four pairs
constexpr int kNabBrute = kE * 8 + 4;
constexpr int kNabTable = kE + (kE + 1) * 4;                  // sorted breakpoints + segment coefficients of one MLP
constexpr int kNabPacked = kNabBrute + 2 * kNabTable;

// one block of kE threads: thread k collapses hidden unit k
__global__ void __launch_bounds__(kE) nab_pack_kernel(const float* __restrict__ w1d, const float* __restrict__ b1d,
                                                      const float* __restrict__ W2d, const float* __restrict__ b2d,
                                                      const float* __restrict__ w1a, const float* __restrict__ b1a,
                                                      const float* __restrict__ W2a, const float* __restrict__ b2a,
                                                      const float* __restrict__ wg, const float* __restrict__ bg,
                                                      const float* __restrict__ wo, const float* __restrict__ bo,
                                                      float* __restrict__ packed) {
  const int k = threadIdx.x;
  double ugd = 0.0, uod = 0.0, uga = 0.0, uoa = 0.0;
  for (int e = 0; e < kE; ++e) {  // column k of the second-layer weights (nn.Linear stores [out, in])
    const double vd = W2d[e * kE + k], va = W2a[e * kE + k];
    ugd += vd * (double)wg[e];
    uod += vd * (double)wo[e];
    uga += va * (double)wg[kE + e];
    uoa += va * (double)wo[e];
  }
  float* q = packed + k * 8;
  q[0] = w1d[k]; q[1] = b1d[k]; q[2] = (float)ugd; q[3] = (float)uod;
  q[4] = w1a[k]; q[5] = b1a[k]; q[6] = (float)uga; q[7] = (float)uoa;
  if (k == 0) {
    double cg = bg[0], cod = 0.0, coa = 0.0;
    for (int e = 0; e < kE; ++e) {
      cg += (double)wg[e] * b2d[e] + (double)wg[kE + e] * b2a[e];
      cod += (double)wo[e] * b2d[e];
      coa += (double)wo[e] * b2a[e];
    }
    float* c = packed + kE * 8;
    c[0] = (float)cg; c[1] = (float)cod; c[2] = (float)coa; c[3] = bo[0];
  }
  // ---- piecewise-linear tables of the two MLPs ----
  __shared__ double s_t[kE], s_w[kE], s_b[kE], s_ug[kE], s_uo[kE];
  __shared__ int s_rank[kE];
#pragma unroll 1
  for (int mlp = 0; mlp < 2; ++mlp) {
    const double w = mlp ? (double)w1a[k] : (double)w1d[k], bb = mlp ? (double)b1a[k] : (double)b1d[k];
    __syncthreads();
    s_w[k] = w; s_b[k] = bb; s_ug[k] = mlp ? uga : ugd; s_uo[k] = mlp ? uoa : uod;
    s_t[k] = w != 0.0 ? -bb / w : INFINITY;  // a unit with zero slope never switches: no breakpoint
    __syncthreads();
    int r = 0;
    for (int j = 0; j < kE; ++j) r += (s_t[j] < s_t[k] || (s_t[j] == s_t[k] && j < k)) ? 1 : 0;
    s_rank[k] = r;
    float* tab = packed + kNabBrute + mlp * kNabTable;
    tab[r] = (float)s_t[k];
    __syncthreads();
    // segment s = arguments with exactly s breakpoints strictly below them; unit j is active there iff
    // (w_j > 0 and rank_j < s) or (w_j < 0 and rank_j >= s) or (w_j == 0 and b_j > 0)
    for (int sgm = k; sgm <= kE; sgm += kE) {
      double sg = 0.0, tg = 0.0, so = 0.0, to = 0.0;
      for (int j = 0; j < kE; ++j) {
        const bool on = s_w[j] > 0.0 ? s_rank[j] < sgm : (s_w[j] < 0.0 ? s_rank[j] >= sgm : s_b[j] > 0.0);
        if (on) {
          sg += s_ug[j] * s_w[j]; tg += s_ug[j] * s_b[j];
          so += s_uo[j] * s_w[j]; to += s_uo[j] * s_b[j];
        }
      }
      float* q4 = tab + kE + sgm * 4;
      q4[0] = (float)sg; q4[1] = (float)tg; q4[2] = (float)so; q4[3] = (float)to;
    }
  }
}

// grid (instances, chunks of kNabThreads * kNabPairs pairs); thread t of a chunk owns pairs t, t + 256, t + 512, t + 768 of
// the chunk (pair p = i N + j): loads and stores are coalesced across the warp.
__global__ void __launch_bounds__(kNabThreads) nab_gating_kernel(int N, const float* __restrict__ coords,
                                                                 const float* __restrict__ cost, int transpose_cost,
                                                                 const float* __restrict__ packed, float scale,
                                                                 float* __restrict__ out) {
  __shared__ __align__(16) float sp[kNabBrute];
  const int tid = threadIdx.x;
  for (int i = tid; i < kNabBrute; i += kNabThreads) sp[i] = packed[i];
  __syncthreads();
  const int64_t b = blockIdx.x;
  const int NN = N * N;
  const float* cb = cost + b * (int64_t)NN;
  const float* xy = coords + b * (int64_t)N * 2;
  float c[kNabPairs], th[kNabPairs];
  int pidx[kNabPairs];
#pragma unroll
  for (int m = 0; m < kNabPairs; ++m) {
    const int p = blockIdx.y * (kNabThreads * kNabPairs) + m * kNabThreads + tid;
    pidx[m] = p;
    c[m] = 0.f;
    th[m] = 0.f;
    if (p < NN) {
      const int i = p / N, j = p - i * N;
      c[m] = transpose_cost ? __ldg(cb + (size_t)j * N + i) : __ldg(cb + p);
      const float2 pi = __ldg(reinterpret_cast<const float2*>(xy) + i), pj = __ldg(reinterpret_cast<const float2*>(xy) + j);
      th[m] = atan2f(pi.y - pj.y, pi.x - pj.x);  // attn_freenet.py:254-262
    }
  }
  float gd[kNabPairs] = {0.f, 0.f, 0.f, 0.f}, od[kNabPairs] = {0.f, 0.f, 0.f, 0.f};
  float ga[kNabPairs] = {0.f, 0.f, 0.f, 0.f}, oa[kNabPairs] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int k = 0; k < kE; ++k) {
    const float4 pd = *reinterpret_cast<const float4*>(&sp[k * 8]);      // w1d, b1d, ug_d, uo_d (shared-memory broadcast)
    const float4 pa = *reinterpret_cast<const float4*>(&sp[k * 8 + 4]);  // w1a, b1a, ug_a, uo_a
#pragma unroll
    for (int m = 0; m < kNabPairs; ++m) {
      const float hd = fmaxf(fmaf(pd.x, c[m], pd.y), 0.f);
      gd[m] = fmaf(pd.z, hd, gd[m]);
      od[m] = fmaf(pd.w, hd, od[m]);
      const float ha = fmaxf(fmaf(pa.x, th[m], pa.y), 0.f);
      ga[m] = fmaf(pa.z, ha, ga[m]);
      oa[m] = fmaf(pa.w, ha, oa[m]);
    }
  }
  const float cg = sp[kE * 8], cod = sp[kE * 8 + 1], coa = sp[kE * 8 + 2], bo = sp[kE * 8 + 3];
#pragma unroll
  for (int m = 0; m < kNabPairs; ++m) {
    if (pidx[m] < NN) {
      const float z = gd[m] + ga[m] + cg;
      const float g = 1.0f / (1.0f + expf(-z));  // nn.Sigmoid (attn_freenet.py:239)
      const float y = g * (od[m] + cod) + (1.0f - g) * (oa[m] + coa) + bo;
      out[b * (int64_t)NN + pidx[m]] = y * scale;
    }
  }
}

// number of sorted breakpoints strictly below x (lower bound over kE = 128 entries, 8 shared-memory probes)
__device__ __forceinline__ int nab_segment(const float* __restrict__ t, float x) {
  int pos = 0;
#pragma unroll
  for (int step = kE / 2; step >= 1; step >>= 1) pos += t[pos + step - 1] < x ? step : 0;
  return pos + (t[pos] < x ? 1 : 0);
}

// Table variant: thread t of a chunk owns pairs t, t + 256, ... (coalesced loads / stores), one segment search per scalar.
__global__ void __launch_bounds__(kNabThreads) nab_gating_table_kernel(int N, const float* __restrict__ coords,
                                                                       const float* __restrict__ cost, int transpose_cost,
                                                                       const float* __restrict__ packed, float scale,
                                                                       float* __restrict__ out) {
  __shared__ __align__(16) float st[2 * kNabTable + 4];
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * kNabTable; i += kNabThreads) st[i] = packed[kNabBrute + i];
  if (tid < 4) st[2 * kNabTable + tid] = packed[kE * 8 + tid];
  __syncthreads();
  const float* td = st;                      // dist: breakpoints, then coefficients
  const float* ta = st + kNabTable;          // angle
  const float cg = st[2 * kNabTable], cod = st[2 * kNabTable + 1], coa = st[2 * kNabTable + 2], bo = st[2 * kNabTable + 3];
  const int64_t b = blockIdx.x;
  const int NN = N * N;
  const float* cb = cost + b * (int64_t)NN;
  const float2* xy = reinterpret_cast<const float2*>(coords + b * (int64_t)N * 2);
#pragma unroll
  for (int m = 0; m < kNabPairs; ++m) {
    const int p = blockIdx.y * (kNabThreads * kNabPairs) + m * kNabThreads + tid;
    if (p < NN) {
      const int i = p / N, j = p - i * N;
      const float c = transpose_cost ? __ldg(cb + (size_t)j * N + i) : __ldg(cb + p);
      const float2 pi = __ldg(xy + i), pj = __ldg(xy + j);
      const float th = atan2f(pi.y - pj.y, pi.x - pj.x);  // attn_freenet.py:254-262
      const float4 qd = *reinterpret_cast<const float4*>(td + kE + 4 * nab_segment(td, c));
      const float4 qa = *reinterpret_cast<const float4*>(ta + kE + 4 * nab_segment(ta, th));
      const float z = fmaf(qd.x, c, qd.y) + fmaf(qa.x, th, qa.y) + cg;
      const float g = 1.0f / (1.0f + expf(-z));
      const float y = g * (fmaf(qd.z, c, qd.w) + cod) + (1.0f - g) * (fmaf(qa.z, th, qa.w) + coa) + bo;
      out[b * (int64_t)NN + p] = y * scale;
    }
  }
}

}  // namespace rrnco

using namespace rrnco;

extern "C" {

int64_t rrnco_nab_packed_floats(void) { return kNabPacked; }

int rrnco_nab_pack(const float* dist_w1, const float* dist_b1, const float* dist_w2, const float* dist_b2,
                   const float* angle_w1, const float* angle_b1, const float* angle_w2, const float* angle_b2,
                   const float* gate_w, const float* gate_b, const float* out_w, const float* out_b, float* packed,
                   void* stream) {
  RRNCO_CHECK_ARG(dist_w1 && dist_b1 && dist_w2 && dist_b2 && angle_w1 && angle_b1 && angle_w2 && angle_b2 && gate_w &&
                  gate_b && out_w && out_b && packed);
  nab_pack_kernel<<<1, kE, 0, (cudaStream_t)stream>>>(dist_w1, dist_b1, dist_w2, dist_b2, angle_w1, angle_b1, angle_w2, angle_b2,
                                                      gate_w, gate_b, out_w, out_b, packed);
  return rrnco_launch_status();
}

int rrnco_nab_gating(int64_t n_inst, int32_t n_nodes, const float* coords, const float* cost, int32_t transpose_cost,
                     const float* packed, float scale, int32_t variant, float* out, void* stream) {
  RRNCO_CHECK_ARG(n_inst > 0 && n_nodes > 0 && coords && cost && packed && out);
  RRNCO_CHECK_ARG((reinterpret_cast<uintptr_t>(coords) & 7u) == 0 && (reinterpret_cast<uintptr_t>(packed) & 15u) == 0);
  const int per_cta = kNabThreads * kNabPairs;
  const int64_t chunks = ((int64_t)n_nodes * n_nodes + per_cta - 1) / per_cta;
  if (n_inst > 0x7fffffffLL || chunks > 65535) return RRNCO_ERR_UNSUPPORTED;
  const dim3 grid((unsigned)n_inst, (unsigned)chunks);
  RRNCO_CHECK_ARG(variant == 0 || variant == 1);
  if (variant == 0)
    nab_gating_table_kernel<<<grid, kNabThreads, 0, (cudaStream_t)stream>>>(n_nodes, coords, cost, transpose_cost, packed, scale, out);
  else
    nab_gating_kernel<<<grid, kNabThreads, 0, (cudaStream_t)stream>>>(n_nodes, coords, cost, transpose_cost, packed, scale, out);
  return rrnco_launch_status();
}

}  // extern "C"
