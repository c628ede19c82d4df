// Environment kernels in REFERENCE LAYOUT (flat [R, ...] td tensors): reset normalisation, instance
// gather, step + action mask for ATSP / RCVRP / RCVRPTW, tour reward.  All are HBM-bound byte / fp32
// streaming kernels: one warp per rollout (a rollout's row of N nodes is 100-400 contiguous bytes),
// lanes stride over nodes so that every global access of a warp is one contiguous segment; the
// any-feasible reduction the depot rule needs is a ballot.  fp32 arithmetic uses explicit _rn
// intrinsics in the reference's evaluation order so that masks are bit-exact (no FMA contraction).
#include <cstdlib>
#include "common.cuh"

namespace rrnco {

constexpr int kWarpsPerBlock = 8;

// ------------------------------------------------------------------------------------------------
// reset: (d - min) / (max - min + 1e-6)      rrnco/envs/rcvrp/env.py:138-145
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_minmax(float& lo, float& hi, float* red /*[64]*/) {
  lo = warp_min(lo);
  hi = warp_max(hi);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    red[warp] = lo;
    red[32 + warp] = hi;
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  lo = lane < nw ? red[lane] : INFINITY;
  hi = lane < nw ? red[32 + lane] : -INFINITY;
  lo = warp_min(lo);
  hi = warp_max(hi);
  __syncthreads();
}

__global__ void __launch_bounds__(256) minmax_normalize_kernel(int n2, const float* __restrict__ in,
                                                               float* __restrict__ out,
                                                               float* __restrict__ mn, float* __restrict__ mx) {
  __shared__ float red[64];
  const float* src = in + (size_t)blockIdx.x * n2;
  float* dst = out + (size_t)blockIdx.x * n2;
  float lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    float v = src[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  block_minmax(lo, hi, red);
  const float denom = __fadd_rn(__fsub_rn(hi, lo), 1e-6f);
  for (int i = threadIdx.x; i < n2; i += blockDim.x) dst[i] = __fdiv_rn(__fsub_rn(src[i], lo), denom);
  if (threadIdx.x == 0) {
    mn[blockIdx.x] = lo;
    mx[blockIdx.x] = hi;
  }
}

// ------------------------------------------------------------------------------------------------
// gather: out[b,i,j] = (float) M[idx[b,i], idx[b,j]]      rrnco/envs/rcvrp/sampler.py:84-90
// One CTA per instance; idx row staged in shared memory; the city matrix (fp64 8 MB, or its fp32 copy 4 MB made once
// per city by rrnco_city_matrix_to_f32) is L2 resident, writes are fully coalesced.  Optional fused reset normalisation
// (second pass hits L1/L2).  The kernel is bound by L2 -> SM sector traffic: n of the L columns of a row are wanted, so
// nearly every element costs its own 32-byte sector (ncu: profiles/r1_ncu_gather.csv); the fp32 copy packs 8 elements
// per sector instead of 4 (~70 instead of ~83 distinct sectors per row at n = 101, L = 1000).  Measured alternatives
// that were slower: a warp per row with the tile kept in shared memory between the two passes (fewer resident CTAs,
// fewer loads in flight: 0.25 ms vs 0.22 ms for 4096 instances).
// ------------------------------------------------------------------------------------------------
// kPer > 0 (normalize, n^2 <= 256 kPer): the thread's gathered values stay in registers between the min/max reduction and
// the single, normalised write - no first write and no re-read of the output - and all its loads are in flight at once.
// kPer == 0: chunks of 4 loads issued before their 4 stores (the first version had one load in flight per thread:
// long_scoreboard 18.9 per issue at 90 % occupancy, profiles/r1_ncu_env_kernels_v3.txt).  (i, j) advance incrementally:
// one integer division per thread instead of one per element.
// normalize == 1: env._reset's law for the distance matrix, (d - min) / (max - min + 1e-6)   (rcvrp/env.py:138-145)
// normalize == 2: the generators' law for the duration matrix, (d - min) / (max - min), range 0 -> 1
//                 (rmtvrp/generator_lazy.py:365-369: np.where(max - min == 0, 1, max - min))
__device__ __forceinline__ float minmax_denom(float lo, float hi, int normalize) {
  const float range = __fsub_rn(hi, lo);
  if (normalize == 2) return range == 0.f ? 1.0f : range;
  return __fadd_rn(range, 1e-6f);
}

template <typename T, int kPer>
__global__ void __launch_bounds__(256) gather_submatrix_kernel(const T* __restrict__ M, int L,
                                                               const int32_t* __restrict__ idx, int n,
                                                               float* __restrict__ out, int normalize,
                                                               float* __restrict__ mn, float* __restrict__ mx) {
  extern __shared__ int32_t sidx[];
  __shared__ float red[64];
  const size_t b = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += 256) sidx[i] = idx[b * n + i];
  __syncthreads();
  float* dst = out + b * (size_t)n * n;
  float lo = INFINITY, hi = -INFINITY;
  const int n2 = n * n;
  const int q = 256 / n, rem = 256 - q * n;
  int i = threadIdx.x / n, j = threadIdx.x - i * n;
  auto advance = [&]() {
    j += rem;
    i += q;
    if (j >= n) {
      j -= n;
      ++i;
    }
  };
  if (kPer > 0) {
    float v[kPer > 0 ? kPer : 1];
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      v[u] = 0.f;
      if (threadIdx.x + u * 256 < n2) v[u] = (float)__ldg(M + (size_t)sidx[i] * L + sidx[j]);
      advance();
    }
#pragma unroll
    for (int u = 0; u < kPer; ++u)
      if (threadIdx.x + u * 256 < n2) {
        lo = fminf(lo, v[u]);
        hi = fmaxf(hi, v[u]);
      }
    block_minmax(lo, hi, red);
    const float denom = minmax_denom(lo, hi, normalize);
#pragma unroll
    for (int u = 0; u < kPer; ++u)
      if (threadIdx.x + u * 256 < n2) dst[threadIdx.x + u * 256] = __fdiv_rn(__fsub_rn(v[u], lo), denom);
    if (threadIdx.x == 0) {
      mn[b] = lo;
      mx[b] = hi;
    }
    return;
  }
  for (int e0 = threadIdx.x; e0 < n2; e0 += 4 * 256) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = 0.f;
      if (e0 + u * 256 < n2) v[u] = (float)__ldg(M + (size_t)sidx[i] * L + sidx[j]);
      advance();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (e0 + u * 256 < n2) {
        dst[e0 + u * 256] = v[u];
        lo = fminf(lo, v[u]);
        hi = fmaxf(hi, v[u]);
      }
  }
  if (!normalize) return;
  block_minmax(lo, hi, red);  // contains __syncthreads: dst writes of this CTA are visible below
  const float denom = minmax_denom(lo, hi, normalize);
  for (int e = threadIdx.x; e < n2; e += 256) dst[e] = __fdiv_rn(__fsub_rn(dst[e], lo), denom);
  if (threadIdx.x == 0) {
    mn[b] = lo;
    mx[b] = hi;
  }
}

__global__ void __launch_bounds__(256) city_to_f32_kernel(int64_t n, const double* __restrict__ in, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (float)in[i];
}

// ------------------------------------------------------------------------------------------------
// ATSP step      rrnco/envs/atsp/env.py:79-105
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) atsp_step_kernel(
    int64_t R, int N, const int64_t* __restrict__ action, const int64_t* __restrict__ step_i, const uint8_t* mask_in,
    const int64_t* first_in, uint8_t* mask_out, int64_t* first_out, int64_t* cur_out, uint8_t* done_out) {
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const int a = (int)action[r];
  bool any = false;
  for (int n0 = 0; n0 < N; n0 += 32) {
    const int n = n0 + lane;
    bool m = false;
    if (n < N) {
      m = mask_in[r * N + n] != 0 && n != a;
      mask_out[r * N + n] = m;
    }
    any |= __any_sync(0xffffffffu, m);
  }
  if (lane == 0) {
    first_out[r] = step_i[0] == 0 ? (int64_t)a : first_in[r];
    cur_out[r] = a;
    done_out[r] = !any;  // count_nonzero(available) <= 0
  }
}

// Staged variant (un-aliased buffers): a warp owns blocks of 32 consecutive rollouts, whose mask rows (32 N bytes) are
// contiguous and 16-byte aligned as a block.  cp.async keeps kStages - 1 blocks per warp in flight; each lane clears its
// rollout's action byte in the staged copy and ORs its row (aligned words + edge bytes) for `done`; the block goes back
// as 128-bit vectors.  Same warp-autonomous persistent pipeline as rcvrp_step_vec_kernel below, no CTA barrier.
template <int kStages, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) atsp_step_vec_kernel(
    int64_t n_groups, int N, const int64_t* __restrict__ action, const int64_t* __restrict__ step_i,
    const uint8_t* __restrict__ mask_in, const int64_t* __restrict__ first_in, uint8_t* __restrict__ mask_out,
    int64_t* __restrict__ first_out, int64_t* __restrict__ cur_out, uint8_t* __restrict__ done_out) {
  extern __shared__ __align__(16) unsigned char sraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = 32 * N;  // bytes of one mask block (multiple of 16)
  unsigned char* wbase = sraw + (size_t)warp * kStages * nb;
  const int64_t stride = (int64_t)gridDim.x * kWarps;
  int64_t gi = (int64_t)blockIdx.x * kWarps + warp;
  const bool first_step = step_i[0] == 0;
  auto stage = [&](int64_t g, int b) {
    unsigned char* s = wbase + b * nb;
    const unsigned char* gm = mask_in + g * nb;
    for (int i = lane; i < nb / 16; i += 32) cp_async16(s + i * 16, gm + i * 16);
  };
#pragma unroll
  for (int k = 0; k < kStages - 1; ++k) {
    if (gi + k * stride < n_groups) stage(gi + k * stride, k);
    cp_async_commit();
  }
  int64_t a_n = 0, f_n = 0;
  if (gi < n_groups) {
    a_n = action[gi * 32 + lane];
    if (!first_step) f_n = first_in[gi * 32 + lane];
  }
  for (int b = 0; gi < n_groups; gi += stride, b = (b + 1) % kStages) {
    if (gi + (kStages - 1) * stride < n_groups) stage(gi + (kStages - 1) * stride, (b + kStages - 1) % kStages);
    cp_async_commit();
    const int64_t r = gi * 32 + lane;
    const int64_t a = a_n, f = f_n;
    if (gi + stride < n_groups) {  // next block's scalars: their latency hides under this block
      a_n = action[(gi + stride) * 32 + lane];
      if (!first_step) f_n = first_in[(gi + stride) * 32 + lane];
    }
    cp_async_wait<kStages - 1>();
    __syncwarp();
    unsigned char* s = wbase + b * nb;
    const int o = lane * N, end = o + N;
    if ((uint64_t)a < (uint64_t)N) s[o + (int)a] = 0;  // an out-of-range action must not touch a neighbour's row
    const int a0 = min((o + 3) & ~3, end), a1 = max(end & ~3, a0);
    uint32_t any = 0;
    for (int q = o; q < a0; ++q) any |= s[q];
    for (int q = a0; q < a1; q += 4) any |= *reinterpret_cast<const uint32_t*>(s + q);
    for (int q = a1; q < end; ++q) any |= s[q];
    first_out[r] = first_step ? a : f;
    cur_out[r] = a;
    done_out[r] = any == 0;  // count_nonzero(available) <= 0
    __syncwarp();
    uint4* om = reinterpret_cast<uint4*>(mask_out + gi * nb);
    for (int i = lane; i < nb / 16; i += 32) om[i] = reinterpret_cast<const uint4*>(s)[i];
    __syncwarp();  // the buffer is re-staged kStages - 1 iterations from now, after this warp's own reads
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// RCVRP step + mask      rrnco/envs/rcvrp/env.py:90-122, 183-195
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) rcvrp_step_kernel(
    int64_t R, int N, int64_t data_rows, const int64_t* __restrict__ action, const float* __restrict__ demand,
    const float* __restrict__ capacity, int64_t cap_rows, const float* used_in, const uint8_t* visited_in,
    const int64_t* current_in, float* used_out, uint8_t* visited_out, int64_t* current_out,
    uint8_t* done_out, uint8_t* mask_out) {
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const float* dem = demand + (r % data_rows) * (int64_t)(N - 1);
  const float cap = capacity[r % cap_rows];
  float used = used_in[r];
  int cur;
  if (action != nullptr) {
    cur = (int)action[r];
    const int di = min(max(cur - 1, 0), N - 2);  // clamp(a - 1, 0, n_loc - 1)
    used = __fmul_rn(__fadd_rn(used, dem[di]), cur != 0 ? 1.0f : 0.0f);
  } else {
    cur = (int)current_in[r];
  }
  int n_visited = 0;
  bool any_free = false;
  for (int n0 = 0; n0 < N; n0 += 32) {
    const int n = n0 + lane;
    bool v = false, free_loc = false;
    if (n < N) {
      uint8_t vb = visited_in[r * N + n];
      if (action != nullptr) {
        if (n == cur) vb = 1;  // scatter(-1, cur, 1)
        visited_out[r * N + n] = vb;
      }
      v = vb != 0;
      if (n >= 1) {
        const bool exceeds = __fadd_rn(dem[n - 1], used) > cap;
        free_loc = !(v || exceeds);
        mask_out[r * N + n] = free_loc;
      }
    }
    n_visited += __popc(__ballot_sync(0xffffffffu, v));
    any_free |= __any_sync(0xffffffffu, free_loc);
  }
  if (lane == 0) {
    mask_out[r * N] = !(cur == 0 && any_free);
    if (action != nullptr) {
      used_out[r] = used;
      current_out[r] = cur;
      done_out[r] = n_visited == N;
    }
  }
}

// Staged variant for the un-aliased, fully replicated reference layout (data_rows == R, cap_rows == R or 1):
// blocks of 32 consecutive rollouts, whose visited / mask rows (32 N bytes) and demand rows (32 (N-1) floats) are
// contiguous and 16-byte aligned as a block, so every global access is a coalesced 128-bit vector; rows are processed
// from shared memory and written back the same way.
constexpr int kGroup = 32;
// Persistent, warp-autonomous pipeline.  ncu of the first staged version (one warp per rollout over the staged bytes,
// lanes = nodes, load -> barrier -> compute -> barrier -> store per CTA) showed 85 % issue-slot utilisation at 31 % DRAM
// utilisation: the reference's bool [R, N] rows are byte-granular and unaligned, so the per-node work was scalar byte
// traffic plus two ballots per 32 nodes.  Here a warp owns blocks of 32 consecutive rollouts (contiguous in every
// array): cp.async stages block k+1 into the second half of the warp's private buffer while the 32 lanes each walk one
// rollout of block k (no cross-lane operations, ~8 instructions per node for 32 rollouts at once, 128-bit demand
// loads), then the warp writes the visited / mask blocks back as 128-bit vectors.  No CTA-wide barrier.
// Inside a block every lane walks its own rollout in chunks of 20 nodes held in registers: the visited bytes and the
// five demand float4 of chunk c+1 are loaded before the mask bytes of chunk c are stored.  (The first version
// interleaved one LDS.U8 / STS.U8 pair per node: the mask stores may alias the visited loads as far as the compiler
// can tell, so every node paid a full shared-memory round trip, ~5 k cycles per block and warp; cuobjdump -sass.)
// The per-rollout scalars (action, capacity, used) of block k+1 are fetched while block k is walked.
constexpr int kStepStagesDefault = 2;  // staged [visited | demand] buffers per warp (prefetch distance stages - 1 blocks)
constexpr int kStepWarpsDefault = 6;   // 6 warps x (2 x 16.0 KB staged visited | demand + 3.2 KB mask) (N = 101) = 212 KB: one CTA per SM
constexpr int kChunkQ = 5;             // float4 of demand per register chunk (20 nodes)
template <int kStepStages, int kStepWarps>
__global__ void __launch_bounds__(kStepWarps * 32, 1) rcvrp_step_vec_kernel(
    int64_t n_groups, int N, const int64_t* __restrict__ action, const float* __restrict__ demand,
    const float* __restrict__ capacity, int64_t cap_rows, const float* __restrict__ used_in,
    const uint8_t* __restrict__ visited_in, float* __restrict__ used_out, uint8_t* __restrict__ visited_out,
    int64_t* __restrict__ current_out, uint8_t* __restrict__ done_out, uint8_t* __restrict__ mask_out) {
  extern __shared__ __align__(16) unsigned char sraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = kGroup * N;                                  // bytes of one visited / mask block (multiple of 16)
  const int nd = kGroup * (N - 1) * (int)sizeof(float);       // bytes of one demand block (multiple of 16)
  const int buf_bytes = nb + nd;                               // one staged buffer: [visited | demand]
  unsigned char* wbase = sraw + (size_t)warp * (kStepStages * buf_bytes + nb);  // this warp: staged buffers + one mask block
  unsigned char* s_mask = wbase + kStepStages * buf_bytes;
  const int64_t stride = (int64_t)gridDim.x * kStepWarps;
  int64_t gi = (int64_t)blockIdx.x * kStepWarps + warp;
  auto stage = [&](int64_t g, int b) {
    unsigned char* s = wbase + b * buf_bytes;
    const unsigned char* gv = visited_in + g * nb;
    for (int i = lane; i < nb / 16; i += 32) cp_async16(s + i * 16, gv + i * 16);
    const unsigned char* gd = reinterpret_cast<const unsigned char*>(demand) + g * nd;
    for (int i = lane; i < nd / 16; i += 32) cp_async16(s + nb + i * 16, gd + i * 16);
  };
#pragma unroll
  for (int k = 0; k < kStepStages - 1; ++k) {
    if (gi + k * stride < n_groups) stage(gi + k * stride, k);
    cp_async_commit();
  }
  int cur_n = 0;
  float cap_n = 0.0f, used_n = 0.0f;
  if (gi < n_groups) {
    const int64_t r0 = gi * kGroup + lane;
    cur_n = (int)action[r0];
    cap_n = capacity[r0 % cap_rows];
    used_n = used_in[r0];
  }
  for (int b = 0; gi < n_groups; gi += stride, b = (b + 1) % kStepStages) {
    if (gi + (kStepStages - 1) * stride < n_groups) stage(gi + (kStepStages - 1) * stride, (b + kStepStages - 1) % kStepStages);
    cp_async_commit();
    const int64_t r = gi * kGroup + lane;
    const int cur = cur_n;
    const float cap = cap_n, used0 = used_n;
    if (gi + stride < n_groups) {  // next block's scalars: their latency hides under this block's walk
      const int64_t rn = (gi + stride) * kGroup + lane;
      cur_n = (int)action[rn];
      cap_n = capacity[rn % cap_rows];
      used_n = used_in[rn];
    }
    cp_async_wait<kStepStages - 1>();
    __syncwarp();
    unsigned char* s = wbase + b * buf_bytes;
    {
      uint8_t* vis = s + lane * N;
      uint8_t* msk = s_mask + lane * N;
      const float* dem = reinterpret_cast<const float*>(s + nb) + lane * (N - 1);
      const int di = min(max(cur - 1, 0), N - 2);
      const float used = __fmul_rn(__fadd_rn(used0, dem[di]), cur != 0 ? 1.0f : 0.0f);
      if ((unsigned)cur < (unsigned)N) vis[cur] = 1;  // an out-of-range action must not touch a neighbour's row
      int n_visited = vis[0];
      bool any_free = false;
      // demand rows are 16-byte aligned whenever (N - 1) % 4 == 0 (e.g. N = 101): 128-bit shared loads, conflict-free at
      // a lane stride of (N - 1) floats; the byte rows of visited / mask stay scalar
      const int n4 = ((N - 1) & 3) == 0 ? (N - 1) >> 2 : 0;
      const int n_chunks = n4 / kChunkQ;
      const float4* dem4 = reinterpret_cast<const float4*>(dem);
      float4 dq[kChunkQ];
      uint32_t vq[4 * kChunkQ];
      auto load_chunk = [&](int c) {
#pragma unroll
        for (int q = 0; q < kChunkQ; ++q) dq[q] = dem4[c * kChunkQ + q];
#pragma unroll
        for (int e = 0; e < 4 * kChunkQ; ++e) vq[e] = vis[c * 4 * kChunkQ + 1 + e];
      };
      if (n_chunks > 0) load_chunk(0);
      for (int c = 0; c < n_chunks; ++c) {
        float dd[4 * kChunkQ];
        uint32_t vv[4 * kChunkQ];
#pragma unroll
        for (int q = 0; q < kChunkQ; ++q) {
          dd[4 * q] = dq[q].x, dd[4 * q + 1] = dq[q].y, dd[4 * q + 2] = dq[q].z, dd[4 * q + 3] = dq[q].w;
        }
#pragma unroll
        for (int e = 0; e < 4 * kChunkQ; ++e) vv[e] = vq[e];
        if (c + 1 < n_chunks) load_chunk(c + 1);  // in flight before this chunk's mask stores
        uint8_t* mc = msk + c * 4 * kChunkQ + 1;
#pragma unroll
        for (int e = 0; e < 4 * kChunkQ; ++e) {
          const bool v = vv[e] != 0;
          const bool free_loc = !(v || __fadd_rn(dd[e], used) > cap);
          mc[e] = free_loc;
          n_visited += v;
          any_free |= free_loc;
        }
      }
      for (int n = 4 * kChunkQ * n_chunks + 1; n < N; ++n) {
        const bool v = vis[n] != 0;
        const bool free_loc = !(v || __fadd_rn(dem[n - 1], used) > cap);
        msk[n] = free_loc;
        n_visited += v;
        any_free |= free_loc;
      }
      msk[0] = !(cur == 0 && any_free);
      used_out[r] = used;
      current_out[r] = cur;
      done_out[r] = n_visited == N;
    }
    __syncwarp();
    uint4* ov = reinterpret_cast<uint4*>(visited_out + gi * nb);
    uint4* om = reinterpret_cast<uint4*>(mask_out + gi * nb);
    for (int i = lane; i < nb / 16; i += 32) {
      ov[i] = reinterpret_cast<const uint4*>(s)[i];
      om[i] = reinterpret_cast<const uint4*>(s_mask)[i];
    }
    __syncwarp();  // the buffer is re-staged kStepStages - 1 iterations from now, after this warp's own reads
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// RCVRPTW (RMTVRP) step + mask      rrnco/envs/rmtvrp/env.py:155-215, 343-428
// ------------------------------------------------------------------------------------------------
struct RmtvrpState {
  int64_t* cur;
  float *time, *route, *used_l, *used_b;
  uint8_t* visited;
};

// kShared: the CTA's warps are rollouts (starts) of ONE instance row (R = data_rows x starts, rollout r = s * data_rows +
// row): the per-instance rows (time windows, service, demands) and column 0 of both matrices - 101 strided 32-byte
// sectors each when read per rollout - are staged once per CTA in shared memory and shared by its warps.
// kIters > 0 (N <= 32 kIters): the node loops are unrolled and every global load of the rollout (visited bytes, the two
// matrix rows of the new node) is issued before the first use, so one memory round trip covers them all; the visited
// bytes stay in registers between the two passes (ncu of the rolled version: long_scoreboard was the top stall).
template <bool kShared, int kIters>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) rmtvrp_step_kernel(
    int64_t R, int N, rrnco_instance_data_t d, const int64_t* __restrict__ action, RmtvrpState in,
    RmtvrpState out, uint8_t* done_out, uint8_t* mask_out) {
  extern __shared__ __align__(16) float s_inst[];  // kShared: [col0 D | col0 Dur | tw (2N) | svc | dl | db]
  int64_t r, row;
  if (kShared) {
    row = blockIdx.x % d.data_rows;
    const int64_t s = (blockIdx.x / d.data_rows) * kWarpsPerBlock + (threadIdx.x >> 5);
    r = s * d.data_rows + row;
  } else {
    r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    row = r % d.data_rows;
  }
  const float* D = d.distance + row * (int64_t)N * N;
  const float* U = d.duration + row * (int64_t)N * N;
  const float* tw = d.time_windows + row * (int64_t)N * 2;
  const float* svc = d.service_time + row * (int64_t)N;
  const float* dl = d.demand + row * (int64_t)N;
  const float* db = d.demand_backhaul + row * (int64_t)N;
  if (kShared) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      s_inst[n] = D[(int64_t)n * N];
      s_inst[N + n] = U[(int64_t)n * N];
      s_inst[2 * N + 2 * n] = tw[2 * n];
      s_inst[2 * N + 2 * n + 1] = tw[2 * n + 1];
      s_inst[4 * N + n] = svc[n];
      s_inst[5 * N + n] = dl[n];
      s_inst[6 * N + n] = db[n];
    }
    __syncthreads();
    tw = s_inst + 2 * N, svc = s_inst + 4 * N, dl = s_inst + 5 * N, db = s_inst + 6 * N;
  }
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const float cap = d.vehicle_capacity[row];
  const float limit = d.distance_limit[row];
  const float closed = d.open_route[row] ? 0.0f : 1.0f;  // "* ~open_route" (bool -> 0/1)
  const float bclass = d.backhaul_class[row];

  int cur = (int)in.cur[r];
  const int prev = cur;
  if (action != nullptr) cur = (int)action[r];
  float time = in.time[r], route = in.route[r], used_l = in.used_l[r], used_b = in.used_b[r];
  // rollout-private loads of the unrolled form, all in flight together
  uint8_t vreg[kIters > 0 ? kIters : 1];
  float dreg[kIters > 0 ? kIters : 1], ureg[kIters > 0 ? kIters : 1];
  if (kIters > 0) {
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int n = it * 32 + lane;
      vreg[it] = 0, dreg[it] = 0.f, ureg[it] = 0.f;
      if (n < N) {
        vreg[it] = in.visited[r * N + n];
        dreg[it] = D[cur * N + n];
        ureg[it] = U[cur * N + n];
      }
    }
  }
  if (action != nullptr) {
    const float away = cur != 0 ? 1.0f : 0.0f;
    const float dist = D[prev * N + cur], dur = U[prev * N + cur];
    time = __fmul_rn(away, __fadd_rn(fmaxf(__fadd_rn(time, dur), tw[cur * 2]), svc[cur]));
    route = __fmul_rn(away, __fadd_rn(route, dist));
    used_l = __fmul_rn(away, __fadd_rn(used_l, dl[cur]));
    used_b = __fmul_rn(away, __fadd_rn(used_b, db[cur]));
  }
  // pass 1: visited update, done, linehauls_missing
  int n_visited = 0;
  bool missing = false;
  auto pass1 = [&](int n, uint8_t vb) -> uint8_t {
    bool v = false, miss = false;
    if (n < N) {
      if (action != nullptr) {
        if (n == cur) vb = 1;
        out.visited[r * N + n] = vb;
      }
      v = vb != 0;
      miss = !v && dl[n] > 0.0f;  // (demand_linehaul * ~visited).sum(-1) > 0 with non-negative demands
    }
    n_visited += __popc(__ballot_sync(0xffffffffu, v));
    missing |= __any_sync(0xffffffffu, miss);
    return vb;
  };
  if (kIters > 0) {
#pragma unroll
    for (int it = 0; it < kIters; ++it) vreg[it] = pass1(it * 32 + lane, vreg[it]);
  } else {
    for (int n0 = 0; n0 < N; n0 += 32) {
      const int n = n0 + lane;
      pass1(n, n < N ? in.visited[r * N + n] : (uint8_t)0);
    }
    __syncwarp();  // out.visited (may alias in.visited) is re-read below by the same lanes only
  }
  const bool carrying_b = db[cur] > 0.0f;
  const float late0 = tw[1];
  bool any_cust = false;
  auto pass2 = [&](int n, bool v, float dist_ij, float dur_ij) {
    bool can = false;
    if (n < N) {
      const float dist_j0 = kShared ? s_inst[n] : D[n * N];
      const float dur_j0 = kShared ? s_inst[N + n] : U[n * N];
      const float early = tw[n * 2], late = tw[n * 2 + 1];
      const float arrival = __fadd_rn(time, dur_ij);
      const bool reach_c = arrival < late;
      const bool reach_d =
          __fmul_rn(__fadd_rn(__fadd_rn(fmaxf(arrival, early), svc[n]), dur_j0), closed) < late0;
      const bool exc_lim = __fadd_rn(__fadd_rn(route, dist_ij), __fmul_rn(dist_j0, closed)) > limit;
      const bool exc_l = __fadd_rn(dl[n], used_l) > cap;
      const bool exc_b = __fadd_rn(db[n], used_b) > cap;
      const bool ok1 = (missing && !exc_l && !carrying_b && dl[n] > 0.0f) || (!exc_b && db[n] > 0.0f);
      const bool cannot_l = dl[n] > __fsub_rn(cap, used_b);
      const bool ok2 = !exc_l && !exc_b && !cannot_l;
      const bool ok = (bclass == 1.0f && ok1) || (bclass == 2.0f && ok2);
      can = reach_c && reach_d && ok && !exc_lim && !v;
      if (n >= 1) mask_out[r * N + n] = can;
    }
    any_cust |= __any_sync(0xffffffffu, can && n >= 1);
  };
  if (kIters > 0) {
#pragma unroll
    for (int it = 0; it < kIters; ++it) pass2(it * 32 + lane, vreg[it] != 0, dreg[it], ureg[it]);
  } else {
    for (int n0 = 0; n0 < N; n0 += 32) {
      const int n = n0 + lane;
      const bool inb = n < N;
      const bool v = inb && (action != nullptr ? out.visited[r * N + n] : in.visited[r * N + n]) != 0;
      pass2(n, v, inb ? D[cur * N + n] : 0.f, inb ? U[cur * N + n] : 0.f);
    }
  }
  if (lane == 0) {
    mask_out[r * N] = !(cur == 0 && any_cust);
    if (action != nullptr) {
      out.cur[r] = cur;
      out.time[r] = time;
      out.route[r] = route;
      out.used_l[r] = used_l;
      out.used_b[r] = used_b;
      done_out[r] = n_visited == N;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tour reward      atsp/env.py:192-211, rcvrp/env.py:197-219, rmtvrp/env.py:430-455
// warp per rollout, lanes over legs, fp64 accumulation (<= 0.5 ulp of the exact fp32-leg sum)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) tour_reward_kernel(
    int64_t R, int T, int N, int64_t data_rows, const int64_t* __restrict__ actions,
    const float* __restrict__ distance, int prepend_depot, const uint8_t* __restrict__ open_route,
    const float* __restrict__ mn, const float* __restrict__ mx, float* norm_out, float* real_out) {
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const int64_t row = r % data_rows;
  const float* D = distance + row * (int64_t)N * N;
  const int64_t* a = actions + r * (int64_t)T;
  const bool open = open_route != nullptr && open_route[row] != 0;
  const int legs = prepend_depot ? T + 1 : T;
  double acc = 0.0;
  for (int l = lane; l < legs; l += 32) {
    int from, to;
    if (prepend_depot) {  // go_from = [0, a...], go_to = roll(go_from, -1)
      from = l == 0 ? 0 : (int)a[l - 1];
      to = l == T ? 0 : (int)a[l];
    } else {
      from = (int)a[l];
      to = (int)a[l + 1 == T ? 0 : l + 1];
    }
    float c = D[from * N + to];
    if (open && to == 0) c = __fmul_rn(c, 0.0f);  // column 0 zeroed for open routes (rmtvrp/env.py:433)
    acc += (double)c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float neg = -(float)acc;
    norm_out[r] = neg;
    if (real_out != nullptr)
      real_out[r] = __fadd_rn(__fmul_rn(neg, __fadd_rn(__fsub_rn(mx[row], mn[row]), 1e-6f)), mn[row]);
  }
}

}  // namespace rrnco

using namespace rrnco;

static inline unsigned warp_grid(int64_t R) { return (unsigned)((R + kWarpsPerBlock - 1) / kWarpsPerBlock); }

template <typename T>
static int gather_launch(const T* city_matrix, int32_t city_len, const int32_t* idx, int64_t batch, int32_t n, float* out,
                         int32_t normalize, float* min_out, float* max_out, void* stream) {
  if (batch == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(batch > 0 && n > 0 && city_len >= n && city_matrix && idx && out);
  RRNCO_CHECK_ARG(normalize >= 0 && normalize <= 2);
  RRNCO_CHECK_ARG(!normalize || (min_out && max_out));
  if ((size_t)n * sizeof(int32_t) > 48 * 1024) return RRNCO_ERR_UNSUPPORTED;
  constexpr int kPer = 40;  // values per thread kept in registers: n <= 101
  if (normalize && (int64_t)n * n <= (int64_t)kPer * 256)
    gather_submatrix_kernel<T, kPer><<<(unsigned)batch, 256, n * sizeof(int32_t), (cudaStream_t)stream>>>(
        city_matrix, city_len, idx, n, out, normalize, min_out, max_out);
  else
    gather_submatrix_kernel<T, 0><<<(unsigned)batch, 256, n * sizeof(int32_t), (cudaStream_t)stream>>>(
        city_matrix, city_len, idx, n, out, normalize, min_out, max_out);
  return rrnco_launch_status();
}

extern "C" {

int rrnco_abi_version(void) { return RRNCO_ABI_VERSION; }
float rrnco_u01(uint32_t x) { return rrnco::u01(x); }

const char* rrnco_strerror(int code) {
  switch (code) {
    case RRNCO_OK: return "ok";
    case RRNCO_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or misaligned buffer)";
    case RRNCO_ERR_UNSUPPORTED: return "unsupported variant or size for this build";
    case RRNCO_ERR_CUDA: return "CUDA launch failed";
    default: return "unknown rrnco status";
  }
}

int rrnco_minmax_normalize(int64_t n_mat, int32_t n_nodes, const float* dist_in, float* dist_out,
                           float* min_out, float* max_out, void* stream) {
  if (n_mat == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(n_mat > 0 && n_nodes > 0 && dist_in && dist_out && min_out && max_out);
  minmax_normalize_kernel<<<(unsigned)n_mat, 256, 0, (cudaStream_t)stream>>>(n_nodes * n_nodes, dist_in, dist_out,
                                                                          min_out, max_out);
  return rrnco_launch_status();
}

int rrnco_gather_submatrix(const double* city_matrix, int32_t city_len, const int32_t* idx, int64_t batch,
                           int32_t n, float* out, int32_t normalize, float* min_out, float* max_out,
                           void* stream) {
  return gather_launch(city_matrix, city_len, idx, batch, n, out, normalize, min_out, max_out, stream);
}

int rrnco_gather_submatrix_f32(const float* city_matrix_f32, int32_t city_len, const int32_t* idx, int64_t batch,
                               int32_t n, float* out, int32_t normalize, float* min_out, float* max_out,
                               void* stream) {
  return gather_launch(city_matrix_f32, city_len, idx, batch, n, out, normalize, min_out, max_out, stream);
}

int rrnco_city_matrix_to_f32(const double* city_matrix, int64_t n_elems, float* out, void* stream) {
  if (n_elems == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(n_elems > 0 && city_matrix && out);
  const int64_t blocks = (n_elems + 255) / 256;
  city_to_f32_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, (cudaStream_t)stream>>>(n_elems, city_matrix, out);
  return rrnco_launch_status();
}

int rrnco_atsp_step(int64_t R, int32_t n_nodes, const int64_t* action, const int64_t* step_i, const uint8_t* mask_in,
                    const int64_t* first_in, uint8_t* mask_out, int64_t* first_out, int64_t* current_out,
                    uint8_t* done_out, void* stream) {
  if (R == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(R > 0 && n_nodes > 0 && action && step_i && mask_in && first_in && mask_out && first_out &&
                  current_out && done_out);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  struct VecCfg {
    int stages, warps;
    decltype(&atsp_step_vec_kernel<4, 16>) fn;
  };
  // deepest pipeline whose per-CTA staging (warps x stages x 32 N bytes) fits; one CTA per SM
  static const VecCfg cfgs[] = {{4, 16, atsp_step_vec_kernel<4, 16>}, {2, 16, atsp_step_vec_kernel<2, 16>},
                                {2, 8, atsp_step_vec_kernel<2, 8>}, {2, 4, atsp_step_vec_kernel<2, 4>},
                                {2, 2, atsp_step_vec_kernel<2, 2>}};
  const VecCfg* cfg = nullptr;
  for (const VecCfg& c : cfgs)
    if (!cfg && (size_t)c.warps * c.stages * 32 * n_nodes <= 200 * 1024) cfg = &c;
  const bool vec_ok = cfg && R >= 32 && mask_in != mask_out && al16(mask_in) && al16(mask_out);
  const int64_t R_vec = vec_ok ? (R / 32) * 32 : 0;
  if (R_vec > 0) {
    static PerDeviceOnce once;  // per device ordinal; idempotent, benign if raced
    const int n_sm = device_sm_count();
    if (n_sm <= 0) return RRNCO_ERR_CUDA;
    if (once.first()) {
      for (const VecCfg& c : cfgs)
        if (cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
          once.undo();
          return RRNCO_ERR_CUDA;
        }
    }
    const int64_t n_groups = R_vec / 32;
    const int64_t want = (n_groups + cfg->warps - 1) / cfg->warps;
    const unsigned grid = (unsigned)(want < n_sm ? want : n_sm);
    cfg->fn<<<grid, cfg->warps * 32, (size_t)cfg->warps * cfg->stages * 32 * n_nodes, (cudaStream_t)stream>>>(
        n_groups, n_nodes, action, step_i, mask_in, first_in, mask_out, first_out, current_out, done_out);
    int rc = rrnco_launch_status();
    if (rc != RRNCO_OK) return rc;
  }
  if (R_vec < R) {  // tail, or aliased / unaligned buffers: one warp per rollout
    const int64_t off = R_vec, Rt = R - R_vec;
    atsp_step_kernel<<<warp_grid(Rt), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        Rt, n_nodes, action + off, step_i, mask_in + off * n_nodes, first_in + off, mask_out + off * n_nodes,
        first_out + off, current_out + off, done_out + off);
  }
  return rrnco_launch_status();
}

int rrnco_rcvrp_step(int64_t R, int32_t n_nodes, int64_t data_rows, const int64_t* action, const float* demand,
                     const float* capacity, int64_t cap_rows, const float* used_in, const uint8_t* visited_in,
                     const int64_t* current_in, float* used_out, uint8_t* visited_out, int64_t* current_out,
                     uint8_t* done_out, uint8_t* mask_out, void* stream) {
  if (R == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(R > 0 && n_nodes > 1 && data_rows > 0 && cap_rows > 0 && demand && capacity && used_in &&
                  visited_in && mask_out);
  RRNCO_CHECK_ARG(action ? (used_out && visited_out && current_out && done_out) : (current_in != nullptr));
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  // staged persistent kernel: each warp pipelines blocks of kGroup rollouts through shared memory
  struct VecCfg {
    int stages, warps;
    decltype(&rcvrp_step_vec_kernel<2, 6>) fn;
  };
  static const VecCfg cfgs[] = {{2, 6, rcvrp_step_vec_kernel<2, 6>}, {3, 4, rcvrp_step_vec_kernel<3, 4>},
                                {4, 3, rcvrp_step_vec_kernel<4, 3>}, {2, 5, rcvrp_step_vec_kernel<2, 5>},
                                {3, 3, rcvrp_step_vec_kernel<3, 3>}, {2, 4, rcvrp_step_vec_kernel<2, 4>}};
  static int cfg_i = -1;  // idempotent; benign if raced
  if (cfg_i < 0) {
    int pick = 0;  // development knob: RRNCO_STEP_CFG="<stages>x<warps>" selects another instantiation
    if (const char* e = std::getenv("RRNCO_STEP_CFG"))
      for (int i = 0; i < (int)(sizeof(cfgs) / sizeof(cfgs[0])); ++i)
        if (e[0] - '0' == cfgs[i].stages && e[1] == 'x' && e[2] - '0' == cfgs[i].warps) pick = i;
    cfg_i = pick;
  }
  const VecCfg& cfg = cfgs[cfg_i];
  const size_t nbk = (size_t)kGroup * n_nodes, buf = nbk + (size_t)kGroup * (n_nodes - 1) * sizeof(float);
  const size_t smem = (size_t)cfg.warps * (cfg.stages * buf + nbk);
  const bool vec_ok = action != nullptr && data_rows == R && (cap_rows == 1 || cap_rows == R) && R >= kGroup &&
                      smem <= 220 * 1024 && al16(demand) && al16(visited_in) && al16(visited_out) && al16(mask_out) &&
                      visited_in != visited_out;
  const int64_t R_vec = vec_ok ? (R / kGroup) * kGroup : 0;
  if (R_vec > 0) {
    static PerDeviceOnce once;  // per device ordinal; idempotent, benign if raced
    const int n_sm = device_sm_count();
    if (n_sm <= 0) return RRNCO_ERR_CUDA;
    if (once.first()) {
      if (cudaFuncSetAttribute(cfg.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
        once.undo();
        return RRNCO_ERR_CUDA;
      }
    }
    const int64_t n_groups = R_vec / kGroup;
    const int64_t want = (n_groups + cfg.warps - 1) / cfg.warps;
    const unsigned grid = (unsigned)(want < n_sm ? want : n_sm);
    cfg.fn<<<grid, cfg.warps * 32, smem, (cudaStream_t)stream>>>(
        n_groups, n_nodes, action, demand, capacity, cap_rows, used_in, visited_in, used_out, visited_out, current_out,
        done_out, mask_out);
    int rc = rrnco_launch_status();
    if (rc != RRNCO_OK) return rc;
  }
  if (R_vec < R) {  // tail (or the general data_rows / mask-only / aliased case): one warp per rollout
    const int64_t off = R_vec, Rt = R - R_vec;
    rcvrp_step_kernel<<<warp_grid(Rt), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        Rt, n_nodes, data_rows, action ? action + off : nullptr, vec_ok ? demand + off * (n_nodes - 1) : demand,
        (vec_ok && cap_rows == R) ? capacity + off : capacity, (vec_ok && cap_rows == R) ? Rt : cap_rows, used_in + off, visited_in + off * n_nodes, current_in ? current_in + off : nullptr,
        used_out ? used_out + off : nullptr, visited_out ? visited_out + off * n_nodes : nullptr,
        current_out ? current_out + off : nullptr, done_out ? done_out + off : nullptr, mask_out + off * n_nodes);
  }
  return rrnco_launch_status();
}

int rrnco_rmtvrp_step(int64_t R, int32_t n_nodes, const rrnco_instance_data_t* data, const int64_t* action,
                      const rrnco_rmtvrp_state_t* state_in, const rrnco_rmtvrp_state_t* state_out,
                      uint8_t* done_out, uint8_t* mask_out, void* stream) {
  if (R == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(R > 0 && n_nodes > 1 && data && state_in && mask_out);
  RRNCO_CHECK_ARG(data->data_rows > 0 && data->distance && data->duration && data->time_windows &&
                  data->service_time && data->demand && data->demand_backhaul && data->vehicle_capacity &&
                  data->distance_limit && data->open_route && data->backhaul_class);
  RRNCO_CHECK_ARG(state_in->current_node && state_in->current_time && state_in->current_route_length &&
                  state_in->used_capacity_linehaul && state_in->used_capacity_backhaul && state_in->visited);
  RmtvrpState in{state_in->current_node, state_in->current_time, state_in->current_route_length,
                 state_in->used_capacity_linehaul, state_in->used_capacity_backhaul, state_in->visited};
  RmtvrpState out = in;
  if (action) {
    RRNCO_CHECK_ARG(state_out && done_out && state_out->current_node && state_out->current_time &&
                    state_out->current_route_length && state_out->used_capacity_linehaul &&
                    state_out->used_capacity_backhaul && state_out->visited);
    out = RmtvrpState{state_out->current_node, state_out->current_time, state_out->current_route_length,
                      state_out->used_capacity_linehaul, state_out->used_capacity_backhaul, state_out->visited};
  }
  const size_t smem = (size_t)7 * n_nodes * sizeof(float);
  const bool unrolled = n_nodes <= 128;
  if (R % data->data_rows == 0 && R / data->data_rows >= 2 && smem <= 48 * 1024) {
    // several rollouts per instance row: CTAs of kWarpsPerBlock starts of one row share its staged per-instance data
    const int64_t starts = R / data->data_rows;
    const int64_t groups = (starts + kWarpsPerBlock - 1) / kWarpsPerBlock;
    auto fn = unrolled ? rmtvrp_step_kernel<true, 4> : rmtvrp_step_kernel<true, 0>;
    fn<<<(unsigned)(data->data_rows * groups), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
        R, n_nodes, *data, action, in, out, done_out, mask_out);
  } else {
    auto fn = unrolled ? rmtvrp_step_kernel<false, 4> : rmtvrp_step_kernel<false, 0>;
    fn<<<warp_grid(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(R, n_nodes, *data, action, in, out, done_out,
                                                                     mask_out);
  }
  return rrnco_launch_status();
}

int rrnco_tour_reward(int64_t R, int32_t T, int32_t n_nodes, int64_t data_rows, const int64_t* actions,
                      const float* distance, int32_t prepend_depot, const uint8_t* open_route, const float* min_d,
                      const float* max_d, float* norm_out, float* real_out, void* stream) {
  if (R == 0) return RRNCO_OK;
  RRNCO_CHECK_ARG(R > 0 && T > 0 && n_nodes > 0 && data_rows > 0 && actions && distance && norm_out);
  RRNCO_CHECK_ARG(real_out == nullptr || (min_d && max_d));
  tour_reward_kernel<<<warp_grid(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      R, T, n_nodes, data_rows, actions, distance, prepend_depot, open_route, min_d, max_d, norm_out, real_out);
  return rrnco_launch_status();
}

}  // extern "C"
