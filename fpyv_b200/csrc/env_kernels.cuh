// Multi-agent gate-race environment kernels (see include/fpv_api.h "Multi-agent gate-race environment").
// One thread per agent; the agents of an env are an aligned group of A = agents_per_env lanes of one warp, so the
// per-env reward sum and termination flag are warp-shuffle / vote reductions -- the only place warp primitives are
// used on the path (north_star).  Reward definition: OURS (parity unpinned, no reference implementation).
#pragma once
#include "../../include/fpv_api.h"
#include "misc_kernels.cuh"
#include "drone_kernels.cuh"

namespace fpv {

// (explicit fused multiply-adds in a fixed order: the same source must give the same bits in every kernel it is inlined in)
__device__ __forceinline__ void gate_metrics(const fpv_gate_t& g, float px, float py, float pz, float& d, float& r) {
  const float dx = px - g.cx, dy = py - g.cy, dz = pz - g.cz;
  d = __fmaf_rn(g.nx, dx, __fmaf_rn(g.ny, dy, g.nz * dz));  // Gate.calculate_distance, components.py:819-822
  r = sqrtf(__fmaf_rn(dx, dx, __fmaf_rn(dy, dy, dz * dz)));
}

// Observation of one agent: [0:3] R^T (c_g - p)  [3:6] R^T n_g  [6:9] R^T v  [9:12] R^T e_z  [12:15] rates  [15] prev_thrust
__device__ __forceinline__ void gate_observation(const float4 p, const float4 v, const float4 q, const float4 w,
                                                 const fpv_gate_t& gt, float4* o) {
  float R[9];
  quat_to_matrix(q, R);
  const float dx = gt.cx - p.x, dy = gt.cy - p.y, dz = gt.cz - p.z;
  // R^T x = columns of R dotted with x
  auto col = [&](int c, float x, float y, float z) { return __fmaf_rn(R[c], x, __fmaf_rn(R[3 + c], y, R[6 + c] * z)); };
  o[0] = make_float4(col(0, dx, dy, dz), col(1, dx, dy, dz), col(2, dx, dy, dz), col(0, gt.nx, gt.ny, gt.nz));
  o[1] = make_float4(col(1, gt.nx, gt.ny, gt.nz), col(2, gt.nx, gt.ny, gt.nz), col(0, v.x, v.y, v.z), col(1, v.x, v.y, v.z));
  o[2] = make_float4(col(2, v.x, v.y, v.z), R[6], R[7], R[8]);
  o[3] = make_float4(w.x, w.y, w.z, p.w);
}

__global__ void gate_env_reset_kernel(const __grid_constant__ fpv_gate_env_params_t k, const float4* state, long long n,
                                      long long stride, const unsigned char* mask, float2* prev, int* progress) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && !mask[i]) return;
  const float4 p = state[i];
  float d, r;
  gate_metrics(k.gates[0], p.x, p.y, p.z, d, r);
  prev[i] = make_float2(d, r);
  progress[i] = 0;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) gate_env_step_kernel(const __grid_constant__ fpv_gate_env_params_t k,
                                                                const float4* state, long long n, long long stride,
                                                                const unsigned char* agent_done, float2* prev,
                                                                int* progress, float* agent_reward, float* env_reward,
                                                                unsigned char* env_done, float4* obs, fpv_stats_t* stats) {
  // The gate table is indexed PER AGENT (every agent is at its own gate): from the kernel-parameter constant bank a warp
  // with 8 different gates would serialise every field load 8 times, so the table is staged in shared memory first.
  __shared__ fpv_gate_t gates[FPV_MAX_GATES];
  {
    const float* src = reinterpret_cast<const float*>(k.gates);
    float* dst = reinterpret_cast<float*>(gates);
    for (int j = threadIdx.x; j < k.n_gates * (int)(sizeof(fpv_gate_t) / sizeof(float)); j += THREADS) dst[j] = src[j];
    __syncthreads();
  }
  const long long i = (long long)blockIdx.x * THREADS + threadIdx.x;
  const int A = k.agents_per_env;
  const bool live = i < n;  // n is a multiple of A (checked by the host), so groups are all-live or all-dead
  float reward = 0.f;
  bool crashed = false, finished = false;
  if (live) {
    const float4 p = state[i], v = state[stride + i], q = state[2 * stride + i], w = state[3 * stride + i];
    const float2 pr = prev[i];
    int prog = progress[i];
    int g = prog & 0xffff, laps = prog >> 16;
    crashed = agent_done[i] != 0;
    float d, r;
    gate_metrics(gates[g], p.x, p.y, p.z, d, r);
    bool passed = false;
    if (!crashed) {
      passed = pr.x < 0.f && d >= 0.f && __fmaf_rn(r, r, -(d * d)) <= gates[g].half_size * gates[g].half_size;
      reward = __fmaf_rn(k.w_progress, pr.y - r, passed ? k.w_gate : 0.f);
    } else {
      reward = -k.w_crash;
    }
    if (passed) {
      g += 1;
      if (g == k.n_gates) { g = 0; laps += 1; }
    }
    if (crashed) { g = 0; laps = 0; }
    finished = k.laps_to_finish > 0 && laps >= k.laps_to_finish;
    if (passed || crashed) gate_metrics(gates[g], p.x, p.y, p.z, d, r);  // re-base on the agent's next gate
    prev[i] = make_float2(d, r);
    progress[i] = (laps << 16) | g;
    if (agent_reward) agent_reward[i] = reward;
    if (obs) gate_observation(p, v, q, w, gates[g], obs + 4 * i);
  }
  // ---- per-env reductions inside the aligned A-lane group: reward sum (xor shuffles), any crashed / finished (vote)
  float team = reward;
  for (int o = A >> 1; o > 0; o >>= 1) team += __shfl_xor_sync(0xffffffffu, team, o);
  const unsigned lane = threadIdx.x & 31;
  const unsigned group_mask = (A == 32 ? 0xffffffffu : ((1u << A) - 1u)) << (lane & ~(unsigned)(A - 1));
  const unsigned flags = __ballot_sync(0xffffffffu, crashed || finished);
  const bool done = (flags & group_mask) != 0;
  const bool head = live && (lane & (unsigned)(A - 1)) == 0;
  if (head) {
    const long long e = i / A;
    env_reward[e] = team;
    env_done[e] = done ? 1 : 0;
  }
  // ---- statistics: block-level sum of env rewards, one atomic pair per CTA
  if (stats) {
    __shared__ float s_sum[THREADS / 32], s_sq[THREADS / 32];
    float a = head ? team : 0.f, b = head ? team * team : 0.f;
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { s_sum[threadIdx.x >> 5] = a; s_sq[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float ta = 0.f, tb = 0.f;
      for (int wi = 0; wi < THREADS / 32; ++wi) { ta += s_sum[wi]; tb += s_sq[wi]; }
      atomicAdd(&stats->reward_sum, (double)ta);
      atomicAdd(&stats->reward_sq_sum, (double)tb);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused env step (fpv_gate_race_step): the env step above as the per-chunk epilogue (`Post`) of the packed ring kernel
// (ring_step_kernel<DroneMode<F2, ANG, false, GatePost, GateIO>>).  A warp's chunk is 64 agents; lane t holds agent
// 64c + t in the low halves of its packed registers and agent 64c + 32 + t in the high halves, so with 32 agents per env
// the warp holds two whole envs and every aligned group of A <= 32 lanes of one slot is one env: the team reward is the
// same xor-shuffle sum over the same lanes as in gate_env_step_kernel, the termination the same ballot.  The agent's state
// is read once (TMA) and written once; the env logic runs on the FINAL rows (after auto-reset) straight from registers.
// Bit-identical to fpv_drone_step (packed hot kernel) followed by fpv_gate_env_step.
// ---------------------------------------------------------------------------------------------------------------
struct GateIO : DroneIO {
  fpv_gate_env_params_t gp;
  float2* prev;
  int* progress;
  float* agent_reward;
  float* env_reward;
  unsigned char* env_done;
  float4* obs;
};

struct GatePost {
  static constexpr bool enabled = true;
  struct Ctx {
    float sum, sq;       // this lane's share of the env-reward statistics
    float2 prev[2];      // race bookkeeping of the thread's two agents, fetched before the substep loop
    int prog[2];
  };
  static __device__ __forceinline__ Ctx begin(const GateIO&) {
    Ctx c;
    c.sum = 0.f; c.sq = 0.f;
    return c;
  }
  template <int L>
  static __device__ __forceinline__ void prefetch(const GateIO& io, const long long (&ei)[L], Ctx& c) {
#pragma unroll
    for (int l = 0; l < L; ++l) { c.prev[l] = io.prev[ei[l]]; c.prog[l] = io.progress[ei[l]]; }
  }
  static __device__ __forceinline__ int bytes(const GateIO& io) { return io.gp.n_gates * (int)sizeof(fpv_gate_t); }
  // the gate table is indexed PER AGENT (every agent is at its own gate): staged in shared memory, see gate_env_step_kernel
  static __device__ __forceinline__ void stage(const GateIO& io, unsigned char* smem, int tid, int nthreads) {
    const float* src = reinterpret_cast<const float*>(io.gp.gates);
    float* dst = reinterpret_cast<float*>(smem);
    for (int j = tid; j < io.gp.n_gates * (int)(sizeof(fpv_gate_t) / sizeof(float)); j += nthreads) dst[j] = src[j];
  }
  template <int L>
  static __device__ __forceinline__ void run(const GateIO& io, const unsigned char* staged,
                                             const float4 (&fin)[FPV_DRONE_PLANES][L], const bool (&crashed_l)[L],
                                             const bool (&live_l)[L], const long long (&ei)[L], Ctx& c) {
    const fpv_gate_t* gates = reinterpret_cast<const fpv_gate_t*>(staged);
    const fpv_gate_env_params_t& gp = io.gp;
    const int A = gp.agents_per_env;
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const long long i = ei[l];
      const bool live = live_l[l];
      float reward = 0.f;
      bool crashed = false, finished = false;
      if (live) {
        const float4 p = fin[0][l], v = fin[1][l], q = fin[2][l], w = fin[3][l];
        const float2 pr = c.prev[l];
        int prog = c.prog[l];
        int g = prog & 0xffff, laps = prog >> 16;
        crashed = crashed_l[l];
        float d, r;
        gate_metrics(gates[g], p.x, p.y, p.z, d, r);
        bool passed = false;
        if (!crashed) {
          passed = pr.x < 0.f && d >= 0.f && __fmaf_rn(r, r, -(d * d)) <= gates[g].half_size * gates[g].half_size;
          reward = __fmaf_rn(gp.w_progress, pr.y - r, passed ? gp.w_gate : 0.f);
        } else {
          reward = -gp.w_crash;
        }
        if (passed) {
          g += 1;
          if (g == gp.n_gates) { g = 0; laps += 1; }
        }
        if (crashed) { g = 0; laps = 0; }
        finished = gp.laps_to_finish > 0 && laps >= gp.laps_to_finish;
        if (passed || crashed) gate_metrics(gates[g], p.x, p.y, p.z, d, r);  // re-base on the agent's next gate
        io.prev[i] = make_float2(d, r);
        io.progress[i] = (laps << 16) | g;
        if (io.agent_reward) io.agent_reward[i] = reward;
        if (io.obs) gate_observation(p, v, q, w, gates[g], io.obs + 4 * i);
      }
      // ---- per-env reductions inside the aligned A-lane group of this slot
      float team = reward;
      for (int o = A >> 1; o > 0; o >>= 1) team += __shfl_xor_sync(0xffffffffu, team, o);
      const unsigned group_mask = (A == 32 ? 0xffffffffu : ((1u << A) - 1u)) << (lane & ~(unsigned)(A - 1));
      const unsigned flags = __ballot_sync(0xffffffffu, crashed || finished);
      const bool done = (flags & group_mask) != 0;
      const bool head = live && (lane & (unsigned)(A - 1)) == 0;
      if (head) {
        const long long e = i / A;
        io.env_reward[e] = team;
        io.env_done[e] = done ? 1 : 0;
        c.sum += team;
        c.sq += team * team;
      }
    }
  }
  static __device__ __forceinline__ void finish(const GateIO& io, Ctx& c) {
    if (!io.stats) return;
    float a = c.sum, b = c.sq;
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0 && (a != 0.f || b != 0.f)) {
      atomicAdd(&io.stats->reward_sum, (double)a);
      atomicAdd(&io.stats->reward_sq_sum, (double)b);
    }
  }
};

}  // namespace fpv
