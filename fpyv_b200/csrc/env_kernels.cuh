// Multi-agent gate-race environment kernels (see include/fpv_api.h "Multi-agent gate-race environment").
// One thread per agent; the agents of an env are an aligned group of A = agents_per_env lanes of one warp, so the
// per-env reward sum and termination flag are warp-shuffle / vote reductions -- the only place warp primitives are
// used on the path (north_star).  Reward definition: OURS (parity unpinned, no reference implementation).
#pragma once
#include "../../include/fpv_api.h"
#include "misc_kernels.cuh"

namespace fpv {

__device__ __forceinline__ void gate_metrics(const fpv_gate_t& g, float px, float py, float pz, float& d, float& r) {
  const float dx = px - g.cx, dy = py - g.cy, dz = pz - g.cz;
  d = g.nx * dx + g.ny * dy + g.nz * dz;  // Gate.calculate_distance, components.py:819-822
  r = sqrtf(dx * dx + dy * dy + dz * dz);
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
      passed = pr.x < 0.f && d >= 0.f && (r * r - d * d) <= gates[g].half_size * gates[g].half_size;
      reward = k.w_progress * (pr.y - r) + (passed ? k.w_gate : 0.f);
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
    if (obs) {
      float R[9];
      quat_to_matrix(q, R);
      const fpv_gate_t& gt = gates[g];
      const float dx = gt.cx - p.x, dy = gt.cy - p.y, dz = gt.cz - p.z;
      float4* o = obs + 4 * i;
      // R^T x = columns of R dotted with x
      o[0] = make_float4(R[0] * dx + R[3] * dy + R[6] * dz, R[1] * dx + R[4] * dy + R[7] * dz,
                         R[2] * dx + R[5] * dy + R[8] * dz, R[0] * gt.nx + R[3] * gt.ny + R[6] * gt.nz);
      o[1] = make_float4(R[1] * gt.nx + R[4] * gt.ny + R[7] * gt.nz, R[2] * gt.nx + R[5] * gt.ny + R[8] * gt.nz,
                         R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z);
      o[2] = make_float4(R[2] * v.x + R[5] * v.y + R[8] * v.z, R[6], R[7], R[8]);
      o[3] = make_float4(w.x, w.y, w.z, p.w);
    }
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

}  // namespace fpv
