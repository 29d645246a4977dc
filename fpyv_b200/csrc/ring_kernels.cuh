// The persistent TMA-ring kernel shared by every dynamics mode (mode A hot path and general path, the gate-race env step,
// mode B `Racer`, mode C acro).
//
// Design (B200): gridDim.x = SMs x resident CTAs persistent CTAs of THREADS/32 warps.  Every WARP owns a private 2-slot ring
// in shared memory.  One elected lane is the producer for the warp's own 64-env chunks: per chunk it arms the slot's
// mbarrier with the byte count and issues Mode::ROWS 1-KiB TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx, SASS
// UBLKCP) -- one per float4 plane of the mode's state plus the actions.  The 32 lanes are the consumers: they drain the slot
// with LDS.128 into registers, the slot is refilled at once, and the arithmetic (Mode::tile) runs from registers while the
// next chunk lands.  __syncwarp() is the only synchronisation; the single CTA barrier stages the mode's shared tables
// (motor-curve LUT, gate table, obstacle table).
//   smem: [ staged tables | WARPS x 2 x ROWS x 64 float4 | WARPS x 2 mbarriers ]
// Chunks are PULLED (one atomicAdd per chunk after the two static first ones): the SM's warp arbiter is not fair, a static
// split leaves the slowest warp of a scheduler finishing alone.
//
// A Mode provides:
//   using V (float or F2), K (launch constants, passed by value), IO (pointers; must carry n, work, chunk_epoch, epoch, err,
//   trace), Ctx (per-thread accumulators);  static constexpr int ROWS;
//   static bool chained(const K&);                         FPV_F_CHAINED honoured by this launch
//   static int row_bytes(int r);                           bytes per env of row r: 16 (a float4 plane), 8 or 6 (raw sticks)
//   static const void* row_ptr(const IO&, int r, long long first);   global address of env `first` in row r
//   static void stage(const K&, const IO&, unsigned char* smem, int tid, int nthreads);   fill the staged tables
//   static Ctx begin(const K&, const IO&);                 once per thread (also the grid-wide "env_steps" bookkeeping)
//   static void tile(const K&, const IO&, const unsigned char* staged, const float4 (&rows)[ROWS][L], const long long (&ei)[L],
//                    long long base, Ctx&, PreStore);      arithmetic + stores of this thread's L envs of the chunk
//   static void finish(const K&, const IO&, Ctx&);         once per thread at kernel end (statistics flush)
#pragma once
#include <cuda_runtime.h>

#include "vec.cuh"

namespace fpv {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};

#ifndef FPV_PUBLISH_BATCH
#define FPV_PUBLISH_BATCH 4
#endif
// a chained launch that waits longer than this for a chunk's epoch raises io.err[0] and falls back to a grid-wide wait
#ifndef FPV_CHAIN_TIMEOUT_NS
#define FPV_CHAIN_TIMEOUT_NS 2000000000ull
#endif

template <class Mode, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) ring_step_kernel(const __grid_constant__ typename Mode::K k,
                                                                  const typename Mode::IO io, const int stage_bytes) {
  using V = typename Mode::V;
  constexpr int L = Lane<V>::N;
  constexpr int CHUNK = 32 * L;               // envs per warp-chunk
  constexpr int ROWS = Mode::ROWS;            // state planes + actions
  constexpr int WARPS = THREADS / 32;
  constexpr int STAGES = 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* ring_all = reinterpret_cast<float4*>(smem_raw + stage_bytes);  // [WARPS][STAGES][ROWS][CHUNK]
  unsigned long long* full_all = reinterpret_cast<unsigned long long*>(ring_all + WARPS * STAGES * ROWS * CHUNK);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // warp-uniform id
  float4* ring = ring_all + (size_t)warp * STAGES * ROWS * CHUNK;
  unsigned long long* full = full_all + warp * STAGES;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();  // this warp's mbarriers are initialised: its ring can be primed before the CTA-wide staging
  // Programmatic dependent launch: let the NEXT launch on the stream become resident as our CTAs retire (its prologue then
  // overlaps our tail), and wait for the PREVIOUS launch -- which may have written this very state -- before the first
  // byte of state is touched.  Without the launch attribute both instructions are no-ops.
  // A CHAINED launch skips the grid-wide wait: the caller vouches for the side inputs, and the state is ordered chunk by
  // chunk through io.chunk_epoch (acquire before a chunk's TMA loads, release after its stores), so this grid's first
  // chunks run on the SMs the previous grid has already left while that grid's last chunks are still computing.
  const bool chained = Mode::chained(k);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (!chained) asm volatile("griddepcontrol.wait;" ::: "memory");

  const long long n_chunks = (io.n + CHUNK - 1) / CHUNK;
  const long long my_warp = (long long)blockIdx.x * WARPS + warp;
  const long long total_warps = (long long)gridDim.x * WARPS;

  bool chain_broken = false;   // a chained wait of this warp timed out: the rest of the launch runs on the grid-wide wait
  // producer: arm the slot's mbarrier with the byte count, then one bulk copy per row (all operands warp-uniform)
  auto issue = [&](long long chunk, int slot) {
    const long long first = chunk * CHUNK;
    const long long rem = io.n - first;
    const unsigned count = (unsigned)(rem < (long long)CHUNK ? rem : (long long)CHUNK);
    float4* dst = ring + (size_t)slot * ROWS * CHUNK;
    if (chained && !chain_broken) {  // the previous step of THIS chunk must have been stored (possibly by a grid still running)
      const unsigned* f = io.chunk_epoch + chunk;
      unsigned v;
      unsigned long long t0 = 0;
      for (unsigned spins = 0;; ++spins) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v == io.epoch) break;
        __nanosleep(64);
        // A chunk that does not reach this epoch within FPV_CHAIN_TIMEOUT_NS of wall-clock time means the caller broke the
        // FPV_F_CHAINED contract (e.g. replayed a captured launch with a stale epoch) -- or the device is shared / being
        // debugged and the producer is merely slow.  Neither may poison the context: raise the error word, fall back to the
        // grid-wide wait every plain launch uses, and go on (the host reports the word; results of a broken contract are
        // the caller's).
        if ((spins & 1023u) == 1023u) {
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          else if (now - t0 > FPV_CHAIN_TIMEOUT_NS) {
            if (io.err) atomicAdd(io.err, 1u);
            asm volatile("griddepcontrol.wait;" ::: "memory");
            chain_broken = true;
            break;
          }
        }
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");  // generic-proxy stores -> async-proxy (TMA) loads
    }
    unsigned total = 0;   // every copy is a multiple of 16 bytes (a narrow row of a ragged tail chunk is rounded up; the
#pragma unroll          // caller pads such a buffer to 16 bytes)
    for (int r = 0; r < ROWS; ++r) total += (count * (unsigned)Mode::row_bytes(r) + 15u) & ~15u;
    mbar_expect_tx(&full[slot], total);
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
      tma_load_1d(dst + r * CHUNK, Mode::row_ptr(io, r, first), (count * (unsigned)Mode::row_bytes(r) + 15u) & ~15u, &full[slot]);
  };

  const bool dynamic = io.work != nullptr;
  unsigned long long t_start = 0;
  if (io.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
  const bool leader = elect_one();
  const int leader_lane = __ffs(__ballot_sync(0xffffffffu, leader)) - 1;
  // a pull is split in two so that the atomic's round trip (~1k cycles) hides behind other warps' arithmetic
  auto pull = [&]() -> long long {
    unsigned v = 0;
    if (leader) v = atomicAdd(io.work, 1u);
    v = __shfl_sync(0xffffffffu, v, leader_lane);
    return 2 * total_warps + (long long)v;   // chunks [0, 2*total_warps) are the static first two of every warp
  };

  // Ring protocol: slot (it & 1) holds the chunk computed at iteration it.  At the top of iteration it the OTHER slot --
  // drained at it-1 -- is refilled with the next chunk, which then lands while chunk `it` is being computed.
  long long cur = my_warp;                 // static first chunk
  if (cur < n_chunks && leader) issue(cur, 0);
  long long static_next = my_warp + total_warps;
  // while the first chunk is in flight: stage the mode's tables (the only CTA-wide barrier of the kernel)
  Mode::stage(k, io, smem_raw, (int)threadIdx.x, THREADS);
  __syncthreads();
  typename Mode::Ctx ctx = Mode::begin(k, io);
  unsigned pend[FPV_PUBLISH_BATCH + 1];   // chunks whose stores are issued but whose epochs are not published yet
  int n_pend = 0;
  for (int it = 0; cur < n_chunks; ++it) {
    const int slot = it & 1;
    // next chunk: the second one is static too (no start-up burst of atomics), later ones are pulled
    long long nxt = static_next;
    if (dynamic && it > 0) nxt = pull();
    static_next += total_warps;
    if (nxt < n_chunks && leader) issue(nxt, slot ^ 1);
    mbar_wait(&full[slot], (unsigned)(it >> 1) & 1u);
    const float4* src = ring + (size_t)slot * ROWS * CHUNK;
    const long long base = cur * CHUNK + lane;
    long long ei[L];
    float4 rows[ROWS][L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      ei[l] = min(base + (long long)l * 32, io.n - 1);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (Mode::row_bytes(r) == 16) {
          rows[r][l] = src[r * CHUNK + l * 32 + lane];
        } else if (Mode::row_bytes(r) == 8) {    // 8-byte elements (uint16 x 4): bit patterns in .x .y
          const uint2 u = reinterpret_cast<const uint2*>(src + r * CHUNK)[l * 32 + lane];
          rows[r][l] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), 0.f, 0.f);
        } else {                                 // 6-byte elements (3 half-words): one per component
          const unsigned short* h = reinterpret_cast<const unsigned short*>(src + r * CHUNK) + 3 * (l * 32 + lane);
          rows[r][l] = make_float4(__uint_as_float((unsigned)h[0]), __uint_as_float((unsigned)h[1]), __uint_as_float((unsigned)h[2]), 0.f);
        }
      }
    }
    __syncwarp();  // all lanes have drained this slot: it is refilled at the top of the next iteration
    // Publishing a chunk's epoch needs its stores to be performed first: a release is MEMBAR.GPU + ERRBAR, which drains the
    // warp's memory pipeline.  Two things keep that off the critical path: the flags go out in the MIDDLE of a later chunk
    // (after that chunk's arithmetic, when the stores are long done), and they go out in batches of FPV_PUBLISH_BATCH chunks
    // under ONE fence.  Order: all lanes' stores of chunk i -> the __syncwarp() of a later iteration -> the leader's fence ->
    // the flag stores.
    const bool flush_now = n_pend >= FPV_PUBLISH_BATCH;   // warp-uniform
    auto publish_pending = [&]() {
      if (flush_now && leader) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        for (int j = 0; j < n_pend; ++j)
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(io.chunk_epoch + pend[j]), "r"(io.epoch + 1u) : "memory");
      }
    };
    Mode::tile(k, io, smem_raw, rows, ei, base, ctx, publish_pending);
    if (flush_now) n_pend = 0;
    if (io.chunk_epoch) pend[n_pend++] = (unsigned)cur;
    cur = nxt;
  }
  if (n_pend > 0) {  // whatever is still unpublished, the warp's last chunk included
    __syncwarp();
    if (leader) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      for (int j = 0; j < n_pend; ++j)
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(io.chunk_epoch + pend[j]), "r"(io.epoch + 1u) : "memory");
    }
  }
  if (dynamic && leader) {   // the last warp to finish puts the two counters back to zero for the next launch
    const unsigned finished = atomicAdd(io.work + 1, 1u);
    if (finished == (unsigned)total_warps - 1u) { io.work[0] = 0u; io.work[1] = 0u; }
  }
  Mode::finish(k, io, ctx);
  if (io.trace && lane == 0) {
    unsigned long long t_end;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* o = io.trace + 3 * ((size_t)blockIdx.x * WARPS + warp);
    o[0] = t_start; o[1] = t_end; o[2] = smid;
  }
}

}  // namespace fpv
