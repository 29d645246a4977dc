// C ABI of libfpyv_b200.so (see include/fpv_api.h).  Validation + double-precision derivation of the
// launch constants on the host, then one kernel launch on the caller's stream.  No global state.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/fpv_api.h"
#include "drone_kernels.cuh"
#include "misc_kernels.cuh"
#include "racer_kernels.cuh"
#include "env_kernels.cuh"
#include "chase_kernels.cuh"
#include "acro_kernels.cuh"
#include "probe_kernels.cuh"

namespace {

thread_local char g_err[512] = "";
std::mutex g_mu;   // guards every launch-fact cache below (occupancy, opted-in shared memory, SM counts, host pipes)

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return FPV_OK;
  if (e == cudaErrorNoKernelImageForDevice || e == cudaErrorInvalidDeviceFunction)
    return fail(FPV_ENODEV, "%s: no sm_100a kernel image for this device (%s)", what, cudaGetErrorString(e));
  return fail(FPV_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kThreads = 128;
#ifndef FPV_MINB
#define FPV_MINB 4  // resident CTAs per SM the compiler must fit (register budget = 65536 / (kThreads * FPV_MINB))
#endif
#ifndef FPV_GENERAL_MINB
#define FPV_GENERAL_MINB 4  // the general (obstacle) path: 128 registers with ~48 bytes of spills in the (rare) contact code;
                            // measured equal to 3 CTAs/SM without spills in plain order, and it lets chained launches of
                            // independent batches sit side by side (2 + 2 CTA slots per SM) like the hot path's
#endif

using fpv::DroneIO;
using fpv::DroneK;
using fpv::F2;

// cudaFuncSetAttribute, occupancy and the SM count are properties of (kernel, DEVICE): one process may drive several
// devices through this library, so every launch-fact cache below is indexed by the current device.
constexpr int kMaxDevices = 64;

int current_device_slot() {   // -1 = beyond the cached range: such devices recompute their facts on every call
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}

int sm_count_of_current_device() {
  static int cached[kMaxDevices] = {0};
  const int slot = current_device_slot();
  std::lock_guard<std::mutex> lk(g_mu);
  if (slot >= 0 && cached[slot]) return cached[slot];
  int dev = 0, v = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  v = v > 0 ? v : 148;
  if (slot >= 0) cached[slot] = v;
  return v;
}

// Dynamic shared memory a kernel has been opted into, per device.
struct SmemOptIn {
  size_t bytes[kMaxDevices] = {};
};

template <class K>
void opt_in_smem(K kern, SmemOptIn& c, size_t smem, size_t above = 48 * 1024) {
  const int slot = current_device_slot();
  std::lock_guard<std::mutex> lk(g_mu);
  if (smem <= above || (slot >= 0 && smem <= c.bytes[slot])) return;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (slot >= 0) c.bytes[slot] = smem;
}

// Persistent launch: one wave of CTAs (SMs x resident CTAs per SM) walking the tiles.
template <class V, int ANG, bool GENERAL>
void launch_drone(const DroneK& k, const DroneIO& io, cudaStream_t st) {
  constexpr int L = fpv::Lane<V>::N;
  auto kern = fpv::drone_step_kernel<V, ANG, GENERAL, kThreads, GENERAL ? 3 : FPV_MINB>;
  const size_t smem = (k.flags & FPV_F_THRUST_LUT) ? sizeof(float) * (size_t)k.lut_n : 0;
  static SmemOptIn opted;                       // per template instantiation
  static int occ_cache[kMaxDevices][2] = {};    // [device][lut?]
  if (smem > 48 * 1024) opt_in_smem(kern, opted, 200 * 1024);
  const int slot = current_device_slot();
  int occ;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    occ = slot >= 0 ? occ_cache[slot][smem ? 1 : 0] : 0;
    if (occ == 0) {
      int o = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, smem > 48 * 1024 ? 200 * 1024 : smem);
      occ = o > 0 ? o : 1;
      if (slot >= 0) occ_cache[slot][smem ? 1 : 0] = occ;
    }
  }
  const long long per_block = (long long)kThreads * L;
  const long long tiles = (io.n + per_block - 1) / per_block;
  const long long wave = (long long)sm_count_of_current_device() * occ;
  const unsigned grid = (unsigned)(tiles < wave ? tiles : wave);
  kern<<<grid, kThreads, smem, st>>>(k, io);
}

// The persistent TMA-ring kernel (ring_kernels.cuh) for any mode.  One wave of CTAs: SMs x resident CTAs per SM (capped by
// `cta_cap` when > 0).  Launched with the programmatic-stream-serialisation attribute, so the grid may become resident
// while the previous one on the stream drains (the kernel itself waits for it before touching the state unless chained).
// Returns false if the ring does not fit into shared memory.  `chained_ok` tells the caller whether FPV_F_CHAINED could
// be honoured (full persistent grid only); the caller clears the flag in `k` beforehand otherwise.
template <class Mode, int MINB>
bool launch_ring(typename Mode::K& k, const typename Mode::IO& io, size_t stage_bytes, unsigned cta_cap, bool wants_chain,
                 unsigned* chain_flag_word, unsigned chain_flag_bit, cudaStream_t st) {
  constexpr int L = fpv::Lane<typename Mode::V>::N;
  constexpr int TILE = kThreads * L;
  auto kern = fpv::ring_step_kernel<Mode, kThreads, MINB>;
  stage_bytes = (stage_bytes + 127) / 128 * 128;
  const size_t smem = stage_bytes + (size_t)2 * Mode::ROWS * TILE * sizeof(float4) + (size_t)(kThreads / 32) * 2 * sizeof(unsigned long long);
  if (smem > 220 * 1024) return false;
  static SmemOptIn opted;  // per instantiation: largest dynamic smem opted in so far, per device
  static int occ_of[kMaxDevices] = {};
  static size_t occ_smem[kMaxDevices] = {};
  opt_in_smem(kern, opted, smem, 0);
  const int slot = current_device_slot();
  int occ;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    occ = (slot >= 0 && occ_smem[slot] == smem) ? occ_of[slot] : 0;
    if (occ == 0) {
      int o = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, smem);
      occ = o > 0 ? o : 1;
      if (slot >= 0) { occ_of[slot] = occ; occ_smem[slot] = smem; }
    }
  }
  const long long tiles = (io.n + TILE - 1) / TILE;
  const int occ_used = (cta_cap > 0 && (int)cta_cap < occ) ? (int)cta_cap : occ;
  const long long wave = (long long)sm_count_of_current_device() * occ_used;
  const unsigned grid = (unsigned)(tiles < wave ? tiles : wave);
  // FPV_F_CHAINED is honoured only by full persistent grids (every CTA slot a launch may use, on every SM): launch i+1
  // cannot become fully resident before launch i has left the slots it needs, which bounds the number of launches alive
  // at once (4 / slots-per-launch + 1) -- what the eight pull-counter pairs and the forward-progress argument of the
  // per-chunk waits rely on (the producer of an awaited chunk is always resident).  Anything else keeps plain stream order.
  if (wants_chain && (long long)grid != wave && chain_flag_word) *chain_flag_word &= ~chain_flag_bit;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, k, io, (int)stage_bytes);
  return true;
}

// mode A on the ring: hot path (GENERAL = false) or general path (obstacles / overrides / damped spring)
template <class V, int ANG, bool GENERAL>
bool launch_drone_ring(const DroneK& k, const DroneIO& io, cudaStream_t st) {
  using Mode = fpv::DroneMode<V, ANG, GENERAL>;
  DroneK kk = k;
  size_t stage = (k.flags & FPV_F_THRUST_LUT) ? ((size_t)k.lut_n * sizeof(float) + 127) / 128 * 128 : 0;
  if (GENERAL) stage += ((size_t)k.n_objects * sizeof(fpv_object_t) + 15) / 16 * 16;
  return launch_ring<Mode, GENERAL ? FPV_GENERAL_MINB : FPV_MINB>(kk, io, stage, io.cta_cap, (kk.flags & FPV_F_CHAINED) != 0, &kk.flags,
                                                   FPV_F_CHAINED, st);
}

__global__ void fill_u32_kernel(unsigned* p, long long n, unsigned v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void pack_done_bits_kernel(const unsigned char* done, long long n, unsigned* bits) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned word = __ballot_sync(0xffffffffu, e < n && done[e] != 0);
  if ((threadIdx.x & 31) == 0 && e < n) bits[e >> 5] = word;
}

// Kernels without the per-chunk protocol run in plain stream order and publish all epochs afterwards.
template <class V, int ANG, bool GENERAL>
void launch_drone_plain(const DroneK& k, const DroneIO& io, cudaStream_t st) {
  DroneK kk = k;
  kk.flags &= ~FPV_F_CHAINED;
  launch_drone<V, ANG, GENERAL>(kk, io, st);
  if (io.done_bits && io.done)
    pack_done_bits_kernel<<<(unsigned)((io.n + 255) / 256), 256, 0, st>>>(io.done, io.n, io.done_bits);
  if (io.chunk_epoch) {
    const long long chunks = (io.n + 63) / 64;
    fill_u32_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(io.chunk_epoch, chunks, io.epoch + 1u);
  }
}

// hot path with the raw sticks calibrated inside the step (io.sticks): row 4 of every chunk is 8 or 6 bytes per env
template <int ANG, int FMT>
bool launch_drone_ring_sticks(const DroneK& k, const DroneIO& io, cudaStream_t st) {
  using Mode = fpv::DroneMode<F2, ANG, false, fpv::NoPost, DroneIO, FMT>;
  DroneK kk = k;
  const size_t stage = (k.flags & FPV_F_THRUST_LUT) ? ((size_t)k.lut_n * sizeof(float) + 127) / 128 * 128 : 0;
  return launch_ring<Mode, FPV_MINB>(kk, io, stage, io.cta_cap, (kk.flags & FPV_F_CHAINED) != 0, &kk.flags, FPV_F_CHAINED, st);
}
template <int FMT>
bool launch_drone_sticks(const DroneK& k, const DroneIO& io, int ang, cudaStream_t st) {
  if (ang == 4) return launch_drone_ring_sticks<4, FMT>(k, io, st);
  if (ang == 3) return launch_drone_ring_sticks<3, FMT>(k, io, st);
  if (ang == 2) return launch_drone_ring_sticks<2, FMT>(k, io, st);
  if (ang == 1) return launch_drone_ring_sticks<1, FMT>(k, io, st);
  return launch_drone_ring_sticks<0, FMT>(k, io, st);
}

template <class V, int ANG>
void launch_drone_g(const DroneK& k, const DroneIO& io, bool general, cudaStream_t st) {
  // chunk_epoch is indexed by 64-env chunks = the packed kernels' warp-chunk: the scalar instantiations (32-env
  // chunks) take the ring only when no epochs are kept
  if (fpv::Lane<V>::N == 2 || !io.chunk_epoch) {
    if (general ? launch_drone_ring<V, ANG, true>(k, io, st) : launch_drone_ring<V, ANG, false>(k, io, st)) return;
  }
  if (general) launch_drone_plain<V, ANG, true>(k, io, st);
  else launch_drone_plain<V, ANG, false>(k, io, st);
}
template <class V>
void launch_drone_a(const DroneK& k, const DroneIO& io, int ang, bool general, cudaStream_t st) {
  if (ang == 4) launch_drone_g<V, 4>(k, io, general, st);
  else if (ang == 3) launch_drone_g<V, 3>(k, io, general, st);
  else if (ang == 2) launch_drone_g<V, 2>(k, io, general, st);
  else if (ang == 1) launch_drone_g<V, 1>(k, io, general, st);
  else launch_drone_g<V, 0>(k, io, general, st);
}

}  // namespace

extern "C" {

int fpv_abi_version(void) { return FPV_ABI_VERSION; }

const char* fpv_last_error(void) { return g_err; }

int fpv_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(fpv_drone_params_t);
    case 1: return (int)sizeof(fpv_drone_io_t);
    case 2: return (int)sizeof(fpv_object_t);
    case 3: return (int)sizeof(fpv_stats_t);
    case 4: return (int)sizeof(fpv_stick_calib_t);
    case 5: return (int)sizeof(fpv_racer_params_t);
    case 6: return (int)sizeof(fpv_gate_env_params_t);
    case 7: return (int)sizeof(fpv_camera_params_t);
    case 8: return (int)sizeof(fpv_autopilot_params_t);
    case 9: return (int)sizeof(fpv_acro_params_t);
    default: return -1;
  }
}

int fpv_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) return fail(FPV_ENODEV, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return FPV_OK;
}

int fpv_host_alloc(int64_t bytes, int32_t write_combined, void** out) {
  if (!out || bytes <= 0) return fail(FPV_EINVAL, "fpv_host_alloc: bad arguments");
  const cudaError_t e = cudaHostAlloc(out, (size_t)bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(FPV_ECUDA, "fpv_host_alloc(%lld bytes): %s", (long long)bytes, cudaGetErrorString(e));
  return FPV_OK;
}

int fpv_host_free(void* p) {
  if (!p) return FPV_OK;
  const cudaError_t e = cudaFreeHost(p);
  if (e != cudaSuccess) return fail(FPV_ECUDA, "fpv_host_free: %s", cudaGetErrorString(e));
  return FPV_OK;
}

int fpv_probe_fp32(int32_t packed, int32_t iters, float* sink, int64_t sink_floats, double* flop_out, void* stream) {
  if (iters < 1) return fail(FPV_EINVAL, "fpv_probe_fp32: iters must be >= 1");
  const int blocks = sm_count_of_current_device() * 8;
  if (!sink || sink_floats < (int64_t)blocks * 256) return fail(FPV_EINVAL, "fpv_probe_fp32: sink must hold %d floats", blocks * 256);
  if (packed) fpv::fp32_probe_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(sink, iters, 1.0001f, 0.5f);
  else fpv::fp32_probe_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(sink, iters, 1.0001f, 0.5f);
  if (flop_out) *flop_out = (double)blocks * 256.0 * (double)iters * 16.0 * (packed ? 2.0 : 1.0) * 2.0;
  return check_launch("fpv_probe_fp32");
}

int fpv_drone_reset(void* state, int64_t n, int64_t plane_stride, const float* pos, const float* vel,
                    const float* rpy_deg, const uint8_t* mask, void* stream) {
  if (!state || !pos || !vel || !rpy_deg) return fail(FPV_EINVAL, "fpv_drone_reset: null pointer");
  if (n < 0 || plane_stride < n) return fail(FPV_EINVAL, "fpv_drone_reset: bad n=%lld stride=%lld", (long long)n, (long long)plane_stride);
  if (!aligned16(state)) return fail(FPV_EINVAL, "fpv_drone_reset: state must be 16-byte aligned");
  if (n == 0) return FPV_OK;
  const unsigned grid = (unsigned)((n + 255) / 256);
  fpv::drone_reset_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)state, n, plane_stride, pos, vel, rpy_deg, mask);
  return check_launch("fpv_drone_reset");
}

namespace {
int make_sticks(const fpv_stick_calib_t* c, fpv::StickK& k, const char* who) {
  if (!c) return fail(FPV_EINVAL, "%s: null calibration", who);
  for (int i = 0; i < 6; ++i) {
    const double span = (double)c->max_vals[i] - (double)c->min_vals[i];
    if (span == 0.0) return fail(FPV_EINVAL, "%s: axis %d has max == min", who, i);
    k.min_v[i] = c->min_vals[i];
    k.inv_span2[i] = (float)(2.0 / span);
    k.sign[i] = c->sign_reverse[i];
  }
  for (int s = 0; s < 4; ++s) {
    if (c->stick_idx[s] < 0 || c->stick_idx[s] > 5) return fail(FPV_EINVAL, "%s: stick idx out of range", who);
    const double ctr = c->stick_center[s];
    if (ctr <= -1.0 || ctr >= 1.0) return fail(FPV_EINVAL, "%s: stick centre must be inside (-1,1)", who);
    k.idx[s] = c->stick_idx[s];
    k.center[s] = c->stick_center[s];
    k.inv_lo[s] = (float)(1.0 / (ctr + 1.0));
    k.inv_hi[s] = (float)(1.0 / (1.0 - ctr));
  }
  return FPV_OK;
}
}  // namespace

namespace {
// Validation and the double-precision derivation of the launch constants shared by fpv_drone_step and
// fpv_drone_rollout.  Returns FPV_OK, a negative error, or 1 for an empty batch.
int prepare_drone(const fpv_drone_params_t* p, const fpv_drone_io_t* io, bool need_actions, DroneK& k, DroneIO& d,
                  int& ang, bool& general) {
  if (!p || !io) return fail(FPV_EINVAL, "fpv_drone_step: null params/io");
  if (io->n < 0 || io->plane_stride < io->n)
    return fail(FPV_EINVAL, "fpv_drone_step: bad n=%lld stride=%lld", (long long)io->n, (long long)io->plane_stride);
  if (io->n == 0) return 1;   /* empty batch: nothing to launch */
  if (!io->state || (need_actions && !io->actions && !io->sticks))
    return fail(FPV_EINVAL, "fpv_drone_step: state/actions must not be null");
  if (io->sticks) {
    if (io->stick_format != FPV_STICKS_U16 && io->stick_format != FPV_STICKS_CRSF)
      return fail(FPV_EINVAL, "fpv_drone_step: io.sticks needs stick_format FPV_STICKS_U16 or FPV_STICKS_CRSF");
    if (!aligned16(io->sticks)) return fail(FPV_EINVAL, "fpv_drone_step: io.sticks must be 16-byte aligned");
    if (int rc = make_sticks(io->stick_calib, d.stick, "fpv_drone_step")) return rc;
  }
  if (!aligned16(io->state) || !aligned16(io->actions) || !aligned16(io->wind_env) || !aligned16(io->acc_out) ||
      !aligned16(io->reset_state) || !aligned16(io->override_q))
    return fail(FPV_EINVAL, "fpv_drone_step: float4 planes must be 16-byte aligned");
  if (p->substeps < 1) return fail(FPV_EINVAL, "fpv_drone_step: substeps must be >= 1 (got %d)", p->substeps);
  if (!(p->dt > 0.f) || !(p->mass > 0.f)) return fail(FPV_EINVAL, "fpv_drone_step: dt and mass must be positive");
  if (p->n_objects < 0 || p->n_objects > FPV_MAX_OBJECTS)
    return fail(FPV_EINVAL, "fpv_drone_step: n_objects=%d out of range [0,%d]", p->n_objects, FPV_MAX_OBJECTS);
  if (p->n_objects > 0 && !io->objects) return fail(FPV_EINVAL, "fpv_drone_step: n_objects > 0 but objects is null");
  if (io->override_q && p->substeps != 1)
    return fail(FPV_EINVAL, "fpv_drone_step: the rotation/thrust override is a per-step input; it needs substeps == 1");
  if (io->override_q && !io->override_thrust)
    return fail(FPV_EINVAL, "fpv_drone_step: override_q needs override_thrust (components.py:230-232)");
  if ((p->flags & FPV_F_AUTO_RESET) && !io->reset_state)
    return fail(FPV_EINVAL, "fpv_drone_step: FPV_F_AUTO_RESET needs io.reset_state");
  if ((p->flags & FPV_F_AUTO_RESET) && (p->flags & FPV_F_FREEZE_DONE))
    return fail(FPV_EINVAL, "fpv_drone_step: AUTO_RESET and FREEZE_DONE are mutually exclusive");
  if (io->done_bits && !io->done)
    return fail(FPV_EINVAL, "fpv_drone_step: io.done_bits needs io.done as well (the scalar fallback packs the bytes)");
  if (io->done_bits && (reinterpret_cast<uintptr_t>(io->done_bits) & 3u))
    return fail(FPV_EINVAL, "fpv_drone_step: io.done_bits must be 4-byte aligned");
  if (p->flags & FPV_F_THRUST_LUT) {
    if (!io->lut || io->lut_n < 2) return fail(FPV_EINVAL, "fpv_drone_step: FPV_F_THRUST_LUT needs io.lut with lut_n >= 2");
    if ((size_t)io->lut_n * sizeof(float) > 200 * 1024) return fail(FPV_EINVAL, "fpv_drone_step: lut_n=%d does not fit in shared memory", io->lut_n);
  }

  std::memset(&k, 0, sizeof(k));
  const double rtr = p->rates_transition_rate, ttr = p->thrust_transition_rate;
  k.dt = p->dt;
  k.substeps = p->substeps;
  k.one_minus_rtr = (float)(1.0 - rtr);
  k.max_rates = p->max_rates;
  k.rtr = p->rates_transition_rate;
  k.ttr = p->thrust_transition_rate;
  k.one_minus_ttr = (float)(1.0 - ttr);
  k.kd0 = p->k_drag[0];
  k.kd_a = (float)((double)p->k_drag[1] - (double)p->k_drag[0]);
  k.kd_b = (float)((double)p->k_drag[2] - (double)p->k_drag[0]);
  {  // bound on the drag force per |u|^2 for the per-control-step reach test of the general path; NaN (= "no bound, test
     // every obstacle every substep") when the low-pass weights are not convex weights
    const double kmax = std::fmax(std::fabs((double)p->k_drag[0]), std::fmax(std::fabs((double)p->k_drag[1]), std::fabs((double)p->k_drag[2])));
    const bool convex = rtr >= 0.0 && rtr <= 1.0 && ttr >= 0.0 && ttr <= 1.0;
    k.kd_max = convex ? (float)(kmax * 1.001) : NAN;
  }
  for (int m = 0; m < 4; ++m)
    for (int j = 0; j < 2; ++j) { k.motor_xy[m][j] = p->motor_xy[m][j]; k.neg_motor_xy[m][j] = -p->motor_xy[m][j]; }
  k.motor_radius = p->motor_radius;
  {
    double arm = 0.0;
    for (int m = 0; m < 4; ++m) arm = std::fmax(arm, std::hypot((double)p->motor_xy[m][0], (double)p->motor_xy[m][1]));
    k.arm_reach = (float)(arm + (double)p->motor_radius + 2e-3);   // margin covers the fp32 rounding of the test itself
  }
  k.spring_k = p->spring_k;
  k.spring_c = p->spring_c;
  for (int i = 0; i < 4; ++i) k.poly[i] = p->thrust_poly[i];
  for (int i = 0; i < 3; ++i) k.wind[i] = p->wind[i];
  k.grav_force_z = (float)(-(double)p->gravity * (double)p->mass);
  k.inv_mass = (float)(1.0 / (double)p->mass);
  k.dt_over_mass = (float)((double)p->dt / (double)p->mass);
  k.half_ang_scale = (float)(0.5 * 0.017453292519943295 * (double)p->dt);
  k.inv_half_ang_scale = (float)(1.0 / (0.5 * 0.017453292519943295 * (double)p->dt));
  k.lut_n = (p->flags & FPV_F_THRUST_LUT) ? io->lut_n : 0;
  k.lut_scale = (float)((io->lut_n - 1) * 0.5);
  k.flags = p->flags;
  k.n_objects = p->n_objects;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};   // bounding box of all obstacles
  for (int i = 0; i < p->n_objects; ++i) {
    k.objects[i] = io->objects[i];
    const fpv_object_t& o = k.objects[i];
    if (o.kind != FPV_OBJ_SPHERE && o.kind != FPV_OBJ_CYLINDER)
      return fail(FPV_EINVAL, "fpv_drone_step: object %d has unknown kind %d", i, o.kind);
    const double r = std::fabs((double)o.a), zlo = o.kind == FPV_OBJ_SPHERE ? o.z - r : std::fmin((double)o.z, (double)o.z + o.b),
                 zhi = o.kind == FPV_OBJ_SPHERE ? o.z + r : std::fmax((double)o.z, (double)o.z + o.b);
    lo[0] = std::fmin(lo[0], o.x - r); hi[0] = std::fmax(hi[0], o.x + r);
    lo[1] = std::fmin(lo[1], o.y - r); hi[1] = std::fmax(hi[1], o.y + r);
    lo[2] = std::fmin(lo[2], zlo);     hi[2] = std::fmax(hi[2], zhi);
  }
  if (p->n_objects > 0) {   // the sphere around that box (plus a margin for the float32 rounding of the test itself)
    double r2 = 0.0;
    for (int a = 0; a < 3; ++a) { k.bound[a] = (float)(0.5 * (lo[a] + hi[a])); r2 += 0.25 * (hi[a] - lo[a]) * (hi[a] - lo[a]); }
    k.bound[3] = (float)(std::sqrt(r2) * 1.0001 + 1e-3);
    if (!(k.bound[3] < 1e18f)) k.bound[3] = INFINITY;   // non-finite obstacle data: the broad phase never culls
  }

  d.state = (float4*)io->state;
  d.n = io->n;
  d.stride = io->plane_stride;
  d.actions = (const float4*)io->actions;
  d.sticks = io->sticks;
  d.wind_env = (const float4*)io->wind_env;
  d.lut = io->lut;
  d.done = io->done;
  d.done_bits = (unsigned*)io->done_bits;
  d.acc_out = (float4*)io->acc_out;
  d.reset_state = (const float4*)io->reset_state;
  d.override_q = (const float4*)io->override_q;
  d.override_thrust = io->override_thrust;
  d.stats = io->stats;
  // pulled chunks pay one atomic round trip per chunk: worth it once a chunk carries enough arithmetic to hide it
  d.work = p->substeps < 4 ? nullptr : (unsigned*)io->work;
  d.chunk_epoch = (unsigned*)io->chunk_epoch;
  d.epoch = io->epoch;
  d.cta_cap = io->max_ctas_per_sm;
  d.err = io->work ? (unsigned*)io->work + 16 : nullptr;
  // launches that overlap (FPV_F_CHAINED) must not share the pull counters.  A chained launch is honoured only as a full
  // persistent grid of k CTA slots per SM, so at most 4/k + 1 <= 5 launches are ever resident together: eight counter
  // pairs indexed by the epoch are never shared by two live launches.
  if (d.work && d.chunk_epoch) d.work += 2 * (io->epoch & 7u);
  if ((p->flags & FPV_F_CHAINED) && !io->chunk_epoch)
    return fail(FPV_EINVAL, "fpv_drone_step: FPV_F_CHAINED needs io.chunk_epoch");
  d.trace = (unsigned long long*)io->trace;

  // |rates| <= max_rates is an invariant of action2force (a convex mix of clipped commands), so the
  // per-substep Euler angles are bounded by max_rates*dt in radians; the polynomial degree follows that bound
  // (see vsincos in vec.cuh), full-range sincosf beyond it.
  const double max_angle = std::fabs((double)p->max_rates) * 0.017453292519943295 * (double)p->dt;
  const double half = 0.5 * max_angle;   // the kernel evaluates sin/cos of the HALF angles (quaternion update)
  ang = half <= 0.008 ? 4 : (half <= 0.03 ? 3 : (half <= 0.05 ? 2 : (half <= 0.25 ? 1 : 0)));
  // hot kernel = reference configuration (ground plane, undamped contact spring); everything else is general
  general = p->n_objects > 0 || io->override_q != nullptr || p->spring_c != 0.f || !(p->flags & FPV_F_GROUND);
  return FPV_OK;
}
}  // namespace

int fpv_drone_step(const fpv_drone_params_t* p, const fpv_drone_io_t* io, void* stream) {
  DroneK k;
  DroneIO d;
  int ang = 0;
  bool general = false;
  const int rc = prepare_drone(p, io, true, k, d, ang, general);
  if (rc != FPV_OK) return rc > 0 ? FPV_OK : rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (d.sticks) {
    if (general || (p->flags & FPV_F_SCALAR))
      return fail(FPV_EINVAL, "fpv_drone_step: io.sticks is served by the packed hot-path kernel only (no obstacles, overrides, "
                              "damped spring or FPV_F_SCALAR); convert with fpv_sticks_to_actions instead");
    const bool ok = io->stick_format == FPV_STICKS_U16 ? launch_drone_sticks<FPV_STICKS_U16>(k, d, ang, st)
                                                        : launch_drone_sticks<FPV_STICKS_CRSF>(k, d, ang, st);
    if (!ok) return fail(FPV_EINVAL, "fpv_drone_step: the motor-curve table does not fit next to the ring in shared memory");
    return check_launch("fpv_drone_step");
  }
  if (p->flags & FPV_F_SCALAR) launch_drone_a<float>(k, d, ang, general, st);
  else launch_drone_a<F2>(k, d, ang, general, st);
  return check_launch("fpv_drone_step");
}


namespace {
// One pipe per device: two copy streams and the events that order them against the caller's stream.  `mu` is held
// for the whole enqueue of one *_step_host call, so concurrent callers on the same device (other host threads, other
// caller streams) queue their slices one call after the other instead of re-recording each other's events.
struct HostPipe {
  cudaStream_t in = nullptr, out = nullptr;
  cudaEvent_t ready = nullptr, joined = nullptr;
  cudaEvent_t h2d[16] = {}, stepped[16] = {};
  bool ok = false;
  std::mutex mu;
};
HostPipe& host_pipe_of_current_device() {
  static HostPipe pipes[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  HostPipe& p = pipes[dev];
  std::lock_guard<std::mutex> lk(g_mu);
  if (!p.ok) {
    cudaStreamCreateWithFlags(&p.in, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&p.out, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&p.ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&p.joined, cudaEventDisableTiming);
    for (int i = 0; i < 16; ++i) {
      cudaEventCreateWithFlags(&p.h2d[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&p.stepped[i], cudaEventDisableTiming);
    }
    p.ok = cudaGetLastError() == cudaSuccess;
  }
  return p;
}
// Env ranges of a sliced host step: (slices - 1) equal ranges followed by one SHORT tail range.  Everything after the
// last byte of the H2D copy is exposed latency (the last slice's launch and its D2H copy), so the last slice is the
// smallest one a launch still fills (65,536 envs or n / 16); ranges are multiples of 64 envs (the kernels' chunk) and
// at least 65,536 envs (below that a slice is launch-latency bound).  Returns the number of ranges; bound[i]..bound[i+1].
int host_slice_bounds(long long n, int slices, long long* bound) {
  constexpr long long kMin = 65536;
  if (slices <= 0) slices = 4;
  if (slices > 16) slices = 16;
  long long body = n;
  if (slices >= 2 && n >= 4 * kMin) {
    long long tail = n / 16 > kMin ? n / 16 : kMin;
    body = (n - tail) / 64 * 64;   // the tail range starts on a chunk boundary too
    --slices;
  }
  long long count = body / kMin < slices ? body / kMin : slices;
  if (count < 1) count = 1;
  const long long per = ((body + count - 1) / count + 63) / 64 * 64;
  int c = 0;
  bound[0] = 0;
  for (long long a = 0; a < body; a += per) bound[++c] = a + per < body ? a + per : body;
  if (body < n) bound[++c] = n;
  return c;
}
}  // namespace

int fpv_drone_step_host(const fpv_drone_params_t* p, const fpv_drone_io_t* io, const float* actions_host,
                        uint8_t* done_host, int32_t slices, void* stream) {
  if (!p || !io) return fail(FPV_EINVAL, "fpv_drone_step_host: null params/io");
  if (!actions_host) return fail(FPV_EINVAL, "fpv_drone_step_host: null host buffer");
  if (!done_host && !io->done_bits)
    return fail(FPV_EINVAL, "fpv_drone_step_host: done_host may be NULL only if io.done_bits points at (pinned host) memory the "
                            "step kernels write the flags to themselves");
  if (!io->actions || !io->done) return fail(FPV_EINVAL, "fpv_drone_step_host: io.actions / io.done must be device staging buffers");
  if (io->n < 0 || io->plane_stride < io->n) return fail(FPV_EINVAL, "fpv_drone_step_host: bad n/stride");
  if (io->n == 0) return FPV_OK;
  const long long n = io->n;
  long long bound[18];
  const int n_slices = host_slice_bounds(n, slices, bound);
  HostPipe& hp = host_pipe_of_current_device();
  if (!hp.ok) return fail(FPV_ECUDA, "fpv_drone_step_host: could not create the copy streams");
  std::lock_guard<std::mutex> pipe_lock(hp.mu);
  cudaStream_t st = (cudaStream_t)stream;
  cudaEventRecord(hp.ready, st);            // everything queued so far (the previous step reads the staging buffer) ...
  cudaStreamWaitEvent(hp.in, hp.ready, 0);  // ... precedes the first byte of the new actions
  cudaStreamWaitEvent(hp.out, hp.ready, 0);
  fpv_drone_params_t pp = *p;
  pp.flags &= ~FPV_F_CHAINED;
  for (int c = 0; c < n_slices; ++c) {
    const long long a = bound[c], b = bound[c + 1];
    cudaMemcpyAsync((char*)io->actions + 16 * a, (const char*)actions_host + 16 * a, (size_t)(16 * (b - a)), cudaMemcpyHostToDevice, hp.in);
    cudaEventRecord(hp.h2d[c], hp.in);
    cudaStreamWaitEvent(st, hp.h2d[c], 0);
    fpv_drone_io_t s = *io;
    s.state = (char*)io->state + 16 * a;
    s.n = b - a;
    s.actions = (const char*)io->actions + 16 * a;
    s.done = io->done + a;
    if (io->done_bits) s.done_bits = (char*)io->done_bits + a / 8;
    if (io->wind_env) s.wind_env = (const char*)io->wind_env + 16 * a;
    if (io->acc_out) s.acc_out = (char*)io->acc_out + 16 * a;
    if (io->reset_state) s.reset_state = (const char*)io->reset_state + 16 * a;
    if (io->override_q) s.override_q = (const char*)io->override_q + 16 * a;
    if (io->override_thrust) s.override_thrust = io->override_thrust + a;
    s.chunk_epoch = nullptr;
    s.trace = nullptr;
    if (int rc = fpv_drone_step(&pp, &s, stream)) {   // the copies already queued still join the caller's stream
      cudaEventRecord(hp.joined, hp.in);
      cudaStreamWaitEvent(hp.out, hp.joined, 0);
      cudaEventRecord(hp.joined, hp.out);
      cudaStreamWaitEvent(st, hp.joined, 0);
      return rc;
    }
    if (done_host) {
      cudaEventRecord(hp.stepped[c], st);
      cudaStreamWaitEvent(hp.out, hp.stepped[c], 0);
      if (io->done_bits)   // bitmask form: slices start on 64-env boundaries, so every slice owns whole 32-bit words
        cudaMemcpyAsync(done_host + a / 8, (const char*)io->done_bits + a / 8, (size_t)((b - a + 31) / 32 * 4), cudaMemcpyDeviceToHost, hp.out);
      else
        cudaMemcpyAsync(done_host + a, io->done + a, (size_t)(b - a), cudaMemcpyDeviceToHost, hp.out);
    }
  }
  if (done_host) {   // (without done_host the flags are written by the step kernels themselves: `stream` alone orders them)
    cudaEventRecord(hp.joined, hp.out);
    cudaStreamWaitEvent(st, hp.joined, 0);
  }
  return check_launch("fpv_drone_step_host");
}

int fpv_drone_rollout(const fpv_drone_params_t* p, const fpv_drone_io_t* io, const void* actions_seq,
                      int64_t action_stride, int32_t n_steps, uint8_t* done_seq, int64_t done_stride, void* stream) {
  DroneK k;
  DroneIO d;
  int ang = 0;
  bool general = false;
  const int rc = prepare_drone(p, io, false, k, d, ang, general);
  if (rc != FPV_OK) return rc > 0 ? FPV_OK : rc;
  if (n_steps < 0) return fail(FPV_EINVAL, "fpv_drone_rollout: n_steps must be >= 0");
  if (n_steps == 0) return FPV_OK;
  if (!actions_seq || !aligned16(actions_seq) || action_stride < io->n)
    return fail(FPV_EINVAL, "fpv_drone_rollout: actions_seq must be a 16-byte aligned float4[T][action_stride >= n]");
  if (done_seq && done_stride < io->n) return fail(FPV_EINVAL, "fpv_drone_rollout: done_stride must be >= n");
  if (general || io->wind_env || io->sticks || (p->flags & (FPV_F_SCALAR | FPV_F_FREEZE_DONE)) || io->chunk_epoch)
    return fail(FPV_EINVAL, "fpv_drone_rollout: only the hot-path configuration is supported (no obstacles, overrides, "
                            "per-env wind, FPV_F_SCALAR, FPV_F_FREEZE_DONE or chunk_epoch); step instead");
  if (!io->work) return fail(FPV_EINVAL, "fpv_drone_rollout: io.work is required");
  d.work = (unsigned*)io->work;
  d.chunk_epoch = nullptr;
  const size_t smem = (k.flags & FPV_F_THRUST_LUT) ? sizeof(float) * (size_t)k.lut_n : 0;
  auto launch = [&](auto kern) {
    static SmemOptIn opted;
    opt_in_smem(kern, opted, smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem);
    if (occ < 1) occ = 1;
    const long long chunks = (io->n + 63) / 64;
    const long long warps_needed = chunks;                       // one warp-chunk at a time per warp
    const long long wave = (long long)sm_count_of_current_device() * occ;
    const long long ctas = (warps_needed + kThreads / 32 - 1) / (kThreads / 32);
    const unsigned grid = (unsigned)(ctas < wave ? ctas : wave);
    kern<<<grid, kThreads, smem, (cudaStream_t)stream>>>(k, d, (const float4*)actions_seq, (long long)action_stride, (int)n_steps,
                                                        done_seq, (long long)done_stride);
  };
  if (ang == 4) launch(fpv::drone_rollout_kernel<F2, 4, kThreads, FPV_MINB>);
  else if (ang == 3) launch(fpv::drone_rollout_kernel<F2, 3, kThreads, FPV_MINB>);
  else if (ang == 2) launch(fpv::drone_rollout_kernel<F2, 2, kThreads, FPV_MINB>);
  else if (ang == 1) launch(fpv::drone_rollout_kernel<F2, 1, kThreads, FPV_MINB>);
  else launch(fpv::drone_rollout_kernel<F2, 0, kThreads, FPV_MINB>);
  return check_launch("fpv_drone_rollout");
}

int fpv_drone_get_rotation(const void* state, int64_t n, int64_t plane_stride, float* R, void* stream) {
  if (!state || !R) return fail(FPV_EINVAL, "fpv_drone_get_rotation: null pointer");
  if (n < 0 || plane_stride < n) return fail(FPV_EINVAL, "fpv_drone_get_rotation: bad n/stride");
  if (!aligned16(state)) return fail(FPV_EINVAL, "fpv_drone_get_rotation: state must be 16-byte aligned");
  if (n == 0) return FPV_OK;
  fpv::drone_get_rotation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)state, n, plane_stride, R);
  return check_launch("fpv_drone_get_rotation");
}

int fpv_drone_set_rotation(void* state, int64_t n, int64_t plane_stride, const float* R, const uint8_t* mask,
                           void* stream) {
  if (!state || !R) return fail(FPV_EINVAL, "fpv_drone_set_rotation: null pointer");
  if (n < 0 || plane_stride < n) return fail(FPV_EINVAL, "fpv_drone_set_rotation: bad n/stride");
  if (!aligned16(state)) return fail(FPV_EINVAL, "fpv_drone_set_rotation: state must be 16-byte aligned");
  if (n == 0) return FPV_OK;
  fpv::drone_set_rotation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((float4*)state, n, plane_stride, R, mask);
  return check_launch("fpv_drone_set_rotation");
}

int fpv_matrix_to_quat(const float* R, int64_t n, void* q, void* stream) {
  if (!R || !q) return fail(FPV_EINVAL, "fpv_matrix_to_quat: null pointer");
  if (n < 0 || !aligned16(q)) return fail(FPV_EINVAL, "fpv_matrix_to_quat: bad n or misaligned q");
  if (n == 0) return FPV_OK;
  fpv::matrix_to_quat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(R, n, (float4*)q);
  return check_launch("fpv_matrix_to_quat");
}

int fpv_drone_observe(const void* state, int64_t n, int64_t plane_stride, const void* acc, float* Rt, float* gyro,
                      float* accel, void* stream) {
  if (!state) return fail(FPV_EINVAL, "fpv_drone_observe: null state");
  if (n < 0 || plane_stride < n) return fail(FPV_EINVAL, "fpv_drone_observe: bad n/stride");
  if (accel && !acc) return fail(FPV_EINVAL, "fpv_drone_observe: accel output needs the acc plane");
  if (!aligned16(state) || !aligned16(acc)) return fail(FPV_EINVAL, "fpv_drone_observe: planes must be 16-byte aligned");
  if (n == 0) return FPV_OK;
  const unsigned grid = (unsigned)((n + 255) / 256);
  fpv::drone_observe_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)state, n, plane_stride,
                                                                     (const float4*)acc, Rt, gyro, accel);
  return check_launch("fpv_drone_observe");
}


int fpv_sticks_to_actions(const fpv_stick_calib_t* c, const int32_t* raw, int64_t n, void* actions, float* calibrated,
                          void* stream) {
  if (!c || !raw || !actions) return fail(FPV_EINVAL, "fpv_sticks_to_actions: null pointer");
  if (n < 0) return fail(FPV_EINVAL, "fpv_sticks_to_actions: bad n");
  if (!aligned16(actions)) return fail(FPV_EINVAL, "fpv_sticks_to_actions: actions must be 16-byte aligned");
  fpv::StickK k;
  if (int rc = make_sticks(c, k, "fpv_sticks_to_actions")) return rc;
  if (n == 0) return FPV_OK;
  const unsigned grid = (unsigned)((n + 255) / 256);
  fpv::sticks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(k, raw, n, (float4*)actions, calibrated);
  return check_launch("fpv_sticks_to_actions");
}

int fpv_drone_step_host_sticks(const fpv_drone_params_t* p, const fpv_drone_io_t* io, const fpv_stick_calib_t* calib,
                               const void* sticks_host, void* sticks_dev, int32_t format, uint8_t* done_host, int32_t slices,
                               void* stream) {
  if (format != FPV_STICKS_U16 && format != FPV_STICKS_CRSF)
    return fail(FPV_EINVAL, "fpv_drone_step_host_sticks: format must be FPV_STICKS_U16 or FPV_STICKS_CRSF");
  const long long esz = format == FPV_STICKS_U16 ? 8 : 6;
  if (!p || !io) return fail(FPV_EINVAL, "fpv_drone_step_host_sticks: null params/io");
  if (!sticks_host || !sticks_dev || !done_host) return fail(FPV_EINVAL, "fpv_drone_step_host_sticks: null buffer");
  if (!io->actions || !io->done) return fail(FPV_EINVAL, "fpv_drone_step_host_sticks: io.actions / io.done must be device buffers");
  if (reinterpret_cast<uintptr_t>(sticks_dev) & 7u) return fail(FPV_EINVAL, "fpv_drone_step_host_sticks: sticks_dev must be 8-byte aligned");
  if (io->n < 0 || io->plane_stride < io->n) return fail(FPV_EINVAL, "fpv_drone_step_host_sticks: bad n/stride");
  fpv::StickK k;
  if (int rc = make_sticks(calib, k, "fpv_drone_step_host_sticks")) return rc;
  if (io->n == 0) return FPV_OK;
  const long long n = io->n;
  long long bound[18];
  const int n_slices = host_slice_bounds(n, slices, bound);
  HostPipe& hp = host_pipe_of_current_device();
  if (!hp.ok) return fail(FPV_ECUDA, "fpv_drone_step_host_sticks: could not create the copy streams");
  std::lock_guard<std::mutex> pipe_lock(hp.mu);
  cudaStream_t st = (cudaStream_t)stream;
  cudaEventRecord(hp.ready, st);
  cudaStreamWaitEvent(hp.in, hp.ready, 0);
  cudaStreamWaitEvent(hp.out, hp.ready, 0);
  fpv_drone_params_t pp = *p;
  pp.flags &= ~FPV_F_CHAINED;
  for (int c = 0; c < n_slices; ++c) {
    const long long a = bound[c], b = bound[c + 1];
    cudaMemcpyAsync((char*)sticks_dev + esz * a, (const char*)sticks_host + esz * a, (size_t)(esz * (b - a)), cudaMemcpyHostToDevice, hp.in);
    cudaEventRecord(hp.h2d[c], hp.in);
    cudaStreamWaitEvent(st, hp.h2d[c], 0);
    if (format == FPV_STICKS_U16)
      fpv::sticks4_u16_kernel<<<(unsigned)((b - a + 255) / 256), 256, 0, st>>>(k, (const ushort4*)sticks_dev + a, b - a,
                                                                               (float4*)io->actions + a);
    else
      fpv::sticks4_crsf_kernel<<<(unsigned)((b - a + 255) / 256), 256, 0, st>>>(k, (const unsigned short*)sticks_dev + 3 * a, b - a,
                                                                                (float4*)io->actions + a);
    fpv_drone_io_t s = *io;
    s.state = (char*)io->state + 16 * a;
    s.n = b - a;
    s.actions = (const char*)io->actions + 16 * a;
    s.done = io->done + a;
    if (io->done_bits) s.done_bits = (char*)io->done_bits + a / 8;
    if (io->wind_env) s.wind_env = (const char*)io->wind_env + 16 * a;
    if (io->acc_out) s.acc_out = (char*)io->acc_out + 16 * a;
    if (io->reset_state) s.reset_state = (const char*)io->reset_state + 16 * a;
    s.override_q = nullptr;
    s.override_thrust = nullptr;
    s.sticks = nullptr;
    s.chunk_epoch = nullptr;
    s.trace = nullptr;
    if (int rc = fpv_drone_step(&pp, &s, stream)) {
      cudaEventRecord(hp.joined, hp.in);
      cudaStreamWaitEvent(hp.out, hp.joined, 0);
      cudaEventRecord(hp.joined, hp.out);
      cudaStreamWaitEvent(st, hp.joined, 0);
      return rc;
    }
    cudaEventRecord(hp.stepped[c], st);
    cudaStreamWaitEvent(hp.out, hp.stepped[c], 0);
    if (io->done_bits)   // bitmask form: slices start on 64-env boundaries, so every slice owns whole 32-bit words
      cudaMemcpyAsync(done_host + a / 8, (const char*)io->done_bits + a / 8, (size_t)((b - a + 31) / 32 * 4), cudaMemcpyDeviceToHost, hp.out);
    else
      cudaMemcpyAsync(done_host + a, io->done + a, (size_t)(b - a), cudaMemcpyDeviceToHost, hp.out);
  }
  cudaEventRecord(hp.joined, hp.out);
  cudaStreamWaitEvent(st, hp.joined, 0);
  return check_launch("fpv_drone_step_host_sticks");
}

int fpv_racer_reset(void* state, int64_t n, int64_t plane_stride, const uint8_t* mask, void* stream) {
  if (!state) return fail(FPV_EINVAL, "fpv_racer_reset: null state");
  if (n < 0 || plane_stride < n) return fail(FPV_EINVAL, "fpv_racer_reset: bad n/stride");
  if (!aligned16(state)) return fail(FPV_EINVAL, "fpv_racer_reset: state must be 16-byte aligned");
  if (n == 0) return FPV_OK;
  const unsigned grid = (unsigned)((n + 255) / 256);
  fpv::racer_reset_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)state, n, plane_stride, mask);
  return check_launch("fpv_racer_reset");
}

int fpv_racer_observe(const void* state, int64_t n, int64_t plane_stride, float* R, float* omega, void* stream) {
  if (!state) return fail(FPV_EINVAL, "fpv_racer_observe: null state");
  if (n < 0 || plane_stride < n || !aligned16(state)) return fail(FPV_EINVAL, "fpv_racer_observe: bad n/stride/alignment");
  if (n == 0) return FPV_OK;
  fpv::racer_observe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)state, n, plane_stride, R, omega);
  return check_launch("fpv_racer_observe");
}

int fpv_racer_step(const fpv_racer_params_t* p, void* state, int64_t n, int64_t plane_stride, const void* actions,
                   void* torque_out, void* work, void* stream) {
  if (!p || !state || !actions) return fail(FPV_EINVAL, "fpv_racer_step: null pointer");
  if (n < 0 || plane_stride < n) return fail(FPV_EINVAL, "fpv_racer_step: bad n/stride");
  if (!aligned16(state) || !aligned16(actions) || !aligned16(torque_out))
    return fail(FPV_EINVAL, "fpv_racer_step: planes must be 16-byte aligned");
  if (p->substeps < 1 || !(p->dt > 0.f) || !(p->mass > 0.f)) return fail(FPV_EINVAL, "fpv_racer_step: bad dt/mass/substeps");
  fpv::RacerK k;
  k.dt = p->dt;
  k.inv_dt = (float)(1.0 / (double)p->dt);
  k.substeps = p->substeps;
  k.dt_over_m = (float)((double)p->dt / (double)p->mass);
  for (int i = 0; i < 3; ++i) {
    if (!(p->inertia[i] > 0.f)) return fail(FPV_EINVAL, "fpv_racer_step: inertia must be positive");
    k.dt_over_I[i] = (float)((double)p->dt / (double)p->inertia[i]);
    for (int j = 0; j < 3; ++j) k.gains[i][j] = p->gains[i][j];
  }
  k.vel_decay = p->vel_decay;
  if (n == 0) return FPV_OK;
  fpv::RacerIO io = {};
  io.state = (float4*)state;
  io.n = n;
  io.stride = plane_stride;
  io.actions = (const float4*)actions;
  io.torque_out = (float4*)torque_out;
  io.work = p->substeps >= 4 ? (unsigned*)work : nullptr;
  if (p->flags & FPV_F_SCALAR) launch_ring<fpv::RacerMode<float>, 4>(k, io, 0, 0, false, nullptr, 0, (cudaStream_t)stream);
  else launch_ring<fpv::RacerMode<F2>, 4>(k, io, 0, 0, false, nullptr, 0, (cudaStream_t)stream);
  return check_launch("fpv_racer_step");
}

namespace {
int check_gate_params(const fpv_gate_env_params_t* p, int64_t n, const char* who) {
  if (!p) return fail(FPV_EINVAL, "%s: null params", who);
  if (p->n_gates < 1 || p->n_gates > FPV_MAX_GATES) return fail(FPV_EINVAL, "%s: n_gates=%d out of [1,%d]", who, p->n_gates, FPV_MAX_GATES);
  const int A = p->agents_per_env;
  if (A < 1 || A > 32 || (A & (A - 1)) != 0) return fail(FPV_EINVAL, "%s: agents_per_env=%d must be a power of two <= 32", who, A);
  if (n < 0 || n % A != 0) return fail(FPV_EINVAL, "%s: n_agents=%lld must be a multiple of agents_per_env=%d", who, (long long)n, A);
  return FPV_OK;
}
}  // namespace

int fpv_gate_env_reset(const fpv_gate_env_params_t* p, const void* state, int64_t n, int64_t plane_stride,
                       const uint8_t* mask, void* prev, int32_t* progress, void* stream) {
  if (int rc = check_gate_params(p, n, "fpv_gate_env_reset")) return rc;
  if (!state || !prev || !progress) return fail(FPV_EINVAL, "fpv_gate_env_reset: null pointer");
  if (plane_stride < n || !aligned16(state)) return fail(FPV_EINVAL, "fpv_gate_env_reset: bad stride/alignment");
  if (n == 0) return FPV_OK;
  fpv::gate_env_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*p, (const float4*)state, n, plane_stride,
                                                                                              mask, (float2*)prev, progress);
  return check_launch("fpv_gate_env_reset");
}

int fpv_gate_env_step(const fpv_gate_env_params_t* p, const void* state, int64_t n, int64_t plane_stride,
                      const uint8_t* agent_done, void* prev, int32_t* progress, float* agent_reward, float* env_reward,
                      uint8_t* env_done, float* obs, fpv_stats_t* stats, void* stream) {
  if (int rc = check_gate_params(p, n, "fpv_gate_env_step")) return rc;
  if (!state || !agent_done || !prev || !progress || !env_reward || !env_done)
    return fail(FPV_EINVAL, "fpv_gate_env_step: null pointer");
  if (plane_stride < n || !aligned16(state) || !aligned16(obs)) return fail(FPV_EINVAL, "fpv_gate_env_step: bad stride/alignment");
  if (n == 0) return FPV_OK;
  constexpr int T = 256;
  fpv::gate_env_step_kernel<T><<<(unsigned)((n + T - 1) / T), T, 0, (cudaStream_t)stream>>>(
      *p, (const float4*)state, n, plane_stride, agent_done, (float2*)prev, progress, agent_reward, env_reward, env_done,
      (float4*)obs, stats);
  return check_launch("fpv_gate_env_step");
}

int fpv_gate_race_step(const fpv_drone_params_t* p, const fpv_drone_io_t* io, const fpv_gate_env_params_t* gp, void* prev,
                       int32_t* progress, float* agent_reward, float* env_reward, uint8_t* env_done, float* obs,
                       void* stream) {
  DroneK k;
  DroneIO d;
  int ang = 0;
  bool general = false;
  if (!io) return fail(FPV_EINVAL, "fpv_gate_race_step: null io");
  if (int rc = check_gate_params(gp, io->n, "fpv_gate_race_step")) return rc;
  const int rc = prepare_drone(p, io, true, k, d, ang, general);
  if (rc != FPV_OK) return rc > 0 ? FPV_OK : rc;
  if (!prev || !progress || !env_reward || !env_done) return fail(FPV_EINVAL, "fpv_gate_race_step: null pointer");
  if (!aligned16(obs)) return fail(FPV_EINVAL, "fpv_gate_race_step: obs must be 16-byte aligned");
  const bool wind = p->wind[0] != 0.f || p->wind[1] != 0.f || p->wind[2] != 0.f;
  if (general || io->wind_env || wind || (p->flags & FPV_F_FREEZE_DONE))
    return fail(FPV_EINVAL, "fpv_gate_race_step: only the hot-path configuration is supported (no obstacles, overrides, wind "
                            "or FPV_F_FREEZE_DONE); call fpv_drone_step and fpv_gate_env_step instead");
  if (p->flags & FPV_F_SCALAR) return fail(FPV_EINVAL, "fpv_gate_race_step: the fused step runs the packed kernel (no FPV_F_SCALAR)");
  fpv::GateIO g;
  static_cast<DroneIO&>(g) = d;       // chunk_epoch / FPV_F_CHAINED carry over: the env arrays are per agent, i.e. per chunk,
  g.gp = *gp;                         // and are read after / written before the chunk's epoch like the state itself
  g.prev = (float2*)prev;
  g.progress = progress;
  g.agent_reward = agent_reward;
  g.env_reward = env_reward;
  g.env_done = env_done;
  g.obs = (float4*)obs;
  const size_t stage = ((k.flags & FPV_F_THRUST_LUT) ? ((size_t)k.lut_n * sizeof(float) + 127) / 128 * 128 : 0) +
                       (size_t)gp->n_gates * sizeof(fpv_gate_t);
  auto launch = [&](auto mode_tag) {
    using Mode = decltype(mode_tag);
    return launch_ring<Mode, FPV_MINB>(k, g, stage, d.cta_cap, (k.flags & FPV_F_CHAINED) != 0, &k.flags, FPV_F_CHAINED,
                                       (cudaStream_t)stream);
  };
  bool ok;
  if (ang == 4) ok = launch(fpv::DroneMode<F2, 4, false, fpv::GatePost, fpv::GateIO>{});
  else if (ang == 3) ok = launch(fpv::DroneMode<F2, 3, false, fpv::GatePost, fpv::GateIO>{});
  else if (ang == 2) ok = launch(fpv::DroneMode<F2, 2, false, fpv::GatePost, fpv::GateIO>{});
  else if (ang == 1) ok = launch(fpv::DroneMode<F2, 1, false, fpv::GatePost, fpv::GateIO>{});
  else ok = launch(fpv::DroneMode<F2, 0, false, fpv::GatePost, fpv::GateIO>{});
  if (!ok) return fail(FPV_EINVAL, "fpv_gate_race_step: the motor-curve table does not fit next to the ring in shared memory");
  return check_launch("fpv_gate_race_step");
}

namespace {
int make_cam(const fpv_camera_params_t* c, fpv::CamK& k, const char* who) {
  if (!c) return fail(FPV_EINVAL, "%s: null camera params", who);
  if (c->width < 1 || c->height < 1) return fail(FPV_EINVAL, "%s: bad resolution %dx%d", who, c->width, c->height);
  if (!(c->fx != 0.0) || !(c->fy != 0.0)) return fail(FPV_EINVAL, "%s: focal length must be non-zero", who);
  for (int i = 0; i < 9; ++i) k.rel_rot[i] = c->rel_rot[i];
  for (int i = 0; i < 3; ++i) k.rel_pos[i] = c->rel_pos[i];
  k.fx = c->fx; k.fy = c->fy; k.cx = c->cx; k.cy = c->cy;
  k.W = c->width; k.H = c->height;
  return FPV_OK;
}
int check_world(const double* pose, int64_t n, const double* points, int32_t n_points, const double* boxes,
                int32_t n_objects, const char* who) {
  if (!pose || !points || !boxes) return fail(FPV_EINVAL, "%s: null pointer", who);
  if (n < 0 || n_points < 0) return fail(FPV_EINVAL, "%s: bad n / n_points", who);
  if (n_objects < 1 || n_objects > FPV_CAM_MAX_OBJECTS)
    return fail(FPV_EINVAL, "%s: n_objects=%d out of [1,%d]", who, n_objects, FPV_CAM_MAX_OBJECTS);
  if (!aligned16(points)) return fail(FPV_EINVAL, "%s: points must be 16-byte aligned", who);
  return FPV_OK;
}
}  // namespace

int fpv_camera_update(const fpv_camera_params_t* cam, const void* state, int64_t n, int64_t plane_stride, double* pose,
                      void* stream) {
  fpv::CamK k;
  if (int rc = make_cam(cam, k, "fpv_camera_update")) return rc;
  if (!state || !pose) return fail(FPV_EINVAL, "fpv_camera_update: null pointer");
  if (n < 0 || plane_stride < n || !aligned16(state)) return fail(FPV_EINVAL, "fpv_camera_update: bad n/stride/alignment");
  if (n == 0) return FPV_OK;
  fpv::camera_update_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(k, (const float4*)state, n, plane_stride, pose);
  return check_launch("fpv_camera_update");
}

int fpv_camera_update_pose(const fpv_camera_params_t* cam, const double* pos, const double* rot, int64_t n, double* pose,
                           void* stream) {
  fpv::CamK k;
  if (int rc = make_cam(cam, k, "fpv_camera_update_pose")) return rc;
  if (!pos || !rot || !pose) return fail(FPV_EINVAL, "fpv_camera_update_pose: null pointer");
  if (n < 0) return fail(FPV_EINVAL, "fpv_camera_update_pose: bad n");
  if (n == 0) return FPV_OK;
  fpv::camera_update_pose_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(k, pos, rot, n, pose);
  return check_launch("fpv_camera_update_pose");
}

int fpv_camera_render(const fpv_camera_params_t* cam, const double* pose, int64_t n, const double* points,
                      int32_t n_points, const double* boxes, int32_t n_objects, const double* obj_offset, double max_depth,
                      uint8_t* keep, uint8_t* image, void* stream) {
  fpv::CamK k;
  if (int rc = make_cam(cam, k, "fpv_camera_render")) return rc;
  if (int rc = check_world(pose, n, points, n_points, boxes, n_objects, "fpv_camera_render")) return rc;
  if (!keep || !image) return fail(FPV_EINVAL, "fpv_camera_render: null keep / image");
  if (((int64_t)k.W * k.H) % 4 != 0 || (reinterpret_cast<uintptr_t>(image) & 3u))
    return fail(FPV_EINVAL, "fpv_camera_render: width*height must be a multiple of 4 and the image 4-byte aligned");
  if (n > 65535) return fail(FPV_EINVAL, "fpv_camera_render: at most 65535 cameras per call (got %lld)", (long long)n);
  if (n == 0) return FPV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(image, 0, (size_t)n * k.W * k.H, st);
  const long long no = (long long)n * n_objects;
  fpv::camera_prune_kernel<<<(unsigned)((no + 127) / 128), 128, 0, st>>>(k, pose, n, boxes, n_objects, obj_offset, keep);
  if (n_points > 0) {
    dim3 grid((unsigned)((n_points + 255) / 256), (unsigned)n);
    fpv::camera_splat_kernel<<<grid, 256, 0, st>>>(k, pose, (const double4*)points, n_points, n_objects, obj_offset, keep,
                                                   max_depth, image);
  }
  return check_launch("fpv_camera_render");
}

int fpv_camera_target_pixel(const fpv_camera_params_t* cam, const double* pose, int64_t n, const double* points,
                            int32_t n_points, const double* boxes, int32_t n_objects, const double* obj_offset,
                            double max_depth, double* pixel, uint8_t* seen, void* stream) {
  fpv::CamK k;
  if (int rc = make_cam(cam, k, "fpv_camera_target_pixel")) return rc;
  if (int rc = check_world(pose, n, points, n_points, boxes, n_objects, "fpv_camera_target_pixel")) return rc;
  if (!pixel || !seen) return fail(FPV_EINVAL, "fpv_camera_target_pixel: null pixel / seen");
  if (!(max_depth > 0.0)) return fail(FPV_EINVAL, "fpv_camera_target_pixel: max_depth must be positive");
  const size_t smem = (((size_t)k.W * k.H + 31) / 32) * sizeof(unsigned);
  if (smem > 200 * 1024) return fail(FPV_EINVAL, "fpv_camera_target_pixel: %dx%d frame bitmap does not fit in shared memory", k.W, k.H);
  if (n == 0) return FPV_OK;
  static SmemOptIn opted;
  opt_in_smem(fpv::camera_target_pixel_kernel, opted, smem);
  fpv::camera_target_pixel_kernel<<<(unsigned)n, 256, smem, (cudaStream_t)stream>>>(
      k, pose, (const double4*)points, n_points, n_objects, boxes, obj_offset, max_depth, pixel, seen);
  return check_launch("fpv_camera_target_pixel");
}

int fpv_camera_rays(const fpv_camera_params_t* cam, const double* pose, int64_t n, const double* pixel, int32_t frame,
                    double* dir, void* stream) {
  fpv::CamK k;
  if (int rc = make_cam(cam, k, "fpv_camera_rays")) return rc;
  if (!pose || !pixel || !dir) return fail(FPV_EINVAL, "fpv_camera_rays: null pointer");
  if (frame < 0 || frame > 2) return fail(FPV_EINVAL, "fpv_camera_rays: ref_frame must be world, drone or camera");
  if (n < 0) return fail(FPV_EINVAL, "fpv_camera_rays: bad n");
  if (n == 0) return FPV_OK;
  fpv::camera_rays_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(k, pose, n, pixel, frame, dir);
  return check_launch("fpv_camera_rays");
}

namespace {
int make_autopilot(const fpv_autopilot_params_t* ap, fpv::AutopilotK& a, const char* who) {
  if (!ap) return fail(FPV_EINVAL, "%s: null params", who);
  if (ap->ref_frame < 0 || ap->ref_frame > 1) return fail(FPV_EINVAL, "%s: Unknown reference frame", who);
  if (ap->mode < 0 || ap->mode > 1) return fail(FPV_EINVAL, "%s: Unknown mode", who);
  if (!(ap->dt > 0.0)) return fail(FPV_EINVAL, "%s: dt must be positive", who);
  a.mass = ap->mass; a.dt = ap->dt; a.vdrag_coef = ap->virtual_drag_coef; a.vlift_coef = ap->virtual_lift_coef;
  a.tof_dist = ap->tof_effective_dist; a.keep_distance = ap->keep_distance; a.uwb_max = ap->uwb_max_range;
  a.kP = ap->kP; a.kI = ap->kI; a.kD = ap->kD; a.integral_clip = ap->integral_clip; a.min_out = ap->min_output;
  a.max_out = ap->max_output; a.dtr = ap->derivative_transition_rate; a.ref_frame = ap->ref_frame; a.mode = ap->mode;
  a.max_force = ap->max_throttle_force;
  a.max_iter = ap->max_limit_iterations > 0 ? ap->max_limit_iterations : 64;
  return FPV_OK;
}
}  // namespace

int fpv_autopilot(const fpv_autopilot_params_t* ap, const fpv_camera_params_t* cam, const void* state, int64_t n,
                  int64_t plane_stride, const double* pixel, const uint8_t* seen, const double* target_pos,
                  const double* target_radius, double* pid, float* rot, void* quat, float* force, void* stream) {
  fpv::CamK k;
  fpv::AutopilotK a;
  if (int rc = make_cam(cam, k, "fpv_autopilot")) return rc;
  if (int rc = make_autopilot(ap, a, "fpv_autopilot")) return rc;
  if (!state || !pixel || !target_pos || !target_radius || !pid) return fail(FPV_EINVAL, "fpv_autopilot: null pointer");
  if (n < 0 || plane_stride < n || !aligned16(state) || !aligned16(quat)) return fail(FPV_EINVAL, "fpv_autopilot: bad n/stride/alignment");
  if (n == 0) return FPV_OK;
  fpv::autopilot_kernel<0><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      a, k, (const float4*)state, n, plane_stride, pixel, seen, target_pos, target_radius, nullptr, pid, rot, (float4*)quat, force, nullptr);
  return check_launch("fpv_autopilot");
}

int fpv_point_and_shoot(const fpv_autopilot_params_t* ap, const fpv_camera_params_t* cam, const void* state, int64_t n,
                        int64_t plane_stride, const double* pixel, const double* action, const uint8_t* seen, double* pid,
                        float* rot, void* quat, float* force, double* shifted_pixel, void* stream) {
  fpv::CamK k;
  fpv::AutopilotK a;
  if (int rc = make_cam(cam, k, "fpv_point_and_shoot")) return rc;
  if (int rc = make_autopilot(ap, a, "fpv_point_and_shoot")) return rc;
  if (!state || !pixel || !action || !pid) return fail(FPV_EINVAL, "fpv_point_and_shoot: null pointer");
  if (n < 0 || plane_stride < n || !aligned16(state) || !aligned16(quat)) return fail(FPV_EINVAL, "fpv_point_and_shoot: bad n/stride/alignment");
  if (!(ap->max_throttle_force > 0.0)) return fail(FPV_EINVAL, "fpv_point_and_shoot: max_throttle_force must be positive");
  if (n == 0) return FPV_OK;
  fpv::autopilot_kernel<1><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      a, k, (const float4*)state, n, plane_stride, pixel, seen, nullptr, nullptr, action, pid, rot, (float4*)quat, force, shifted_pixel);
  return check_launch("fpv_point_and_shoot");
}

int fpv_acro_reset(void* state, int64_t n, int64_t plane_stride, const float* pos, const float* vel,
                   const float* rpy_deg, const uint8_t* mask, void* stream) {
  if (!state || !pos || !vel || !rpy_deg) return fail(FPV_EINVAL, "fpv_acro_reset: null pointer");
  if (n < 0 || plane_stride < n || !aligned16(state)) return fail(FPV_EINVAL, "fpv_acro_reset: bad n/stride/alignment");
  if (n == 0) return FPV_OK;
  fpv::acro_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((float4*)state, n, plane_stride, pos, vel, rpy_deg, mask);
  return check_launch("fpv_acro_reset");
}

namespace {
int acro_launch(const fpv_acro_params_t* p, void* state, int64_t n, int64_t plane_stride, const void* actions,
                const float* lut, int32_t lut_n, uint8_t* done, void* motor_thrust, const void* reset_state,
                fpv_stats_t* stats, int32_t T, int64_t act_stride, uint8_t* done_seq, int64_t done_stride, void* work,
                void* stream) {
  if (!p || !state || !actions) return fail(FPV_EINVAL, "fpv_acro_step: null pointer");
  if (n < 0 || plane_stride < n) return fail(FPV_EINVAL, "fpv_acro_step: bad n/stride");
  if (!aligned16(state) || !aligned16(actions) || !aligned16(motor_thrust) || !aligned16(reset_state))
    return fail(FPV_EINVAL, "fpv_acro_step: float4 planes must be 16-byte aligned");
  if (p->substeps < 1 || !(p->dt > 0.f) || !(p->mass > 0.f)) return fail(FPV_EINVAL, "fpv_acro_step: bad dt/mass/substeps");
  if ((p->flags & FPV_F_AUTO_RESET) && !reset_state) return fail(FPV_EINVAL, "fpv_acro_step: FPV_F_AUTO_RESET needs reset_state");
  if (p->flags & FPV_F_THRUST_LUT) {
    if (!lut || lut_n < 2) return fail(FPV_EINVAL, "fpv_acro_step: FPV_F_THRUST_LUT needs lut with lut_n >= 2");
    if ((size_t)lut_n * sizeof(float) > 200 * 1024) return fail(FPV_EINVAL, "fpv_acro_step: lut_n=%d does not fit in shared memory", lut_n);
  }
  if (!(p->u_min < p->u_max)) return fail(FPV_EINVAL, "fpv_acro_step: u_min must be below u_max");
  if ((p->flags & FPV_F_THRUST_LUT) && !(p->u_min >= -1.f && p->u_max <= 1.f))
    return fail(FPV_EINVAL, "fpv_acro_step: with FPV_F_THRUST_LUT the motor throttle limits must lie in [-1, 1] (the table's range)");
  fpv::AcroK k;
  std::memset(&k, 0, sizeof(k));
  k.dt = p->dt; k.inv_dt = (float)(1.0 / (double)p->dt); k.substeps = p->substeps;
  k.max_rates = p->max_rates;
  k.rtr = p->rates_transition_rate; k.one_minus_rtr = (float)(1.0 - (double)p->rates_transition_rate);
  k.ttr = p->thrust_transition_rate; k.one_minus_ttr = (float)(1.0 - (double)p->thrust_transition_rate);
  k.deg2rad = (float)0.017453292519943295;
  for (int i = 0; i < 3; ++i) {
    if (!(p->inertia[i] > 0.f)) return fail(FPV_EINVAL, "fpv_acro_step: inertia must be positive");
    for (int j = 0; j < 3; ++j) k.gains[i][j] = p->gains[i][j];
    const double ki = p->gains[i][1] > 1e-12f ? (double)p->gains[i][1] : 1e-12;
    k.i_lim[i] = (float)((double)p->integral_limit / ki);
    k.inertia[i] = p->inertia[i]; k.inv_inertia[i] = (float)(1.0 / (double)p->inertia[i]);
    k.kd[i] = p->k_drag[i]; k.wind[i] = p->wind[i];
    k.rc_cen[i] = p->rate_curve[i][0];
    k.rc_span[i] = (float)std::fmax(0.0, (double)p->rate_curve[i][1] - (double)p->rate_curve[i][0]);
    k.rc_expo[i] = p->rate_curve[i][2];
    if ((p->flags & FPV_F_RATE_CURVE) && !(p->rate_curve[i][2] >= 0.f && p->rate_curve[i][2] <= 1.f))
      return fail(FPV_EINVAL, "fpv_acro_step: rate_curve expo must be in [0,1]");
  }
  for (int m = 0; m < 4; ++m) {
    const float x = p->motor_xy[m][0], y = p->motor_xy[m][1];
    k.motor_xy[m][0] = x; k.motor_xy[m][1] = y;
    k.mix[m][0] = y > 0.f ? 1.f : (y < 0.f ? -1.f : 0.f);
    k.mix[m][1] = x > 0.f ? -1.f : (x < 0.f ? 1.f : 0.f);
    k.mix[m][2] = p->spin[m];
    k.spin_kappa[m] = p->spin[m] * p->kappa;
  }
  k.kd_a = (float)((double)p->k_drag[1] - (double)p->k_drag[0]);
  k.kd_b = (float)((double)p->k_drag[2] - (double)p->k_drag[0]);
  k.u_min = p->u_min; k.u_max = p->u_max;
  for (int i = 0; i < 4; ++i) k.poly[i] = p->thrust_poly[i];
  k.lut_n = (p->flags & FPV_F_THRUST_LUT) ? lut_n : 0;
  k.lut_scale = (float)((lut_n - 1) * 0.5);
  k.grav_z = (float)(-(double)p->gravity * (double)p->mass);
  k.inv_mass = (float)(1.0 / (double)p->mass);
  k.motor_radius = p->motor_radius; k.spring_k = p->spring_k;
  k.flags = p->flags;
  if (n == 0) return FPV_OK;
  const size_t smem = (p->flags & FPV_F_THRUST_LUT) ? sizeof(float) * ((size_t)lut_n + 1) : 0;   // + the padding entry
  auto launch = [&](auto kern, int envs_per_thread) {
    static SmemOptIn opted;
    opt_in_smem(kern, opted, smem);
    const long long tile = (long long)kThreads * envs_per_thread;
    kern<<<(unsigned)((n + tile - 1) / tile), kThreads, smem, (cudaStream_t)stream>>>(
        k, (float4*)state, n, plane_stride, (const float4*)actions, lut, done, (float4*)motor_thrust, (const float4*)reset_state, stats,
        (int)T, (long long)act_stride, done_seq, (long long)done_stride);
  };
  if (T == 1 && !(p->flags & FPV_F_SCALAR)) {   // one control step, packed: the ring kernel
    fpv::AcroIO io = {};
    io.state = (float4*)state;
    io.n = n;
    io.stride = plane_stride;
    io.actions = (const float4*)actions;
    io.lut = lut;
    io.done = done;
    io.motor_out = (float4*)motor_thrust;
    io.reset_state = (const float4*)reset_state;
    io.stats = stats;
    io.work = p->substeps >= 4 ? (unsigned*)work : nullptr;
    if (launch_ring<fpv::AcroMode<F2>, 3>(k, io, smem, 0, false, nullptr, 0, (cudaStream_t)stream)) return check_launch("fpv_acro_step");
  }
  if (p->flags & FPV_F_SCALAR) launch(fpv::acro_step_kernel<float, kThreads>, 1);
  else launch(fpv::acro_step_kernel<F2, kThreads>, 2);
  return check_launch("fpv_acro_step");
}
}  // namespace

int fpv_acro_step(const fpv_acro_params_t* p, void* state, int64_t n, int64_t plane_stride, const void* actions,
                  const float* lut, int32_t lut_n, uint8_t* done, void* motor_thrust, const void* reset_state,
                  fpv_stats_t* stats, void* work, void* stream) {
  return acro_launch(p, state, n, plane_stride, actions, lut, lut_n, done, motor_thrust, reset_state, stats, 1, 0, nullptr, 0, work,
                     stream);
}

int fpv_acro_rollout(const fpv_acro_params_t* p, void* state, int64_t n, int64_t plane_stride, const void* actions_seq,
                     int64_t action_stride, int32_t n_steps, const float* lut, int32_t lut_n, uint8_t* done_seq,
                     int64_t done_stride, uint8_t* done_last, void* motor_thrust, const void* reset_state, fpv_stats_t* stats,
                     void* stream) {
  if (n_steps < 0) return fail(FPV_EINVAL, "fpv_acro_rollout: n_steps must be >= 0");
  if (n_steps == 0) return FPV_OK;
  if (action_stride < n) return fail(FPV_EINVAL, "fpv_acro_rollout: action_stride must be >= n");
  if (done_seq && done_stride < n) return fail(FPV_EINVAL, "fpv_acro_rollout: done_stride must be >= n");
  return acro_launch(p, state, n, plane_stride, actions_seq, lut, lut_n, done_last, motor_thrust, reset_state, stats, n_steps,
                     action_stride, done_seq, done_stride, nullptr, stream);
}

}  // extern "C"
