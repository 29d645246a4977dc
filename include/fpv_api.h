/* fpv_api.h -- C ABI of libfpyv_b200.so: the batched FPV-drone dynamics step for NVIDIA B200
 * (sm_100a).  This is the drop-in boundary for the dynamics path of omrijsharon/FpyV.
 *
 * The reference has no FFI for this path: its boundary is the Python object protocol of
 * `utils.components.Drone` (src/utils/components.py:72-253) plus the free functions of
 * src/utils/kinematics.py.  Each entry point below names the reference interface it replaces.
 * The Python mirror of that protocol lives in fpyv_b200/drone.py and binds these symbols with
 * ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - Every function returns 0 on success and a negative FPV_E* code otherwise; nothing throws
 *     across the ABI.  fpv_last_error() returns a thread-local message for the last failure.
 *   - All data pointers are DEVICE pointers owned by the caller (PyTorch owns the memory; the
 *     library only borrows them for the launch) unless an argument is named *_host.  Parameter
 *     structs are HOST pointers to PODs that are copied by value into the launch (kernel-parameter
 *     constant bank).  The library keeps no state about the simulation, so it is re-entrant and
 *     usable from one host thread per GPU; the only things it caches are occupancy numbers per
 *     kernel and device and, for fpv_drone_step_host, two copy streams + events per device.
 *   - Like the CUDA runtime, every call launches on the CURRENT device (cudaSetDevice): the
 *     buffers and `stream` must belong to it.  One process may drive several devices by switching
 *     the current device between calls.
 *   - Launches are asynchronous and ordered on `stream` (a cudaStream_t / CUstream handle; 0 =
 *     legacy default stream).  No hidden synchronisation, no allocation.  Two documented
 *     relaxations of plain stream order exist, both opt-in: FPV_F_CHAINED (a launch may overlap the
 *     end of the previous one; the state is ordered per 64-env chunk instead) and
 *     fpv_drone_step_host (internal copy streams, joined back to `stream`).
 *   - "float4 plane" = an array of n 16-byte elements, 16-byte aligned.  Env i lives at
 *     element i of every plane.  Planes of one state buffer are `plane_stride` ELEMENTS apart.
 *   - Arithmetic is FP32 (the reference is float64 NumPy); parity tolerance is stated in
 *     tests/test_gpu_parity.py (<= 1e-5 relative per step).
 */
#ifndef FPV_API_H
#define FPV_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPV_ABI_VERSION 13

/* error codes */
#define FPV_OK 0
#define FPV_EINVAL (-22)  /* bad argument (null/misaligned pointer, bad size, bad flag combo) */
#define FPV_ECUDA (-5)    /* CUDA runtime error at launch; see fpv_last_error()               */
#define FPV_ENODEV (-19)  /* no sm_100 device / kernel image not loadable on this device      */

/* ---------------------------------------------------------------------------------------------
 * Drone state layout (mode A, reference `Drone`): 4 float4 planes = 64 B per env
 *   plane 0: position.x  position.y  position.z  prev_thrust      components.py:151-153,:161
 *   plane 1: velocity.x  velocity.y  velocity.z  episode (int32 bits: control steps since reset; -1-steps
 *                                                 once the env has crashed and is frozen)
 *   plane 2: attitude as a unit quaternion  w x y z               components.py:154 (rotation_matrix)
 *   plane 3: prev_rates[0..2] (deg/s, components.py:160), one spare float preserved by the step
 * The reference stores the body->world rotation MATRIX and multiplies it by an Euler increment twice per
 * step (components.py:216-218, kinematics.py:27-30).  The same rotation is carried here as the quaternion
 * of the reference's own convention (helper_functions.py:65-117): R(q1 q2) = R(q1) R(q2), so the update is a
 * quaternion product; fpv_drone_get_rotation / fpv_drone_set_rotation convert at the boundary.
 * Consequence: a rotation override must be a proper rotation (the reference would accept any 3x3).
 * -------------------------------------------------------------------------------------------*/
#define FPV_DRONE_PLANES 4

/* flags for fpv_drone_params_t.flags */
#define FPV_F_GROUND 1u         /* plane z=0 in the object list (components.py:649-680)          */
#define FPV_F_AUTO_RESET 2u     /* an env whose control step raised `done` is reloaded from
                                   io.reset_state at the end of that control step              */
#define FPV_F_FREEZE_DONE 4u    /* a crashed env stops integrating (the reference caller breaks
                                   out of its loop, src/core/simulator.py:91-93); default is the
                                   reference's own behaviour: done is reported, integration goes on */
#define FPV_F_THRUST_LUT 8u     /* throttle->thrust through io.lut (shared-memory table, linear
                                   interpolation) instead of the cubic                          */
#define FPV_F_CHAINED 64u       /* this launch may overlap the END of the previous launch on the stream: it does not
                                   wait for the previous grid (programmatic dependent launch) and orders its accesses
                                   to `state` chunk by chunk through io.chunk_epoch instead.  The caller promises that
                                   (1) every input other than `state` (actions, lut, wind, reset_state) is not being
                                   written by work that may still be running, and (2) the last writer of `state` was a
                                   fpv_drone_step launch with the same chunk_epoch and epoch - 1.  Ignored (plain stream
                                   order) whenever the launch cannot honour it.  Do not capture a chained launch in a
                                   CUDA graph (the replay would carry a stale epoch).  A chunk that does not reach the
                                   expected epoch within 2 s of wall-clock time does not hang or poison the context:
                                   the launch adds 1 to work[16], falls back to waiting for the whole previous grid
                                   and goes on; the host reads the word when it wants to know (BatchedDrone.
                                   episode_stats()["chain_timeouts"]).                                            */
#define FPV_F_RATE_CURVE 128u    /* fpv_acro_params_t.rate_curve is used (mode C only)                            */
#define FPV_F_SCALAR 32u        /* one env per thread (plain FP32 instructions) instead of the
                                   default two envs per thread on packed f32x2 instructions     */

/* Mirrors what Drone.__init__ derives from params.yaml (components.py:84-142). */
typedef struct fpv_drone_params {
  float dt;               /* seconds per substep (reference: 1/fps, components.py:96)            */
  int32_t substeps;       /* K >= 1 reference steps per launch, same action (north_star)         */
  float gravity;          /* m/s^2, components.py:92                                             */
  float mass;             /* kg, components.py:97                                                */
  float max_rates;        /* deg/s, components.py:85                                             */
  float rates_transition_rate;   /* components.py:105 */
  float thrust_transition_rate;  /* components.py:106 */
  float k_drag[3];        /* -0.5*Cd*rho*A per body axis, kinematics.py:36                       */
  float motor_xy[4][2];   /* body-frame motor offsets (z = 0), components.py:123-125             */
  float motor_radius;     /* components.py:121 */
  float spring_k;         /* components.py:198 */
  float spring_c;         /* components.py:198 */
  float thrust_poly[4];   /* cubic in throttle PERCENT, highest power first (model_xy, flight_time_calculator.py:43-52) */
  float wind[3];          /* uniform wind, world frame (step's wind_velocity_vector)             */
  uint32_t flags;
  int32_t n_objects;      /* extra obstacles after the ground plane, <= FPV_MAX_OBJECTS          */
} fpv_drone_params_t;

#define FPV_MAX_OBJECTS 16
#define FPV_OBJ_SPHERE 1    /* Target,   components.py:773-777 : x,y,z = centre, a = radius            */
#define FPV_OBJ_CYLINDER 2  /* Cylinder, components.py:710-729 : x,y,z = base centre, a = radius, b = height */
typedef struct fpv_object {
  int32_t kind;
  float x, y, z, a, b;
} fpv_object_t;

/* Episode statistics accumulated on the device (one struct per GPU; all-reduced by the host with
 * NCCL outside the step).  Doubles are updated with one atomicAdd per CTA. */
typedef struct fpv_stats {
  double env_steps;       /* control steps executed                                       */
  double crashes;         /* control steps that raised done                               */
  double episodes;        /* episodes ended (auto-reset or frozen)                        */
  double episode_len_sum; /* sum of lengths (control steps) of ended episodes              */
  double reward_sum;      /* env kernels only                                              */
  double reward_sq_sum;
  double nonfinite;       /* envs whose state went NaN/Inf this step                       */
  double reserved;
} fpv_stats_t;

/* Stick calibration (Joystick.calib_read, get_sticks.py:245-265) -- see "Stick front-end" below. */
typedef struct fpv_stick_calib {
  float min_vals[6], max_vals[6], sign_reverse[6];   /* calibration JSON arrays           */
  int32_t stick_idx[4];                              /* JSON "sticks" in file order: Throttle, Roll, Pitch, Yaw */
  float stick_center[4];
} fpv_stick_calib_t;

/* raw-stick transport formats (fpv_drone_io_t.stick_format, fpv_drone_step_host_sticks) */
#define FPV_STICKS_U16 1    /* uint16[n][4]: raw readings 0..65535 of axes 0, 1, 2, 5 (throttle, roll, pitch, yaw): 8 B per env */
#define FPV_STICKS_CRSF 2   /* what an RC link carries: four 11-bit channels (same order) packed little-endian into 6 B per
                               env; an 11-bit value v stands for the raw reading (v << 5) | (v >> 6)                      */

typedef struct fpv_drone_io {
  void* state;              /* float4[FPV_DRONE_PLANES][plane_stride], in/out                     */
  int64_t n;                /* number of envs                                                     */
  int64_t plane_stride;     /* elements between planes, >= n                                      */
  const void* actions;      /* float4[n]: roll, pitch, yaw, throttle in [-1,1] (components.py:220).  Like `sticks` this may be
                               a PINNED HOST pointer (UVA): the step's TMA engine then fetches the chunks straight over
                               PCIe while earlier chunks compute -- one launch instead of copy + launch ("zero copy").   */
  const void* sticks;       /* Drone.step(action=None): raw stick readings in `stick_format` instead of actions (then
                               `actions` may be NULL), calibrated inside the step with the arithmetic of
                               fpv_sticks_to_actions (bit-identical).  Device or pinned host memory, 16-byte aligned, padded
                               to a multiple of 16 bytes.  Hot-path configuration and packed kernel only.  NULL = actions. */
  const fpv_stick_calib_t* stick_calib;  /* HOST pointer; required with `sticks`                                */
  int32_t stick_format;     /* FPV_STICKS_U16 or FPV_STICKS_CRSF                                                  */
  const void* wind_env;     /* float4[n] per-env wind (xyz), or NULL -> params.wind               */
  const float* lut;         /* float[lut_n] thrust [N] sampled at throttle -1..1, or NULL         */
  int32_t lut_n;
  uint8_t* done;            /* uint8[n] out: 1 if any substep raised done (components.py:236-240); may be NULL */
  void* done_bits;          /* uint32[ceil(n / 32)] out or NULL: the same flags as a bitmask, bit (e % 32) of word e / 32
                               (written with one warp ballot per 32 envs; 1/8 of the bytes of `done` for a host that
                               only needs to know WHICH envs ended).  Needs `done` as well.                   */
  void* acc_out;            /* float4[n] out: world acceleration of the last substep (components.py:243); may be NULL */
  const void* reset_state;  /* float4[FPV_DRONE_PLANES][plane_stride]: source for FPV_F_AUTO_RESET */
  const void* override_q;   /* float4[n]: rotation override as quaternion (w,x,y,z), see fpv_matrix_to_quat
                               (step(rotation_matrix=, thrust_force=), components.py:230-232); NULL = none.
                               Requires substeps == 1 and override_thrust. */
  const float* override_thrust; /* float[n]: thrust_force of the same call; NaN = no override for that env */
  const fpv_object_t* objects; /* HOST pointer, params.n_objects entries, or NULL                  */
  fpv_stats_t* stats;       /* device, may be NULL                                                */
  void* work;               /* device uint32[32], zeroed ONCE by the caller.  [0..15]: chunk counters for dynamic load
                               balancing (warps pull the next 64-env chunk with one atomic); every launch leaves them
                               zeroed.  [16]: number of chained waits that timed out (see FPV_F_CHAINED), never reset
                               by the library.  NULL = static round-robin distribution, no error word.        */
  void* chunk_epoch;        /* device uint32[ceil(n / 64)], zeroed once by the caller, or NULL.  After the launch
                               every entry holds epoch + 1 (published chunk by chunk as the chunk's state is stored);
                               with FPV_F_CHAINED chunk c is loaded only once chunk_epoch[c] == epoch.            */
  uint32_t epoch;           /* number of fpv_drone_step launches already applied to `state` through chunk_epoch  */
  uint32_t max_ctas_per_sm; /* 0 = every CTA slot of every SM (default).  k > 0: the persistent grid takes at most k slots
                               per SM, so that CHAINED launches of INDEPENDENT batches stepped round-robin are resident
                               side by side (the idle tail and start-up of one launch are covered by the others' bulk);
                               do not combine with chaining consecutive steps of the SAME batch -- the trailing launch
                               would hold its slots while it waits for the leading one chunk by chunk.
                               The same field serves the CLOSED-LOOP form: split the population into 4 / k independent
                               parts, one CUDA stream each, plain (unchained) launches with max_ctas_per_sm = k --
                               the parts' launches then run side by side without anything being known ahead of time
                               (fpyv_b200.TwoStreamDrones; 37 us per 1,048,576-env K = 8 step on B200 against 44-48 us
                               for one launch at a time).                                                          */
  void* trace;              /* developer profiling hook: device uint64[3 * warps] receiving per-warp
                               (start ns, end ns, SM id) of the hot kernel; NULL in production           */
} fpv_drone_io_t;

int fpv_abi_version(void);
const char* fpv_last_error(void);

/* sizeof() of the ABI structs as this library was compiled, so that a foreign-language binding can
 * verify its own struct layout: which = 0 fpv_drone_params_t, 1 fpv_drone_io_t, 2 fpv_object_t,
 * 3 fpv_stats_t, 4 fpv_stick_calib_t, 5 fpv_racer_params_t, 6 fpv_gate_env_params_t, 7 fpv_camera_params_t,
 * 8 fpv_autopilot_params_t, 9 fpv_acro_params_t; -1 for an unknown index. */
int fpv_sizeof(int which);

/* Number of SMs / compute capability of `device`; used by hosts to size persistent launches. */
int fpv_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* Diagnostic: one launch of a pure-FMA kernel (SMs x 8 CTAs x 256 threads x iters x 16 independent chains with shared
 * multiplicand / addend; packed != 0: fma.rn.f32x2, else scalar fma.rn.f32).  *flop_out (host) receives the flops of the
 * launch; the caller times it with events.  bench.py reports the result as roofline.peak_measured next to the nominal
 * FP32 peak.  sink: device float[>= SMs * 8 * 256] (never written in practice). */
int fpv_probe_fp32(int32_t packed, int32_t iters, float* sink, int64_t sink_floats, double* flop_out, void* stream);

/* Drone.reset(position, velocity, ypr)  -- components.py:150-169.
 * pos, vel, rpy_deg: float[n][3] row-major (rpy in DEGREES, consumed as roll, pitch, yaw exactly like
 * the reference's `ypr` argument).  mask: uint8[n] or NULL; only envs with mask != 0 are reset. */
int fpv_drone_reset(void* state, int64_t n, int64_t plane_stride, const float* pos, const float* vel,
                    const float* rpy_deg, const uint8_t* mask, void* stream);

/* Drone.step(action, wind_velocity_vector, object_list, rotation_matrix=None, thrust_force=None)
 * -- components.py:220-248, i.e. action2force :179-196, calculate_drag kinematics.py:33-38,
 * gravity_vector :41-45, handle_collisions components.py:198-214 (+Ground :674-680, spring_force
 * kinematics.py:56-59), the crash test :239, the force sum :242-243 and update :216-218
 * (update_kinematic_step kinematics.py:15-24 + rotate_body_by_rates :27-30, applied twice).
 * Runs params->substeps reference steps per env with the state held in registers. */
int fpv_drone_step(const fpv_drone_params_t* params, const fpv_drone_io_t* io, void* stream);

/* Drone.step with HOST buffers (the call a CPU-side caller of the reference makes: NumPy in, flags out):
 * actions_host float[n][4] and done_host uint8[n] are host pointers (page-locked memory for full speed); io->actions
 * must point at a device staging buffer float4[n], io->done at a device uint8[n].  The batch is cut into `slices` env
 * ranges (multiples of 64 envs, at least 65,536 envs each; the last one short -- n/16 -- because everything after the
 * last H2D byte is exposed latency) pipelined over three streams -- H2D copy of slice c+1 | step of slice c | D2H copy
 * of slice c-1 -- because the call is PCIe-bound (16 B/env in, 1 B/env out).  Everything is ordered after the work already
 * queued on `stream`, and `stream` is joined to the last copy: synchronising it means done_host is valid.  The two
 * extra streams and the events are created once per device and cached inside the library (the only state it keeps).
 * io->chunk_epoch / FPV_F_CHAINED are ignored.  slices <= 0: 4.
 * With io->done_bits set, done_host receives the BITMASK instead (uint32[ceil(n / 32)], 1/8 of the bytes; done_host must
 * then be 4-byte aligned and hold ceil(n / 32) * 4 bytes).  done_host == NULL (only with io->done_bits): no device-to-host
 * copies at all -- io->done_bits then points at pinned HOST memory and the step kernels write the flag words there
 * themselves; synchronising `stream` means they have arrived. */
int fpv_drone_step_host(const fpv_drone_params_t* params, const fpv_drone_io_t* io, const float* actions_host,
                        uint8_t* done_host, int32_t slices, void* stream);

/* Open-loop rollout: T consecutive calls of Drone.step (components.py:220-248, params->substeps reference steps each)
 * in ONE launch, with every env's state held in registers from the first step to the last: the state planes are read
 * once and written once, per control step only the env's action is read and its done flag written (17 B/env/step instead
 * of 145).  Bit-identical to T calls of fpv_drone_step on the same inputs (episode counters, FPV_F_AUTO_RESET restarts
 * and statistics included).
 *   actions_seq: float4[T][action_stride] (step t of env e at t*action_stride + e; action_stride >= n);
 *   done_seq:    uint8[T][done_stride] out, or NULL; io->done (if set) receives the LAST step's flags;
 *   io->actions is ignored; io->work (uint32[16], zeroed once) is required.
 * Supported configuration: the hot path of fpv_drone_step (ground plane, no obstacles / overrides / per-env wind, packed
 * kernel) without FPV_F_FREEZE_DONE and without io->chunk_epoch; anything else returns FPV_EINVAL and the caller steps. */
int fpv_drone_rollout(const fpv_drone_params_t* params, const fpv_drone_io_t* io, const void* actions_seq,
                      int64_t action_stride, int32_t n_steps, uint8_t* done_seq, int64_t done_stride, void* stream);

/* Drone.rotation_matrix (components.py:154) read / write: R is float[n][9] row-major, body->world.
 * set: robust matrix->quaternion (all four Shepperd branches), normalised; mask as in fpv_drone_reset. */
int fpv_drone_get_rotation(const void* state, int64_t n, int64_t plane_stride, float* R, void* stream);
int fpv_drone_set_rotation(void* state, int64_t n, int64_t plane_stride, const float* R, const uint8_t* mask,
                           void* stream);
/* R float[n][9] -> q float4[n] (w,x,y,z); for override inputs. */
int fpv_matrix_to_quat(const float* R, int64_t n, void* q, void* stream);

/* The tuple Drone.step returns -- components.py:247-248: (R^T, euler_matrix(*rates), R @ acc).
 * Rt, gyro: float[n][9] row-major; accel: float[n][3].  Any of the three may be NULL. */
int fpv_drone_observe(const void* state, int64_t n, int64_t plane_stride, const void* acc /*float4[n]*/,
                      float* Rt, float* gyro, float* accel, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stick front-end: Joystick.calib_read + Drone.read_sticks
 * (src/utils/get_sticks.py:245-265, components.py:250-253).
 * -------------------------------------------------------------------------------------------*/

/* raw: int32[n][6] axis readings (dwXpos..dwVpos order, get_sticks.py:55-60).
 * actions: float4[n] = [-roll, pitch, yaw, throttle]; calibrated: float[n][6] or NULL. */
int fpv_sticks_to_actions(const fpv_stick_calib_t* calib, const int32_t* raw, int64_t n, void* actions,
                          float* calibrated, void* stream);

/* Drone.step(action=None, ...) with HOST buffers: the joystick path of the reference (components.py:227-228, :250-253 ->
 * get_sticks.py:254-265) in the compact transport form of a radio link.  sticks_host: uint16[n][4] = the raw readings of
 * axes 0, 1, 2 and 5 (throttle, roll, pitch, yaw of the stock calibrations -- the four of calib_read's six values that
 * read_sticks keeps), 0..65535 as the joystick driver reports them -- or, format FPV_STICKS_CRSF, the 6-byte packing of
 * four 11-bit channels; sticks_dev: device staging buffer of the same size; the
 * calibrated actions are written to io->actions (float4[n]) by the same per-axis arithmetic as fpv_sticks_to_actions
 * (bit-identical), then the step runs and the flags return like in fpv_drone_step_host.  8 B/env in, 1 B/env out. */
int fpv_drone_step_host_sticks(const fpv_drone_params_t* params, const fpv_drone_io_t* io, const fpv_stick_calib_t* calib,
                               const void* sticks_host, void* sticks_dev, int32_t format /* FPV_STICKS_U16 | FPV_STICKS_CRSF */,
                               uint8_t* done_host, int32_t slices, void* stream);

/* Page-locked host memory for the *_host entries and for zero-copy inputs.  write_combined != 0: the CPU writes it through
 * write-combining buffers (never cached, so the device's reads do not snoop the CPU caches; CPU reads of it are slow --
 * use it for producer-written inputs only).  fpv_host_free releases it. */
int fpv_host_alloc(int64_t bytes, int32_t write_combined, void** out);
int fpv_host_free(void* p);

/* ---------------------------------------------------------------------------------------------
 * Mode B: the acro rate-PID drone of tests/racer_drone_test.py (`PID` :11-32, `Racer` :68-103).
 * State: 5 float4 planes (80 B per env)
 *   0: position xyz, first-call flag of the PIDs (1.0 / 0.0)      :20, :47-51
 *   1: velocity xyz, angular_velocity[0]
 *   2: orientation as a unit quaternion w x y z (the reference keeps the 3x3 matrix and re-orthonormalises it every
 *      step through scipy's Rotation, :99; same rotation, see fpv_racer_observe)
 *   3: PID integral per axis, angular_velocity[1]
 *   4: PID last error per axis, angular_velocity[2]
 * Runs on the packed TMA-ring kernel (two envs per thread); sin/cos of the half angles are accurate for every
 * argument a finite-gain PID produces (|omega| ~ 80 rad in the reference's own demo).
 * -------------------------------------------------------------------------------------------*/
#define FPV_RACER_PLANES 5

typedef struct fpv_racer_params {
  float dt;              /* racer_drone_test.py:8  */
  int32_t substeps;
  float mass;            /* :82 */
  float inertia[3];      /* :83 (m r^2 on every axis in the reference) */
  float gains[3][3];     /* [axis roll/pitch/yaw][P,I,D], :113 */
  float vel_decay;       /* :102 (0.9) */
  uint32_t flags;        /* 0, or FPV_F_SCALAR (one env per thread: the cross-check instantiation) */
} fpv_racer_params_t;

/* Racer.reset -- :86-93 (mask as in fpv_drone_reset). */
int fpv_racer_reset(void* state, int64_t n, int64_t plane_stride, const uint8_t* mask, void* stream);

/* Racer.step(action) -- :95-103.  actions: float4[n] = [roll, pitch, yaw rate set-points, thrust N].
 * torque_out: float4[n] (last substep's PID output) or NULL.  work: device uint32[32] zeroed once by the caller (chunk
 * counters, as fpv_drone_io_t.work) or NULL. */
int fpv_racer_step(const fpv_racer_params_t* params, void* state, int64_t n, int64_t plane_stride,
                   const void* actions, void* torque_out, void* work, void* stream);

/* Racer.orientation (the 3x3 matrix of :73, float[n][9] row-major) and Racer.angular_velocity (float[n][3]) read from
 * the state planes; either output may be NULL. */
int fpv_racer_observe(const void* state, int64_t n, int64_t plane_stride, float* R, float* omega, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-agent gate-race environment (BASELINE.json configs[4]; API style of tests/ma_com_simple_env.py:17-57:
 * reset() -> obs, step(a) -> (obs, reward, done, {}) with one observation per agent and ONE reward / done per env).
 * Gate geometry is the reference's: a gate is a plane through `position` with normal = rotation_matrix[:,0]
 * (components.py:811-822), tracks come from generate_track (generators.py:7-18).
 * The reference has NO gate-passing reward anywhere: the reward / termination below are OUR
 * definition -- PARITY UNPINNED, checked only against our own float64 model (oracle/gate_env_oracle.py):
 *   agent a (next gate g): d = n_g.(p - c_g), r = |p - c_g|
 *   passed  = d_prev < 0 <= d  and  r^2 - d^2 <= half_size_g^2            (plane crossed inside the aperture)
 *   reward_a = w_gate*passed + w_progress*(r_prev - r) - w_crash*crashed_a
 *   passed -> g = (g+1) mod n_gates (a wrap counts a lap); crashed -> g = 0, laps = 0 (the drone step already
 *   re-spawned the agent when auto-reset is on); (d_prev, r_prev) are then re-based on the agent's next gate
 *   env reward = sum over its agents (warp shuffle reduction), env done = any agent crashed or reached `laps_to_finish`
 * agents_per_env must be a power of two <= 32: one warp (or an aligned sub-warp) per env.
 * -------------------------------------------------------------------------------------------*/
#define FPV_MAX_GATES 32
#define FPV_ENV_OBS_FLOATS 16
typedef struct fpv_gate {
  float cx, cy, cz;      /* Gate.position                      components.py:786 */
  float nx, ny, nz;      /* Gate.normal = rotation_matrix[:,0] components.py:807-809 */
  float half_size;       /* aperture half-width: Gate.size / 2 components.py:790 */
  float pad;
} fpv_gate_t;

typedef struct fpv_gate_env_params {
  int32_t n_gates;
  int32_t agents_per_env;
  int32_t laps_to_finish;  /* <= 0: never finishes */
  float w_gate, w_progress, w_crash;
  fpv_gate_t gates[FPV_MAX_GATES];
} fpv_gate_env_params_t;

/* (Re)base the per-agent race bookkeeping on gate 0 from the current positions.  prev: float2[n] (d_prev, r_prev);
 * progress: int32[n] (low 16 bits next gate, high 16 bits laps); mask: uint8[n] or NULL. */
int fpv_gate_env_reset(const fpv_gate_env_params_t* params, const void* state, int64_t n_agents, int64_t plane_stride,
                       const uint8_t* mask, void* prev, int32_t* progress, void* stream);

/* One env step AFTER fpv_drone_step: rewards, terminations, observations.
 * agent_done: uint8[n_agents] (the done output of fpv_drone_step); agent_reward: float[n_agents] or NULL;
 * env_reward: float[n_envs]; env_done: uint8[n_envs]; obs: float[n_agents][FPV_ENV_OBS_FLOATS] or NULL:
 *   [0:3] R^T (c_g - p)  [3:6] R^T n_g  [6:9] R^T v  [9:12] R^T e_z  [12:15] rates (deg/s)  [15] prev_thrust
 * stats (may be NULL): reward_sum / reward_sq_sum accumulate the env rewards. */
int fpv_gate_env_step(const fpv_gate_env_params_t* params, const void* state, int64_t n_agents, int64_t plane_stride,
                      const uint8_t* agent_done, void* prev, int32_t* progress, float* agent_reward, float* env_reward,
                      uint8_t* env_done, float* obs, fpv_stats_t* stats, void* stream);

/* fpv_drone_step and fpv_gate_env_step in ONE launch: the env step runs as the per-chunk epilogue of the packed TMA-ring
 * kernel (a warp's 64-agent chunk holds whole envs; team reward / termination by warp shuffle / ballot; every agent's
 * state is read once and written once).  Bit-identical to fpv_drone_step followed by fpv_gate_env_step.  Hot-path
 * configuration of the dynamics only (ground plane, no obstacles / overrides / per-env wind, no FPV_F_FREEZE_DONE, no
 * FPV_F_SCALAR); io->done receives the agents' crash flags, io->stats (may be NULL) the dynamics counters and the reward
 * sums (summed in a different order than by fpv_gate_env_step: equal to rounding). */
int fpv_gate_race_step(const fpv_drone_params_t* params, const fpv_drone_io_t* io, const fpv_gate_env_params_t* gates,
                       void* prev, int32_t* progress, float* agent_reward, float* env_reward, uint8_t* env_done, float* obs,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * Chase pipeline: the callers on either side of Drone.step in src/core/simulator.py:98-110 --
 *   target_img = drone.camera.render_depth_image([target], max_depth)      components.py:614-629
 *   pixel      = mean (x, y) of the non-zero pixels                        simulator.py:104-108
 *   rot, f     = drone.calculate_needed_force_orientation(pixel, target)   components.py:258-304 (+ PID :43-54)
 *   drone.step(action, wind, objects, rotation_matrix=rot, thrust_force=f) components.py:230-232
 * Geometry here is float64 on the device (the splat produces integer pixel indices and bytes by truncation, which
 * only reproduce the float64 reference if the projection is float64).  All arrays are device pointers.
 * -------------------------------------------------------------------------------------------*/
#define FPV_CAM_MAX_OBJECTS 64
typedef struct fpv_camera_params {   /* Camera.__init__, components.py:450-470 */
  double rel_rot[9];                 /* relative_rotation_matrix = WORLD2CAM^T Rx(pitch), row-major   :455 */
  double rel_pos[3];                 /* position_relative_to_frame                                    :452 */
  double fx, fy, cx, cy;             /* intrinsic_matrix(f, f, W/2, H/2)                              :469-470 */
  int32_t width, height;             /* resolution [W, H] */
} fpv_camera_params_t;

typedef struct fpv_autopilot_params { /* Drone.__init__, components.py:96-97, :113-118, :143-145 */
  double mass, dt;
  double virtual_drag_coef, virtual_lift_coef, tof_effective_dist;  /* params["point_and_shoot"] */
  double keep_distance, uwb_max_range;                              /* params["drone"]           */
  double kP, kI, kD, integral_clip, min_output, max_output, derivative_transition_rate;  /* force_multiplier_pid */
  int32_t ref_frame;                 /* 0 'world', 1 'drone'            components.py:270-279 */
  int32_t mode;                      /* 0 'level', 1 'frontarget'       components.py:295-301 */
  double max_throttle_force;         /* Drone.max_throttle_in_force (components.py:142): limit of point_and_shoot's loop */
  int32_t max_limit_iterations;      /* cap on that loop (the reference's has none and can spin forever); <= 0: 64 */
  int32_t reserved;
} fpv_autopilot_params_t;

/* Camera.update (components.py:501-503) for every env: pose[e] = { R_cam row-major [9], camera position [3] }. */
int fpv_camera_update(const fpv_camera_params_t* cam, const void* state, int64_t n, int64_t plane_stride, double* pose,
                      void* stream);

/* The same update from explicit poses (the literal call shape of components.py:501: update(drone_position,
 * drone_rotation_matrix)): pos double[n][3], rot double[n][9] row-major body->world. */
int fpv_camera_update_pose(const fpv_camera_params_t* cam, const double* pos, const double* rot, int64_t n, double* pose,
                           void* stream);

/* Camera.render_depth_image (max_depth > 0, components.py:614-629) or Camera.render_image (max_depth <= 0, :601-612),
 * including pruned_objects_list (:584-599), for n cameras looking at ONE shared world:
 *   points: double[n_points][4] = x, y, z, object index; boxes: double[n_objects][6] = min xyz, max xyz of each object's
 *   points (bbox3d, helper_functions.py:120-136); obj_offset: double[n][n_objects][3] per-env translation of each
 *   object (moving / per-env targets) or NULL; keep: uint8[n][n_objects] scratch (receives the prune flags);
 *   image: uint8[n][height][width] out (width*height must be a multiple of 4; zeroed by the call). */
int fpv_camera_render(const fpv_camera_params_t* cam, const double* pose, int64_t n, const double* points,
                      int32_t n_points, const double* boxes, int32_t n_objects, const double* obj_offset, double max_depth,
                      uint8_t* keep, uint8_t* image, void* stream);

/* simulator.py:104-108 fused with the target's depth image: pixel[e] = mean (x, y) over the distinct non-zero pixels
 * of render_depth_image(objects, max_depth), seen[e] = 0 when there is none (pixel then 0, 0).  The frame lives as a
 * bitmap in shared memory (width*height bits <= 200 KiB); the byte image is never written. */
int fpv_camera_target_pixel(const fpv_camera_params_t* cam, const double* pose, int64_t n, const double* points,
                            int32_t n_points, const double* boxes, int32_t n_objects, const double* obj_offset,
                            double max_depth, double* pixel, uint8_t* seen, void* stream);

/* Camera.pixel2direction (components.py:505-526): pixel double[n][2] -> unit vectors double[n][3];
 * frame 0 'world', 1 'drone', 2 'camera'. */
int fpv_camera_rays(const fpv_camera_params_t* cam, const double* pose, int64_t n, const double* pixel, int32_t frame,
                    double* dir, void* stream);

/* Drone.calculate_needed_force_orientation (components.py:258-304) for every env, reading position / velocity /
 * attitude from `state`.  pixel: double[n][2]; seen: uint8[n] or NULL (envs with seen == 0 are skipped: their PID
 * state is untouched and force = NaN, which fpv_drone_step reads as "no override for this env");
 * target_pos: double[n][3]; target_radius: double[n]; pid: double[n][4] in/out = integral, prev_derivative,
 * previous_error, is_first (reset = 0, 0, 0, 1; components.py:35-41).  Outputs (each may be NULL): rot float[n][9]
 * row-major (columns x, y, force direction), quat float4[n] of the same rotation (for io.override_q), force float[n]. */
int fpv_autopilot(const fpv_autopilot_params_t* ap, const fpv_camera_params_t* cam, const void* state, int64_t n,
                  int64_t plane_stride, const double* pixel, const uint8_t* seen, const double* target_pos,
                  const double* target_radius, double* pid, float* rot, void* quat, float* force, void* stream);

/* Drone.point_and_shoot (components.py:312-381) for every env.  pixel: double[n][2]; action: double[n][4] =
 * (target column, target row on the screen, virtual-target x / y offset), each in [-1, 1] (:316, :322-323, :383-387);
 * seen / pid / rot / quat / force as in fpv_autopilot; shifted_pixel: double[n][2] out (pixel + virtual target, the
 * value the reference stores in prev_pixel, :325-330) or NULL. */
int fpv_point_and_shoot(const fpv_autopilot_params_t* ap, const fpv_camera_params_t* cam, const void* state, int64_t n,
                        int64_t plane_stride, const double* pixel, const double* action, const uint8_t* seen, double* pid,
                        float* rot, void* quat, float* force, double* shifted_pixel, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Mode C ("acro"): stick -> rate set-point -> acro rate PID -> motor mixer -> per-motor thrust / torque from the
 * T-Motor F80 bench curve (shared-memory LUT) -> rigid-body rotation, feeding the reference's translational model.
 * PARITY UNPINNED: the reference has no such model (its Drone applies the commanded rates kinematically,
 * components.py:216-218, and has ONE scalar thrust, :133-137).  Definition and only oracle: oracle/acro_oracle.py.
 * Reused reference pieces: action2force's maps and low-passes (components.py:185-194), the PID form of
 * tests/racer_drone_test.py:22-32, the motor layout (components.py:120-125), the bench curve (:133-136), drag /
 * gravity / ground spring / crash (:233-243) and the translation order of kinematics.py:21-22.
 * State: float4[FPV_ACRO_PLANES][plane_stride]
 *   0 position xyz, filtered collective throttle   1 velocity xyz, episode counter (int bits)   2 quaternion wxyz
 *   3 filtered rate set-point xyz [deg/s], PID first-call flag   4 body rates xyz [rad/s]   5 PID integral xyz
 *   6 previous rate error xyz
 * -------------------------------------------------------------------------------------------*/
#define FPV_ACRO_PLANES 7
typedef struct fpv_acro_params {
  float dt;
  int32_t substeps;
  float gravity, mass;
  float max_rates, rates_transition_rate, thrust_transition_rate;   /* components.py:85, :105-106 */
  float k_drag[3];              /* -0.5 Cd rho A                                  kinematics.py:36 */
  float motor_xy[4][2];         /* body-frame motor offsets                      components.py:123-125 */
  float motor_radius, spring_k; /* components.py:121, :198 */
  float gains[3][3];            /* rows roll, pitch, yaw; columns P, I, D; output in throttle units */
  float integral_limit;         /* clamp on |kI * integral| (throttle units) */
  float inertia[3];             /* diagonal body inertia [kg m^2] */
  float kappa;                  /* rotor reaction torque per thrust [m] */
  float spin[4];                /* +1 / -1 spin direction of motor k */
  float u_min, u_max;           /* motor throttle limits in [-1, 1] (idle = 5 %: -0.9, components.py:138-139) */
  float thrust_poly[4];         /* 4-motor bench cubic in throttle percent        components.py:136 */
  float wind[3];
  uint32_t flags;               /* FPV_F_GROUND | FPV_F_AUTO_RESET | FPV_F_THRUST_LUT | FPV_F_SCALAR (one env per thread)
                                   | FPV_F_RATE_CURVE */
  float rate_curve[3][3];       /* FPV_F_RATE_CURVE: per axis (centre sensitivity [deg/s], maximum rate [deg/s], expo in
                                   [0,1]) of the flight-controller "actual rates" stick curve
                                     rate(s) = s c + max(0, m - c) |s| (s^5 e + s (1 - e)),   s = -stick clipped to [-1,1];
                                   without the flag the reference's linear map clip(-stick * max_rates) (components.py:185) */
} fpv_acro_params_t;

/* pos, vel, rpy_deg: float[n][3]; motors off (throttle -1), zero rates, fresh PID.  mask as in fpv_drone_reset. */
int fpv_acro_reset(void* state, int64_t n, int64_t plane_stride, const float* pos, const float* vel,
                   const float* rpy_deg, const uint8_t* mask, void* stream);

/* params->substeps model steps with the action [roll, pitch, yaw, throttle] in [-1,1]^4 held.  lut: float[lut_n]
 * 4-motor thrust [N] sampled uniformly at throttle -1..1 (FPV_F_THRUST_LUT) or NULL; done: uint8[n] or NULL;
 * motor_thrust: float4[n] out (per-motor thrust of the last substep) or NULL; reset_state: snapshot for
 * FPV_F_AUTO_RESET; stats: crashes / episodes / episode_len_sum, may be NULL. */
int fpv_acro_step(const fpv_acro_params_t* params, void* state, int64_t n, int64_t plane_stride, const void* actions,
                  const float* lut, int32_t lut_n, uint8_t* done, void* motor_thrust, const void* reset_state,
                  fpv_stats_t* stats, void* work /* uint32[32] zeroed once, or NULL */, void* stream);

/* Open-loop rollout of mode C: n_steps control steps in one launch, state in registers across them, bit-identical to
 * n_steps calls of fpv_acro_step (restarts from the snapshot included).  actions_seq: float4[n_steps][action_stride];
 * done_seq: uint8[n_steps][done_stride] or NULL; done_last: uint8[n] (flags of the last step) or NULL. */
int fpv_acro_rollout(const fpv_acro_params_t* params, void* state, int64_t n, int64_t plane_stride, const void* actions_seq,
                     int64_t action_stride, int32_t n_steps, const float* lut, int32_t lut_n, uint8_t* done_seq,
                     int64_t done_stride, uint8_t* done_last, void* motor_thrust, const void* reset_state, fpv_stats_t* stats,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FPV_API_H */
