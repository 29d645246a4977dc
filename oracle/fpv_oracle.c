/* TEST INFRASTRUCTURE ONLY -- plain-C float64 restatement of the reference's `Drone.step`
 * (ground-only object list), used (a) as a second, independent checker next to the NumPy
 * restatement in oracle/fpv_oracle.py and (b) as the CPU baseline timed by bench.py
 * (`cpu_baseline`, `--impl reference`).  Never linked into or called by the product.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks it against the golden vectors that
 * oracle/make_golden.py produced by executing the unmodified reference.
 *
 * Reference lines followed (relative to /root/reference):
 *   action2force            src/utils/components.py:179-196
 *   calculate_drag          src/utils/kinematics.py:33-38
 *   gravity_vector          src/utils/kinematics.py:41-45
 *   motors / collisions     src/utils/components.py:235-239, :198-214, Ground :674-680,
 *                           spring_force src/utils/kinematics.py:56-59
 *   force sum               src/utils/components.py:242-243
 *   update                  src/utils/components.py:216-218, src/utils/kinematics.py:15-30
 *   euler matrix            src/utils/helper_functions.py:19-44
 *
 * Build: gcc -O2 -fPIC -shared -o _build/libfpv_oracle.so fpv_oracle.c -lm   (oracle/Makefile)
 * Threading: none in C; oracle/c_oracle.py runs disjoint env ranges on a thread pool (ctypes drops the GIL).
 */
#include <math.h>
#include <stdint.h>

typedef struct {
  double dt, gravity, mass, max_rates, rtr, ttr;
  double k_drag[3];
  double motor_rel[4][3];
  double poly[4]; /* high -> low, in throttle percent */
  double motor_radius, spring_k, spring_c;
  int32_t ground;
} oracle_consts_t;

static void euler_matrix(double roll, double pitch, double yaw, double m[3][3]) {
  /* Rz(yaw) @ Ry(pitch) @ Rx(roll), helper_functions.py:39-44 */
  const double sr = sin(roll), cr = cos(roll), sp = sin(pitch), cp = cos(pitch), sy = sin(yaw), cy = cos(yaw);
  m[0][0] = cy * cp; m[0][1] = cy * sp * sr - sy * cr; m[0][2] = cy * sp * cr + sy * sr;
  m[1][0] = sy * cp; m[1][1] = sy * sp * sr + cy * cr; m[1][2] = sy * sp * cr - cy * sr;
  m[2][0] = -sp;     m[2][1] = cp * sr;                m[2][2] = cp * cr;
}

static void mul_r_et(double R[3][3], double E[3][3]) {
  /* (E @ R^T)^T = R @ E^T, kinematics.py:30 */
  double T[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[i][j] = R[i][0] * E[j][0] + R[i][1] * E[j][1] + R[i][2] * E[j][2];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = T[i][j];
}

/* One env, one reference step.  Returns this step's `done`. */
static int drone_step_one(const oracle_consts_t* c, double* pos, double* vel, double* Rflat, double* prev_rates,
                          double* prev_thrust, const double* action, const double* wind, double* acc_out) {
  double (*R)[3] = (double (*)[3])Rflat;
  double rates[3];
  for (int i = 0; i < 3; ++i) {
    double cmd = -action[i] * c->max_rates;
    if (cmd < -c->max_rates) cmd = -c->max_rates;
    if (cmd > c->max_rates) cmd = c->max_rates;
    rates[i] = cmd * c->rtr + prev_rates[i] * (1 - c->rtr);
    prev_rates[i] = rates[i];
  }
  const double pct = 100 * (action[3] + 1) / 2;
  const double f = ((c->poly[0] * pct + c->poly[1]) * pct + c->poly[2]) * pct + c->poly[3];
  const double thr = f * c->ttr + (*prev_thrust) * (1 - c->ttr);
  *prev_thrust = thr;
  double F[3];
  for (int i = 0; i < 3; ++i) F[i] = R[i][2] * thr;
  /* drag */
  double vs[3], vb[3], fb[3];
  for (int i = 0; i < 3; ++i) vs[i] = vel[i] + wind[i];
  const double nrm = sqrt(vs[0] * vs[0] + vs[1] * vs[1] + vs[2] * vs[2]);
  for (int j = 0; j < 3; ++j) vb[j] = R[0][j] * vs[0] + R[1][j] * vs[1] + R[2][j] * vs[2];
  for (int j = 0; j < 3; ++j) fb[j] = c->k_drag[j] * vb[j] * nrm;
  for (int i = 0; i < 3; ++i) F[i] += R[i][0] * fb[0] + R[i][1] * fb[1] + R[i][2] * fb[2];
  F[2] += -c->gravity * c->mass;
  /* motors and ground */
  int done = 0, any_below = 0;
  double mz[4], coll = 0.0;
  for (int m = 0; m < 4; ++m) {
    mz[m] = pos[2] + c->motor_rel[m][0] * R[2][0] + c->motor_rel[m][1] * R[2][1] + c->motor_rel[m][2] * R[2][2];
    if (mz[m] < 0) any_below = 1;
  }
  if (c->ground) {
    if (any_below) {
      done = 1; /* early return with the forces summed so far: none */
    } else {
      for (int m = 0; m < 4; ++m) {
        const double pen = mz[m] - c->motor_radius;
        if (pen < 0) coll += -c->spring_k * pen - c->spring_c * vel[2];
      }
    }
  }
  if (any_below) done = 1; /* components.py:239 */
  F[2] += coll;
  double a[3];
  for (int i = 0; i < 3; ++i) a[i] = F[i] / c->mass;
  for (int i = 0; i < 3; ++i) pos[i] += vel[i] * c->dt;
  for (int i = 0; i < 3; ++i) vel[i] += a[i] * c->dt;
  const double d2r = M_PI / 180.0;
  double E[3][3];
  euler_matrix(rates[0] * d2r * c->dt, rates[1] * d2r * c->dt, rates[2] * d2r * c->dt, E);
  mul_r_et(R, E);
  mul_r_et(R, E);
  if (acc_out) { acc_out[0] = a[0]; acc_out[1] = a[1]; acc_out[2] = a[2]; }
  return done;
}

/* Batched: arrays are [n][3], R is [n][9], actions [n][4]; wind is [3] (uniform).  `substeps` reference steps
 * per env with the action held; done[n] = OR over substeps.  Single-threaded; callers shard [n]. */
void fpv_oracle_drone_step(const oracle_consts_t* c, int64_t n, double* pos, double* vel, double* R, double* prev_rates,
                           double* prev_thrust, uint8_t* done, const double* actions, const double* wind, int substeps,
                           double* acc_out) {
  for (int64_t e = 0; e < n; ++e) {
    int d = 0;
    for (int s = 0; s < substeps; ++s)
      d |= drone_step_one(c, pos + 3 * e, vel + 3 * e, R + 9 * e, prev_rates + 3 * e, prev_thrust + e, actions + 4 * e,
                          wind, acc_out ? acc_out + 3 * e : 0);
    done[e] = (uint8_t)d;
  }
}

int fpv_oracle_sizeof_consts(void) { return (int)sizeof(oracle_consts_t); }

