"""TEST / BASELINE INFRASTRUCTURE ONLY -- Numba `@njit(parallel=True)` float64 restatement of the reference's
`Drone.step` (ground-only object list), batched over envs with `prange`.

Why it exists: BASELINE.json's north_star asks for "the reference's Numba CPU path" to be timed next to the GPU
number.  THE REFERENCE SHIPS NO NUMBA PATH: its only Numba lines are commented out (src/utils/kinematics.py:6, :14)
and nothing is jitted.  This file is therefore OUR restatement of the reference step in the form a Numba port of the
reference would take; bench.py reports it as `cpu_baseline_numba` with kind "restatement".  Only `tests/` and the
CPU legs of `bench.py` may import it; the product never does.

Parity status: PINNED -- tests/test_oracle_golden.py checks it against the golden vectors that oracle/make_golden.py
produced by executing the unmodified reference.

Reference lines followed (relative to /root/reference): action2force src/utils/components.py:179-196; calculate_drag
src/utils/kinematics.py:33-38; gravity :41-45; motors / collisions components.py:235-239, :198-214, Ground :674-680,
spring_force kinematics.py:56-59; force sum components.py:242-243; update components.py:216-218 +
kinematics.py:15-30 (the attitude increment is applied TWICE); Euler matrix src/utils/helper_functions.py:19-44."""
from __future__ import annotations

import math

import numba
import numpy as np
from numba import njit, prange


@njit(cache=True, inline="always")
def _euler(roll, pitch, yaw, E):
    sr, cr = math.sin(roll), math.cos(roll)
    sp, cp = math.sin(pitch), math.cos(pitch)
    sy, cy = math.sin(yaw), math.cos(yaw)
    E[0, 0] = cy * cp; E[0, 1] = cy * sp * sr - sy * cr; E[0, 2] = cy * sp * cr + sy * sr
    E[1, 0] = sy * cp; E[1, 1] = sy * sp * sr + cy * cr; E[1, 2] = sy * sp * cr - cy * sr
    E[2, 0] = -sp;     E[2, 1] = cp * sr;                E[2, 2] = cp * cr


@njit(parallel=True, cache=True, fastmath=False)
def drone_step(consts, motor_rel, pos, vel, R, prev_rates, prev_thrust, actions, wind, substeps, done):
    """consts = [dt, gravity, mass, max_rates, rtr, ttr, kd0, kd1, kd2, p0, p1, p2, p3, motor_radius, spring_k, spring_c, ground].
    In place on pos[n,3] vel[n,3] R[n,3,3] prev_rates[n,3] prev_thrust[n]; done[n] = OR of the substeps' crash flags."""
    dt, g, mass, max_rates, rtr, ttr = consts[0], consts[1], consts[2], consts[3], consts[4], consts[5]
    motor_radius, spring_k, spring_c, ground = consts[13], consts[14], consts[15], consts[16] != 0.0
    d2r = math.pi / 180.0
    n = pos.shape[0]
    for e in prange(n):
        E = np.empty((3, 3))
        T = np.empty((3, 3))
        Re = R[e]
        dn = False
        for _ in range(substeps):
            # action2force, components.py:185-194
            rates0 = min(max(-actions[e, 0] * max_rates, -max_rates), max_rates) * rtr + prev_rates[e, 0] * (1 - rtr)
            rates1 = min(max(-actions[e, 1] * max_rates, -max_rates), max_rates) * rtr + prev_rates[e, 1] * (1 - rtr)
            rates2 = min(max(-actions[e, 2] * max_rates, -max_rates), max_rates) * rtr + prev_rates[e, 2] * (1 - rtr)
            prev_rates[e, 0] = rates0; prev_rates[e, 1] = rates1; prev_rates[e, 2] = rates2
            pct = 100.0 * (actions[e, 3] + 1.0) / 2.0
            thr = (((consts[9] * pct + consts[10]) * pct + consts[11]) * pct + consts[12]) * ttr + prev_thrust[e] * (1 - ttr)
            prev_thrust[e] = thr
            # drag, kinematics.py:33-38 (velocity PLUS wind)
            ux, uy, uz = vel[e, 0] + wind[0], vel[e, 1] + wind[1], vel[e, 2] + wind[2]
            nrm = math.sqrt(ux * ux + uy * uy + uz * uz)
            fx = fy = 0.0
            fz = -g * mass
            for j in range(3):
                vb = Re[0, j] * ux + Re[1, j] * uy + Re[2, j] * uz
                fb = consts[6 + j] * vb * nrm
                fx += Re[0, j] * fb; fy += Re[1, j] * fb; fz += Re[2, j] * fb
            fx += Re[0, 2] * thr; fy += Re[1, 2] * thr; fz += Re[2, 2] * thr
            # motors, ground collision, crash test (components.py:235-239, :198-214)
            crashed = False
            cz = 0.0
            any_below = False
            for m in range(4):
                mz = pos[e, 2] + Re[2, 0] * motor_rel[m, 0] + Re[2, 1] * motor_rel[m, 1] + Re[2, 2] * motor_rel[m, 2]
                if mz < 0.0:
                    any_below = True
                pen = mz - motor_radius
                if pen < 0.0:
                    cz += -spring_k * pen - spring_c * vel[e, 2]
            if ground and any_below:
                crashed = True
                cz = 0.0
            if not ground:
                cz = 0.0
            dn = dn or crashed or any_below
            fz += cz
            # update, kinematics.py:21-22 then the doubled attitude increment
            ax, ay, az = fx / mass, fy / mass, fz / mass
            pos[e, 0] += vel[e, 0] * dt; pos[e, 1] += vel[e, 1] * dt; pos[e, 2] += vel[e, 2] * dt
            vel[e, 0] += ax * dt; vel[e, 1] += ay * dt; vel[e, 2] += az * dt
            _euler(rates0 * d2r * dt, rates1 * d2r * dt, rates2 * d2r * dt, E)
            for _rep in range(2):
                for i in range(3):
                    for j in range(3):
                        T[i, j] = Re[i, 0] * E[j, 0] + Re[i, 1] * E[j, 1] + Re[i, 2] * E[j, 2]
                for i in range(3):
                    for j in range(3):
                        Re[i, j] = T[i, j]
        done[e] = dn


def make_consts(c):
    """c: oracle.fpv_oracle.DroneConsts -> (consts[17], motor_rel[4,3])"""
    k = np.array([c.dt, c.gravity, c.mass, c.max_rates, c.rtr, c.ttr, *c.k_drag, *c.poly, c.motor_radius, c.spring_k,
                  c.spring_c, 1.0 if c.ground else 0.0], dtype=np.float64)
    return k, np.ascontiguousarray(c.motor_rel, dtype=np.float64)


def threads():
    return numba.get_num_threads()
