"""TEST INFRASTRUCTURE ONLY.  Generates `tests/golden/*.npz` by executing the UNMODIFIED
reference (through `oracle/ref_shim.py`) on seeded inputs.  Run in the build container, where
/root/reference is mounted:

    python oracle/make_golden.py

The vectors are committed because the GPU box has no /root/reference.  Every file stores the
inputs next to the reference's float64 outputs, so both the CPU restatement
(`tests/test_oracle_golden.py`) and the CUDA path (`tests/test_gpu_parity.py`) replay them.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim as rs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _stack(rollouts):
    """list over envs of dict[T,...] -> dict[T, n, ...]"""
    return {k: np.stack([r[k] for r in rollouts], axis=1) for k in rollouts[0]}


def drone_case(name, n, T, seed, dt=None, z_lo=1.0, z_hi=12.0, wind=None, act_scale=1.0, hold=1):
    rng = np.random.default_rng(seed)
    pos = np.stack([rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(z_lo, z_hi, n)], axis=1)
    vel = rng.normal(0, 1, (n, 3))
    rpy = rng.uniform(-30, 30, (n, 3))
    nh = (T + hold - 1) // hold
    actions = np.repeat(rng.uniform(-1, 1, (nh, n, 4)) * act_scale, hold, axis=0)[:T]
    outs = []
    for e in range(n):
        outs.append(rs.ref_drone_rollout(actions[:, e], pos[e], vel[e], rpy[e], wind=wind, dt=dt))
    o = _stack(outs)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), pos0=pos, vel0=vel, rpy0=rpy, actions=actions,
                        wind=np.zeros(3) if wind is None else np.asarray(wind, dtype=np.float64),
                        dt=np.float64(1 / 60 if dt is None else dt), hold=np.int64(hold),
                        state=o["state"], R=o["R"], prev_rates=o["prev_rates"], prev_thrust=o["prev_thrust"],
                        done=o["done"], ret_Rt=o["ret_Rt"], ret_gyro=o["ret_gyro"], ret_acc=o["ret_acc"],
                        acc=o["acc"])
    print(name, "done-steps:", int(o["done"].sum()), "of", o["done"].size)


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = rs.load()
    # --- derived constants (Drone.__init__, components.py:84-142)
    d = rs.make_drone()
    thr = d.motor_test_report['Throttle'].values
    thrust_n = d.n_motors * d.motor_test_report['Thrust'].values / 1000 * d.gravity
    xs = np.linspace(-1, 1, 41)
    fs = np.linspace(0, 90, 31)
    np.savez_compressed(
        os.path.join(OUT, "consts.npz"), dt=d.dt, mass=d.mass, gravity=d.gravity, max_rates=d.max_rates,
        rtr=d.rates_transition_rate, ttr=d.thrust_transition_rate, drag_coef=d.drag_coef,
        cross_section_areas=d.cross_section_areas, motors_relative_position=d.motors_relative_position,
        poly=ns["ftc"].model_xy(thr, thrust_n).coeffs, inv_poly=ns["ftc"].model_xy(thrust_n, thr).coeffs,
        min_force=d.min_throttle_in_force, max_force=d.max_throttle_in_force, throttle_pct=thr, thrust_n=thrust_n,
        t2t_x=xs, t2t_y=np.array([d.throttle2thrust(x) for x in xs]),
        inv_x=fs, inv_y=np.array([d.thrust2throttle(f) for f in fs]))

    # --- Drone.step trajectories (components.py:220-248)
    # the SURVEY KAT: stock params.yaml initial state, constant action, 60 steps at 1/60 s
    o = rs.ref_drone_rollout(np.tile([0.3, -0.2, 0.1, 0.25], (60, 1)), [0, 0, 10], [1, 0, 0], [0, 0, 0])
    np.savez_compressed(os.path.join(OUT, "drone_kat.npz"), pos0=np.array([[0., 0, 10]]), vel0=np.array([[1., 0, 0]]),
                        rpy0=np.zeros((1, 3)), actions=np.tile([0.3, -0.2, 0.1, 0.25], (60, 1, 1)),
                        wind=np.zeros(3), dt=np.float64(1 / 60), hold=np.int64(1),
                        **{k: v[:, None] for k, v in o.items() if k != "action"})
    drone_case("drone_random", n=16, T=120, seed=1)
    drone_case("drone_ground", n=24, T=40, seed=2, z_lo=0.02, z_hi=0.6)
    drone_case("drone_wind", n=8, T=60, seed=3, wind=[2.0, -1.0, 0.5])
    drone_case("drone_1ms_k8", n=8, T=1000, seed=4, dt=1e-3, hold=8)
    drone_case("drone_overdrive", n=8, T=30, seed=5, act_scale=1.7)     # exercises the rate clip

    # --- step(rotation_matrix=, thrust_force=) override inputs (components.py:230-232)
    rng = np.random.default_rng(6)
    n, T = 6, 20
    pos = np.stack([rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(2, 12, n)], axis=1)
    vel = rng.normal(0, 1, (n, 3))
    rpy = rng.uniform(-30, 30, (n, 3))
    actions = rng.uniform(-1, 1, (T, n, 4))
    eul = np.deg2rad(rng.uniform(-40, 40, (T, n, 3)))
    force = rng.uniform(2, 40, (T, n))
    Ro = np.zeros((T, n, 3, 3))
    outs = {k: np.zeros((T, n) + s) for k, s in (("state", (6,)), ("R", (3, 3)), ("prev_rates", (3,)),
                                                  ("prev_thrust", ()), ("done", ()))}
    ground = rs.make_ground()
    for e in range(n):
        dr = rs.make_drone()
        with rs.quiet():
            dr.reset(pos[e], vel[e], rpy[e])
        for t in range(T):
            Ro[t, e] = ns["helper_functions"].euler_angles_to_rotation_matrix(*eul[t, e])
            with rs.quiet():
                dr.step(actions[t, e], np.zeros(3), [ground], rotation_matrix=Ro[t, e].copy(), thrust_force=force[t, e])
            outs["state"][t, e] = dr.state
            outs["R"][t, e] = dr.rotation_matrix
            outs["prev_rates"][t, e] = dr.prev_rates
            outs["prev_thrust"][t, e] = dr.prev_thrust
            outs["done"][t, e] = dr.done
    np.savez_compressed(os.path.join(OUT, "drone_override.npz"), pos0=pos, vel0=vel, rpy0=rpy, actions=actions,
                        R_override=Ro, thrust_override=force, dt=np.float64(1 / 60), **outs)

    # --- obstacles: ground + cylinder + sphere (components.py:198-214, :710-729, :773-777)
    rng = np.random.default_rng(7)
    comp = ns["components"]
    cyl_p, cyl_r, cyl_h = np.array([3.0, 0.0, 0.0]), 1.0, 6.0
    sph_p, sph_r = np.array([-3.0, 1.0, 4.0]), 1.2
    n, T = 12, 50
    pos = np.stack([rng.uniform(-6, 6, n), rng.uniform(-2, 2, n), rng.uniform(2, 7, n)], axis=1)
    tgt = np.where(rng.random(n)[:, None] < 0.5, cyl_p + [0, 0, 3.0], sph_p)
    vel = (tgt - pos) * rng.uniform(0.8, 1.6, (n, 1))
    rpy = rng.uniform(-20, 20, (n, 3))
    actions = rng.uniform(-0.3, 0.3, (T, n, 4))
    outs = []
    for e in range(n):
        with rs.quiet():
            cyl = comp.Cylinder(cyl_p.copy(), cyl_r, cyl_h, 4, 4)
            sph = comp.Target(sph_p.copy(), sph_r, nu=1)
        outs.append(rs.ref_drone_rollout(actions[:, e], pos[e], vel[e], rpy[e], objects=[sph, cyl, ground]))
    o = _stack(outs)
    np.savez_compressed(os.path.join(OUT, "drone_objects.npz"), pos0=pos, vel0=vel, rpy0=rpy, actions=actions,
                        dt=np.float64(1 / 60), cyl=np.array([*cyl_p, cyl_r, cyl_h]), sph=np.array([*sph_p, sph_r]),
                        state=o["state"], R=o["R"], prev_rates=o["prev_rates"], prev_thrust=o["prev_thrust"],
                        done=o["done"])
    print("drone_objects done-steps:", int(o["done"].sum()), "of", o["done"].size)

    # --- stick calibration (get_sticks.py:245-265, components.py:250-253)
    rng = np.random.default_rng(8)
    for calib in ("frsky.json", "calibration.json"):
        raw = rng.integers(0, 65536, (200, 6)).astype(np.float64)
        raw[0] = [30000, 20000, 50000, 0, 0, 40000]           # SURVEY KAT
        import json
        cal = json.load(open(os.path.join(rs.REF_ROOT, "config", calib)))
        raw[1] = cal["min_vals"]
        raw[2] = cal["max_vals"]
        cal_out = np.array([rs.ref_calib_read(r, calib) for r in raw])
        action = np.stack([-cal_out[:, 1], cal_out[:, 2], cal_out[:, 5], cal_out[:, 0]], axis=1)
        np.savez_compressed(os.path.join(OUT, "sticks_" + calib.split(".")[0] + ".npz"), raw=raw,
                            calibrated=cal_out, action=action)
    # joystick-driven Drone.step(action=None) end to end
    raw = rng.integers(3000, 60000, (30, 6)).astype(np.float64)
    o = rs.ref_drone_rollout(None, [0, 0, 10], [1, 0, 0], [0, 0, 0], raw_axes=raw)
    np.savez_compressed(os.path.join(OUT, "drone_sticks.npz"), raw=raw[:, None], pos0=np.array([[0., 0, 10]]),
                        vel0=np.array([[1., 0, 0]]), rpy0=np.zeros((1, 3)), dt=np.float64(1 / 60),
                        actions=o["action"][:, None], state=o["state"][:, None], R=o["R"][:, None],
                        prev_rates=o["prev_rates"][:, None], prev_thrust=o["prev_thrust"][:, None],
                        done=o["done"][:, None])

    # --- Racer (tests/racer_drone_test.py): the script's own scenario + seeded variants
    gains = {"roll": [2, 0, 0], "pitch": [2, 0, 0], "yaw": [0.1, 0, 0]}
    acts = np.array([[80, 10, 0, 0]] * 22 + [[-30, -50, 0, 0]] * 978, dtype=np.float64)   # :113-122 (switch after t==20)
    o = rs.ref_racer_rollout(acts, gains)
    np.savez_compressed(os.path.join(OUT, "racer_demo.npz"), actions=acts[:, None], gains=np.array([list(gains.values())]),
                        **{k: v[:, None] for k, v in o.items()})
    rng = np.random.default_rng(9)
    n, T = 8, 400
    gains_n = np.stack([rng.uniform(0.5, 3, (n, 3)), rng.uniform(0, 0.5, (n, 3)), rng.uniform(0, 2e-4, (n, 3))], axis=2)
    gains_n[0, :, 1:] = 0
    acts = np.repeat(np.concatenate([rng.uniform(-6, 6, (T // 20, n, 3)), rng.uniform(0, 12, (T // 20, n, 1))], axis=2), 20, axis=0)
    outs = [rs.ref_racer_rollout(acts[:, e], dict(zip(("roll", "pitch", "yaw"), gains_n[e]))) for e in range(n)]
    o = _stack(outs)
    np.savez_compressed(os.path.join(OUT, "racer_random.npz"), actions=acts, gains=gains_n, **o)

    # --- L1 free-function KATs (kinematics.py:33-38, :56-59; helper_functions.py:39-44)
    hf, kin = ns["helper_functions"], ns["kinematics"]
    R = hf.euler_angles_to_rotation_matrix(0.1, 0.2, 0.3)
    with rs.quiet():
        drag = kin.calculate_drag(R, np.array([3., -2, 1]), np.array([0.5, 0, 0]), d.drag_coef, d.cross_section_areas)
    np.savez_compressed(os.path.join(OUT, "l1_kat.npz"), R_euler_0p1_0p2_0p3=R, drag=drag,
                        spring=kin.spring_force(-0.05, np.array([0, 0, 1.]), np.array([1., 2, 3]), 100, 0),
                        rot_by_rates=kin.rotate_body_by_rates(R, np.array([-42., 28, -14]), 1 / 60),
                        quat=hf.rotation_matrix_to_quaternion(R), quat_R=hf.quaternion_to_rotation_matrix(hf.rotation_matrix_to_quaternion(R)))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
