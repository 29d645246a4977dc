"""The modes that moved onto the persistent TMA-ring kernel in round 2 (general / obstacle path with the per-control-step
reach test, gate-race epilogue, packed quaternion `Racer`, acro) and the features added to the ring (done bitmask, the
error word that replaced the device trap of a broken chaining contract)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from oracle import fpv_oracle as fo

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_parity import TOL_STEP, drone_err, fo_consts, make  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("K,dt", [(8, 2e-3), (4, 1 / 240), (16, 1e-3)])
def test_general_path_reach_test_hoisted_out_of_the_substep_loop(packed, K, dt):
    """The general kernel decides ONCE per control step which obstacles any motor can reach during the K substeps (bound on
    the speed over the step, drone_kernels.cuh) and skips the others in every substep.  Drones are placed 0.05 .. 3 m outside
    the reach of a sphere and a cylinder and thrown AT them at up to 60 m/s (plus bystanders far away and drones already in
    contact), so that many enter the contact shell or crash in the middle of the control step: state and crash flags must
    match the float64 oracle, which tests every obstacle in every substep."""
    from fpyv_b200 import Cylinder, Ground, Target
    n = 24_000
    rng = np.random.default_rng(91 + K)
    sph_c, sph_r = np.array([2.0, -1.0, 6.0]), 1.3
    cyl_p, cyl_r, cyl_h = np.array([-4.0, 3.0, 0.0]), 0.8, 7.0
    reach = 0.127 + 0.1
    gap = rng.choice([-0.02, 0.05, 0.2, 0.5, 1.0, 2.0, 3.0, 40.0], n) + reach + rng.normal(0, 0.01, n)
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    pos = sph_c + u * (sph_r + gap)[:, None]
    speed = rng.choice([0.0, 5.0, 20.0, 40.0, 60.0], n)
    vel = -u * speed[:, None] + rng.normal(0, 0.5, (n, 3))
    half = n // 2                                                    # second half: around the cylinder's side
    ang = rng.uniform(0, 2 * np.pi, half)
    rad = cyl_r + gap[half:]
    pos[half:] = np.stack([cyl_p[0] + rad * np.cos(ang), cyl_p[1] + rad * np.sin(ang), rng.uniform(1.0, cyl_h - 1.0, half)], axis=1)
    vel[half:] = np.stack([-np.cos(ang), -np.sin(ang), np.zeros(half)], 1) * speed[half:, None] + rng.normal(0, 0.5, (half, 3))
    pos[:, 2] = np.maximum(pos[:, 2], 0.5)
    rpy = rng.uniform(-40, 40, (n, 3))
    act = rng.uniform(-1, 1, (n, 4)).astype(np.float32).astype(np.float64)       # identical inputs on both sides
    c = fo_consts(dt)
    s = fo.drone_reset(c, pos, vel, rpy)
    d = make(n, packed=packed, substeps=K, dt=dt)
    d.reset(pos, vel, rpy)
    s.pos, s.vel = d.position.double().cpu().numpy(), d.velocity.double().cpu().numpy()
    s.R = fo.quaternion_to_matrix(d.quaternion.double().cpu().numpy())
    objs = [Target(sph_c, sph_r), Cylinder(cyl_p, cyl_r, cyl_h), Ground()]
    extra = [fo.SphereObj(sph_c, sph_r), fo.CylinderObj(cyl_p, cyl_r, cyl_h)]
    touched = np.zeros(n, dtype=bool)
    for _ in range(2):                                               # two control steps: the second starts mid-contact
        a0 = np.abs(s.acc).max(1)
        fo.drone_step(c, s, act, substeps=K, extra_objects=extra)
        d.step(act, np.zeros(3), objs, return_obs=False)
        touched |= np.abs(s.acc).max(1) > 150
        done = d.done.cpu().numpy()
        mism = done != s.done
        err = drone_err(d, np.concatenate([s.pos, s.vel], 1), s.R, s.prev_rates, s.prev_thrust, vel_floor=1.0)
        ok = ~mism & ~s.done                    # a crashed env keeps integrating; compare the survivors tightly
        print(f"\nhoisted reach K={K} dt={dt:.4f} packed={packed}: max rel err {err[ok].max():.2e}, crashed {int(s.done.sum())}, "
              f"flag mismatches {int(mism.sum())}")
        assert mism.sum() <= 6, mism.sum()       # a motor within fp32 rounding of a surface may flip the flag
        assert err[ok].max() <= TOL_STEP         # K free-running substeps, many through contact (measured <= 1.4e-6)
        # re-synchronise the oracle on the device state so that the second step is a single control step again
        s.pos, s.vel = d.position.double().cpu().numpy(), d.velocity.double().cpu().numpy()
        s.R = fo.quaternion_to_matrix(d.quaternion.double().cpu().numpy())
        s.prev_rates, s.prev_thrust = d.prev_rates.double().cpu().numpy(), d.prev_thrust.double().cpu().numpy()
    assert s.done.sum() > 200


@pytest.mark.parametrize("n", [1, 63, 1000, 65_537, 1 << 18])
@pytest.mark.parametrize("mode", ["hot", "general", "scalar", "freeze"])
def test_done_bitmask_equals_the_done_bytes(n, mode):
    """fpv_drone_io_t.done_bits: one ballot per 32 envs.  Bit e % 32 of word e // 32 == done[e], every kernel form, ragged
    sizes, frozen envs (sticky) included; words past the batch are never written."""
    from fpyv_b200 import BatchedDrone, Cylinder, Ground
    rng = np.random.default_rng(n)
    pos = np.stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.uniform(0.05, 0.6, n)], 1)
    vel, rpy = rng.normal(size=(n, 3)), rng.uniform(-50, 50, (n, 3))
    kw = dict(packed=mode != "scalar", freeze_done=mode == "freeze", auto_reset=mode == "hot")
    d = BatchedDrone(None, num_envs=n, device=DEV, substeps=4, dt=2e-3, done_bits=True, **kw)
    d.reset(pos, vel, rpy)
    objs = [Cylinder(np.array([50.0, 50.0, 0.0]), 1.0, 5.0), Ground()] if mode == "general" else None
    seen = 0
    for t in range(6):
        a = rng.uniform(-1, 1, (n, 4))
        a[:, 3] = rng.uniform(-1, -0.5, n)           # low throttle: the drones sink and crash
        d._done_bits.fill_(0x5A5A5A5A)
        d.step(a, None, objs, return_obs=False)
        by = d._done.cpu().numpy().astype(bool)
        words = d.done_bits.cpu().numpy().view(np.uint32)
        bits = ((words[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(bool).reshape(-1)[:n]
        assert np.array_equal(bits, by), t
        if n % 32:
            assert (words[-1] >> np.uint32(n % 32)) == 0        # bits past the batch inside the last word are clear
        seen += int(by.sum())
    assert seen > 0 or n < 8


def test_broken_chaining_contract_raises_the_error_word_not_a_trap():
    """A chained launch whose epoch never arrives (here: the host's epoch counter is pushed ahead, as a replayed captured
    launch would) used to __trap() the context after a spin count.  Now: the launch waits 2 s of wall-clock time, adds to the
    error word, falls back to the grid-wide wait and completes -- with the right result; the context stays usable."""
    from fpyv_b200 import BatchedDrone
    n, K = 200_000, 2
    rng = np.random.default_rng(5)
    pos = np.stack([rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(1, 3, n)], 1)
    vel, rpy = rng.normal(size=(n, 3)), rng.uniform(-30, 30, (n, 3))
    acts = [torch.as_tensor(rng.uniform(-1, 1, (n, 4)), dtype=torch.float32, device=DEV) for _ in range(3)]
    a, b = (BatchedDrone(None, num_envs=n, device=DEV, substeps=K, dt=1e-3) for _ in range(2))
    a.reset(pos, vel, rpy)
    b.reset(pos, vel, rpy)
    for t in range(2):
        a.step(acts[t], return_obs=False, chained=True)
        b.step(acts[t], return_obs=False)
    assert a.episode_stats()["chain_timeouts"] == 0
    a._epoch += 7                                   # break the contract: the kernel will wait for an epoch nobody publishes
    a.step(acts[2], return_obs=False, chained=True)
    b.step(acts[2], return_obs=False)
    torch.cuda.synchronize()                        # (about 2 s)
    assert a.episode_stats()["chain_timeouts"] > 0
    assert torch.equal(a._state, b._state)
    a.step(acts[0], return_obs=False, chained=True)  # the epochs are consistent again: no further timeouts
    b.step(acts[0], return_obs=False)
    t0 = a.episode_stats()["chain_timeouts"]
    a.step(acts[1], return_obs=False, chained=True)
    b.step(acts[1], return_obs=False)
    torch.cuda.synchronize()
    assert a.episode_stats()["chain_timeouts"] == t0 and torch.equal(a._state, b._state)


def test_racer_packed_matches_scalar_and_the_oracle_at_large_rates():
    """Mode B on the ring: the packed (two envs per thread) and scalar instantiations agree bit for bit (the same FP32
    operations per env), ragged sizes leave the padding alone, and set-points that drive |omega| to ~100 rad (half angles far
    outside the polynomial range: Cody-Waite reduction) still match the float64 oracle."""
    from fpyv_b200 import BatchedRacer, _lib
    n = 3001
    gains = {"roll": [2, 0.1, 1e-4], "pitch": [1.5, 0.2, 0], "yaw": [0.1, 0, 0]}
    rng = np.random.default_rng(3)
    sp = np.concatenate([rng.uniform(-100, 100, (n, 3)), rng.uniform(0, 10, (n, 1))], 1)
    sp[: n // 3, :3] = rng.uniform(-3, 3, (n // 3, 3))
    rp = BatchedRacer(5, gains, num_envs=n, device=DEV, dt=1e-3, substeps=5)
    rs = BatchedRacer(5, gains, num_envs=n, device=DEV, dt=1e-3, substeps=5)
    rs._p.flags = _lib.F_SCALAR
    for r in (rp, rs):
        r.reset()
        r._state[:, n:] = 77.0
    for t in range(4):
        rp.step(sp)
        rs.step(sp)
    torch.cuda.synchronize()
    assert torch.equal(rp._state[:, :n], rs._state[:, :n])
    assert bool((rp._state[:, n:] == 77.0).all())
    rc = fo.RacerConsts(gains=np.array(list(gains.values()), dtype=np.float64))
    s = fo.RacerState(n)
    for t in range(20):
        fo.racer_step(rc, s, sp)
    R, w, x = (v.double().cpu().numpy() for v in (rp.orientation, rp.angular_velocity, rp.position))
    worst = max(np.abs(R - s.R).max(), (np.abs(w - s.omega).max(1) / np.maximum(1.0, np.abs(s.omega).max(1))).max(),
                (np.abs(x - s.pos).max(1) / np.maximum(1e-2, np.abs(s.pos).max(1))).max())
    print(f"\nracer packed vs oracle after 20 steps, |omega| up to {np.abs(rp.angular_velocity.cpu().numpy()).max():.0f} rad/s: {worst:.2e}")
    assert worst <= 2e-4


def test_step_host_returns_the_done_bitmask():
    """fpv_drone_step_host with io.done_bits: the D2H copy carries ceil(n/32) words instead of n bytes; same flags, same state
    as the plain step."""
    from fpyv_b200 import BatchedDrone
    n = 300_037
    rng = np.random.default_rng(8)
    pos = np.stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.uniform(0.05, 1.0, n)], 1)
    vel, rpy = rng.normal(size=(n, 3)), rng.uniform(-40, 40, (n, 3))
    a = BatchedDrone(None, num_envs=n, device=DEV, substeps=4, dt=1e-3, auto_reset=True, thrust_lut=513, done_bits=True)
    b = BatchedDrone(None, num_envs=n, device=DEV, substeps=4, dt=1e-3, auto_reset=True, thrust_lut=513)
    a.reset(pos, vel, rpy)
    b.reset(pos, vel, rpy)
    host_a = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    bits_host = torch.zeros((n + 31) // 32, dtype=torch.int32, pin_memory=True)
    with pytest.raises(ValueError):
        a.step_host(host_a, torch.empty(n, dtype=torch.uint8, pin_memory=True))
    crashed = 0
    for t in range(6):
        host_a.copy_(torch.as_tensor(rng.uniform(-1, 1, (n, 4)), dtype=torch.float32))
        # even steps: flags copied back by the library; odd steps: flag words written to the host buffer by the kernels
        a.step_host(host_a, bits_host, slices=4 if t % 2 == 0 else 3, flags_direct=t % 2 == 1)
        b.step(host_a.to(DEV), return_obs=False)
        torch.cuda.synchronize()
        words = bits_host.numpy().view(np.uint32)
        bits = ((words[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(bool).reshape(-1)[:n]
        assert np.array_equal(bits, b._done.cpu().numpy().astype(bool)), t
        crashed += int(bits.sum())
    assert torch.equal(a._state, b._state) and crashed > 0


@pytest.mark.parametrize("K", [1, 8])
def test_general_path_without_contact_gives_the_hot_path_bits(K):
    """An env that touches no obstacle gets, from the general kernel, exactly the bits the hot kernel gives it -- whether its
    warp ran the hot loop (no env of the warp in reach of anything) or the general loop (a neighbour in contact).  One batch:
    the first half far from the obstacles, the second half mixed into contact (so warps of both kinds exist, and warps that
    contain both kinds of env)."""
    from fpyv_b200 import BatchedDrone, Cylinder, Ground, Target
    n = 40_000
    rng = np.random.default_rng(13)
    pos = np.stack([rng.normal(60, 5, n), rng.normal(60, 5, n), rng.uniform(0.05, 3.0, n)], 1)      # far away: ground only
    near = rng.random(n) < 0.3
    near[: n // 2] = False
    pos[near] = np.array([0.0, 0.0, 3.0]) + rng.normal(0, 1.2, (int(near.sum()), 3))                 # around the sphere
    pos[:, 2] = np.maximum(pos[:, 2], 0.05)
    vel, rpy = rng.normal(0, 2, (n, 3)), rng.uniform(-30, 30, (n, 3))
    objs = [Target(np.array([0.0, 0.0, 3.0]), 1.0), Cylinder(np.array([5.0, -4.0, 0.0]), 1.0, 6.0), Ground()]
    a = BatchedDrone(None, num_envs=n, device=DEV, substeps=K, dt=2e-3, thrust_lut=1025)
    b = BatchedDrone(None, num_envs=n, device=DEV, substeps=K, dt=2e-3, thrust_lut=1025)
    a.reset(pos, vel, rpy)
    b.reset(pos, vel, rpy)
    far = torch.as_tensor(~near, device=DEV)
    touched = 0
    for t in range(5):
        act = torch.as_tensor(rng.uniform(-1, 1, (n, 4)), dtype=torch.float32, device=DEV)
        a.step(act, None, objs, return_obs=False)       # general kernel
        b.step(act, return_obs=False)                   # hot kernel (no object list: ground only)
        assert torch.equal(a._state[:, :n][:, far], b._state[:, :n][:, far]), t
        assert torch.equal(a._done[far], b._done[far]) and torch.equal(a._acc[far], b._acc[far])
        touched += int((a._done != b._done).sum())
        # keep the two batches on the same trajectory for the next comparison
        b._state.copy_(a._state)
    assert touched > 50          # the near envs really did hit the obstacles


def _bits_to_bool(words, n):
    w = words.numpy().view(np.uint32)
    return ((w[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(bool).reshape(-1)[:n]


@pytest.mark.parametrize("write_combined", [False, True])
def test_zero_copy_host_step_matches_plain_step(write_combined):
    """step_host(zero_copy=True): ONE launch whose TMA engine reads the actions straight from page-locked HOST memory and
    whose warps write the done bitmask straight back to it -- same state and flags as copy + step, ragged batch size,
    ordinary pinned and write-combined input buffers."""
    from fpyv_b200 import BatchedDrone, hostmem
    n = 300_037
    rng = np.random.default_rng(21)
    pos = np.stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.uniform(0.05, 1.0, n)], 1)
    vel, rpy = rng.normal(size=(n, 3)), rng.uniform(-40, 40, (n, 3))
    a = BatchedDrone(None, num_envs=n, device=DEV, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049, done_bits=True)
    b = BatchedDrone(None, num_envs=n, device=DEV, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
    a.reset(pos, vel, rpy)
    b.reset(pos, vel, rpy)
    host_a = hostmem.pinned((n, 4), torch.float32, write_combined=write_combined)
    bits_host = hostmem.pinned(((n + 31) // 32,), torch.int32)
    with pytest.raises(ValueError):
        a.step_host(torch.zeros(n, 4), bits_host, zero_copy=True)        # pageable memory is refused
    crashed = 0
    for t in range(5):
        act = torch.as_tensor(rng.uniform(-1, 1, (n, 4)), dtype=torch.float32)
        host_a.copy_(act)
        a.step_host(host_a, bits_host, zero_copy=True)
        b.step(act.to(DEV), return_obs=False)
        torch.cuda.synchronize()
        assert np.array_equal(_bits_to_bool(bits_host, n), b._done.cpu().numpy().astype(bool)), t
        crashed += int(b._done.sum())
    assert torch.equal(a._state, b._state) and crashed > 0
    a.step(act.to(DEV), return_obs=False)                               # the drone's own buffers are back in place
    b.step(act.to(DEV), return_obs=False)
    assert torch.equal(a._state, b._state) and np.array_equal(_bits_to_bool(a.done_bits.cpu(), n), b._done.cpu().numpy().astype(bool))


@pytest.mark.parametrize("fmt", ["u16", "crsf"])
@pytest.mark.parametrize("zero_copy", [False, True])
def test_host_stick_formats_match_the_joystick_path(fmt, zero_copy):
    """step_host_sticks in both transport formats (uint16 x 4 = 8 B/env, CRSF 4 x 11 bit = 6 B/env), as the sliced copy
    pipeline and as ONE zero-copy launch that calibrates the sticks in registers: bit-identical to the reference's joystick
    path `rc.feed(raw axes); step(action=None)` (components.py:227-228, :250-253)."""
    from fpyv_b200 import BatchedDrone, hostmem
    from fpyv_b200.sticks import crsf_to_raw16, pack_crsf
    n = 200_011
    rng = np.random.default_rng(31)
    pos = np.stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.uniform(0.05, 1.5, n)], 1)
    vel, rpy = rng.normal(size=(n, 3)), rng.uniform(-40, 40, (n, 3))
    a = BatchedDrone(None, num_envs=n, device=DEV, substeps=4, dt=1e-3, auto_reset=True, thrust_lut=2049, done_bits=True)
    b = BatchedDrone(None, num_envs=n, device=DEV, substeps=4, dt=1e-3, auto_reset=True, thrust_lut=2049)
    a.reset(pos, vel, rpy)
    b.reset(pos, vel, rpy)
    sticks = hostmem.pinned((n, 4), torch.uint16) if fmt == "u16" else hostmem.pinned((n, 6), torch.uint8)
    bits_host = hostmem.pinned(((n + 31) // 32,), torch.int32)
    for t in range(4):
        if fmt == "u16":
            raw4 = rng.integers(0, 65536, (n, 4))
            sticks.copy_(torch.from_numpy(raw4.astype(np.uint16)))
        else:
            v11 = rng.integers(0, 2048, (n, 4))
            sticks.copy_(pack_crsf(v11))
            raw4 = crsf_to_raw16(v11)
        a.step_host_sticks(sticks, bits_host, zero_copy=zero_copy)
        raw6 = np.zeros((n, 6), dtype=np.int32)
        raw6[:, [0, 1, 2, 5]] = raw4
        b.rc.feed(raw6)
        b.step(None, return_obs=False)
        torch.cuda.synchronize()
        assert np.array_equal(_bits_to_bool(bits_host, n), b._done.cpu().numpy().astype(bool)), t
    assert torch.equal(a._state, b._state)


def test_static_world_steps_on_the_fast_path_and_chains():
    """set_static_objects(): the object list is lowered once, `step(action)` then runs the general kernel from the
    allocation-free fast path -- plain and chained -- with the bits of passing the same list to every call."""
    from fpyv_b200 import BatchedDrone, Cylinder, Ground, Target
    n = 180_000
    rng = np.random.default_rng(17)
    pos = np.stack([rng.normal(0, 6, n), rng.normal(0, 6, n), rng.uniform(0.05, 6.0, n)], 1)
    vel, rpy = rng.normal(0, 2, (n, 3)), rng.uniform(-30, 30, (n, 3))
    objs = [Target(np.array([0.0, 0.0, 3.0]), 1.0), Cylinder(np.array([5.0, -4.0, 0.0]), 1.0, 6.0), Ground()]
    kw = dict(num_envs=n, device=DEV, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
    a, b, c = BatchedDrone(None, **kw), BatchedDrone(None, **kw), BatchedDrone(None, **kw)
    for d in (a, b, c):
        d.reset(pos, vel, rpy)
    a.set_static_objects(objs)
    c.set_static_objects(objs)
    acts = [torch.as_tensor(rng.uniform(-1, 1, (n, 4)), dtype=torch.float32, device=DEV) for _ in range(4)]
    for t in range(12):
        a.step(acts[t % 4], return_obs=False)
        b.step(acts[t % 4], None, objs, return_obs=False)
        c.step(acts[t % 4], return_obs=False, chained=True)
        assert a._fast_ok
    torch.cuda.synchronize()
    assert torch.equal(a._state, b._state) and torch.equal(a._done, b._done)
    assert torch.equal(c._state, b._state) and c.episode_stats()["chain_timeouts"] == 0
    assert a.episode_stats()["crashes"] == b.episode_stats()["crashes"] > 0
    hb = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    with pytest.raises(RuntimeError):
        a.step_host(torch.empty(n, 4, pin_memory=True), hb)
    a.set_static_objects(None)                      # back to the hot path
    a.step(acts[0], return_obs=False)
    b.step(acts[0], return_obs=False)
    assert torch.equal(a._state, b._state)


@pytest.mark.parametrize("parts,n", [(2, 200_000), (4, 70_001), (1, 5_000)])
def test_two_stream_population_equals_one_batch(parts, n):
    """`TwoStreamDrones`: the population split over CUDA streams (closed-loop form of side-by-side launches), policy and step
    of every part on the part's own stream.  The envs are independent, so state, flags and statistics equal those of ONE
    BatchedDrone given the same per-env sticks, bit for bit -- through crashes and restarts, with a state-dependent policy."""
    from fpyv_b200 import BatchedDrone, TwoStreamDrones
    rng = np.random.default_rng(5)
    pos = np.stack([rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(0.05, 3.0, n)], 1)
    vel, rpy = rng.normal(0, 1, (n, 3)), rng.uniform(-30, 30, (n, 3))
    kw = dict(device=DEV, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
    one = BatchedDrone(None, num_envs=n, **kw)
    pop = TwoStreamDrones(None, num_envs=n, parts=parts, **kw)
    assert [d.cta_slots for d in pop.parts] == [4 // parts if parts > 1 else 0] * parts and sum(d.num_envs for d in pop.parts) == n
    one.reset(pos, vel, rpy)
    pop.reset(torch.as_tensor(pos, dtype=torch.float32, device=DEV), vel, rpy)
    noise = torch.as_tensor(rng.uniform(-1, 1, (n, 4)), dtype=torch.float32, device=DEV)

    def sticks(d, extra):          # depends on the CURRENT state: the loop is closed
        a = extra.clone()
        a[:, :3] = (a[:, :3] + 0.5 * d.prev_rates / d.max_rates).clamp(-1, 1)
        a[:, 3] = (a[:, 3] + 0.3 * (1.5 - d.position[:, 2])).clamp(-1, 1)
        return a

    for t in range(25):
        one.step(sticks(one, noise), return_obs=False)
        pop.step(lambda d, i: sticks(d, noise[pop.bounds[i]:pop.bounds[i + 1]]))
    assert torch.equal(pop.position, one.position) and torch.equal(pop.velocity, one.velocity)
    assert torch.equal(pop.quaternion, one.quaternion) and torch.equal(pop.done, one.done)
    a, b = pop.episode_stats(), one.episode_stats()
    assert a["crashes"] == b["crashes"] > 0 and a["env_steps"] == b["env_steps"]
    pop.step_actions(noise)
    one.step(noise, return_obs=False)
    assert torch.equal(pop.position, one.position)
    with pytest.raises(ValueError):
        TwoStreamDrones(None, num_envs=n, parts=3, **kw)


@pytest.mark.parametrize("mode", ["hot", "hot_k1", "general", "acro", "racer"])
def test_results_do_not_depend_on_the_neighbours(mode):
    """Size-independent property of every dynamics mode: stepping a PERMUTED population gives the permuted results, bit for
    bit.  A permutation changes which env shares a thread's packed registers with which (lanes t and t + 32 of a 64-env
    chunk), which envs share a warp (the general path's warp-uniform decisions, the accurate-sin/cos fallbacks taken per
    thread) and which chunk a warp pulls when -- none of it may leak into an env's arithmetic.  Ragged size, crashes and
    restarts inside the run."""
    from fpyv_b200 import BatchedAcroDrone, BatchedDrone, BatchedRacer, Cylinder, Ground, Target
    n, steps = 300_007, 6
    rng = np.random.default_rng(31)
    perm = rng.permutation(n)
    pt = torch.as_tensor(perm, device=DEV)
    pos = np.stack([rng.normal(0, 6, n), rng.normal(0, 6, n), rng.uniform(0.05, 4.0, n)], 1)
    vel, rpy = rng.normal(0, 2, (n, 3)), rng.uniform(-40, 40, (n, 3))
    acts = [torch.as_tensor(rng.uniform(-1, 1, (n, 4)), dtype=torch.float32, device=DEV) for _ in range(steps)]
    if mode in ("hot", "hot_k1", "general"):
        kw = dict(num_envs=n, device=DEV, substeps=1 if mode == "hot_k1" else 8, dt=1e-3, auto_reset=True, thrust_lut=2049)
        a, b = BatchedDrone(None, **kw), BatchedDrone(None, **kw)
        a.reset(pos, vel, rpy)
        b.reset(pos[perm], vel[perm], rpy[perm])
        objs = [Target(np.array([0.0, 0.0, 3.0]), 1.0), Cylinder(np.array([4.0, -3.0, 0.0]), 1.0, 6.0), Ground()] if mode == "general" else None
        for t in range(steps):
            a.step(acts[t], None, objs, return_obs=False)
            b.step(acts[t][pt].contiguous(), None, objs, return_obs=False)
            assert torch.equal(a._done[pt], b._done), t
        assert a.episode_stats()["crashes"] == b.episode_stats()["crashes"] > 0
    elif mode == "acro":
        kw = dict(num_envs=n, device=DEV, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
        a, b = BatchedAcroDrone(None, **kw), BatchedAcroDrone(None, **kw)
        a.reset(pos, vel, rpy)
        b.reset(pos[perm], vel[perm], rpy[perm])
        for t in range(steps):
            da, db = a.step(acts[t]), b.step(acts[t][pt].contiguous())
            assert torch.equal(da[pt], db), t
    else:
        gains = {"roll": [2.0, 0.1, 0.01], "pitch": [2.0, 0.1, 0.01], "yaw": [0.1, 0.0, 0.0]}
        a = BatchedRacer(5, gains, num_envs=n, device=DEV, substeps=8)
        b = BatchedRacer(5, gains, num_envs=n, device=DEV, substeps=8)
        a.reset()
        b.reset()
        for t in range(steps):
            sp = acts[t] * torch.tensor([60.0, 60.0, 20.0, 4.0], device=DEV)      # rate set-points up to 60 rad/s: reduced sin/cos
            a.step(sp)
            b.step(sp[pt].contiguous())
    sa, sb = a._state[:, :n], b._state[:, :n]
    assert torch.equal(sa[:, pt].view(torch.int32), sb.view(torch.int32))
