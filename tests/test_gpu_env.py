"""Gate-race env (BASELINE configs[4]) on the GPU against our own float64 model.  Reward rules are OUR definition
(parity unpinned); the gate plane is the reference's (components.py:811-822)."""
import numpy as np
import pytest
import torch

from oracle import gate_env_oracle as go

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gates_array(env):
    return np.array([[*np.asarray(g.position, dtype=np.float64), *np.asarray(g.normal, dtype=np.float64), g.size / 2]
                     for g in env.gates])


@pytest.mark.parametrize("A", [32, 4, 1])
def test_env_rewards_and_reductions_match_model(A):
    from fpyv_b200.env import GateRaceEnv
    from fpyv_b200.objects import Gate
    # small tight track so that gates are actually passed within a few steps
    gates = [Gate([2.0 * i, 0.0, 2.0], np.eye(3), 6.0) for i in range(1, 4)]
    env = GateRaceEnv(None, num_envs=96, agents_per_env=A, device=DEV, substeps=4, dt=2.5e-3, gates=gates,
                      laps_to_finish=1, spawn_height=(1.5, 2.5), seed=3)
    obs = env.reset()
    assert set(obs) == {f"agent_{a}" for a in range(A)} and obs["agent_0"].shape == (96, 16)
    G = gates_array(env)
    n = env.n_agents
    # fling the agents forward through the gates
    env.drone.velocity.copy_(torch.tensor([[14.0, 0.0, 0.5]], device=DEV).expand(n, 3))
    prev, g, laps = go.reset(G, env.drone.position.cpu().numpy().astype(np.float64))
    assert np.allclose(env._prev.cpu().numpy(), prev, atol=1e-5)
    rng = np.random.default_rng(0)
    passed_total, done_total = 0, 0
    for t in range(80):
        act = torch.as_tensor(rng.uniform(-0.2, 0.2, (96, A, 4)), dtype=torch.float32)
        act[..., 3] = -0.55          # ~ hover thrust
        if t == 60:                  # second phase: dive into the ground -> crashes, auto-reset, env terminations
            env.drone.velocity[:, 2] = -40.0
        obs, reward, done, info = env.step(act)
        assert info == {}
        p = env.drone.position.cpu().numpy().astype(np.float64)
        crashed = env.drone.done.cpu().numpy()
        r_a, r_env, d_env, prev, g_new, laps = go.step(G, p, crashed, prev, g, laps, A, 10.0, 1.0, 5.0, 1)
        passed_total += int((g_new != g).sum())
        g = g_new
        assert np.array_equal(env.next_gate.cpu().numpy().reshape(-1), g), t
        assert np.array_equal(env.laps.cpu().numpy().reshape(-1), laps), t
        assert np.allclose(env.agent_reward.cpu().numpy().reshape(-1), r_a, atol=2e-4), t
        assert np.allclose(reward.cpu().numpy(), r_env, atol=2e-4 * A), t
        assert np.array_equal(done.cpu().numpy(), d_env), t
        done_total += int(d_env.sum())
        prev = env._prev.cpu().numpy().astype(np.float64)   # re-base on the fp32 bookkeeping (chaotic thresholds)
    assert passed_total > 50 and done_total > 0
    st = env.episode_stats()
    assert st["reward_sum"] != 0 and st["reward_sq_sum"] > 0


def test_observation_layout():
    from fpyv_b200.env import GateRaceEnv
    env = GateRaceEnv(None, num_envs=8, agents_per_env=4, device=DEV, substeps=1)
    env.reset()
    obs, *_ = env.step(torch.zeros(8, 4, 4))
    o = torch.stack([obs[k] for k in env.agent_names], 1).reshape(-1, 16).cpu().numpy()
    d = env.drone
    R = d.rotation_matrix.cpu().numpy().astype(np.float64)
    p, v = d.position.cpu().numpy(), d.velocity.cpu().numpy()
    G = gates_array(env)
    g = env.next_gate.cpu().numpy().reshape(-1)
    rel = np.einsum("nji,nj->ni", R, G[g, :3] - p)
    assert np.allclose(o[:, 0:3], rel, atol=1e-4)
    assert np.allclose(o[:, 3:6], np.einsum("nji,nj->ni", R, G[g, 3:6]), atol=1e-5)
    assert np.allclose(o[:, 6:9], np.einsum("nji,nj->ni", R, v), atol=1e-4)
    assert np.allclose(o[:, 9:12], R[:, 2, :], atol=1e-5)
    assert np.allclose(o[:, 12:15], d.prev_rates.cpu().numpy(), atol=1e-5)
    assert np.allclose(o[:, 15], d.prev_thrust.cpu().numpy(), atol=1e-5)


def test_config5_shape_runs():
    """262,144 drones = 8,192 envs x 32 agents (one warp per env)."""
    from fpyv_b200.env import GateRaceEnv
    env = GateRaceEnv(None, num_envs=8192, agents_per_env=32, device=DEV, substeps=8, dt=1e-3)
    env.reset()
    g = torch.Generator(device=DEV).manual_seed(0)
    for _ in range(5):
        obs, reward, done, _ = env.step(torch.rand(8192, 32, 4, device=DEV, generator=g) * 2 - 1)
    assert reward.shape == (8192,) and done.shape == (8192,) and torch.isfinite(reward).all()
    assert env.episode_stats()["env_steps"] == 5 * 262144


def test_env_argument_validation():
    from fpyv_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    p = _lib.GateEnvParams()
    p.n_gates, p.agents_per_env = 1, 3
    assert lib.fpv_gate_env_step(C.byref(p), None, 6, 6, None, None, None, None, None, None, None, None, None) == -22
    assert b"power of two" in lib.fpv_last_error()
    p.agents_per_env = 4
    assert lib.fpv_gate_env_step(C.byref(p), None, 6, 6, None, None, None, None, None, None, None, None, None) == -22
    assert b"multiple" in lib.fpv_last_error()


@pytest.mark.parametrize("A", [32, 8])
def test_fused_env_step_is_bit_identical_to_the_two_launch_path(A):
    """fpv_gate_race_step (the env step as the per-chunk epilogue of the packed ring kernel, one launch) against
    fpv_drone_step (packed hot kernel) followed by fpv_gate_env_step: state, rewards, terminations, observations, race
    bookkeeping and the event counters bit for bit over a rollout with gate passes and crashes; the reward sums of the
    statistics are accumulated in a different order (per warp instead of per CTA), so those agree to rounding."""
    from fpyv_b200.env import GateRaceEnv
    envs = 512
    kw = dict(num_envs=envs, agents_per_env=A, device=DEV, substeps=4, dt=2e-3, thrust_lut=2049, seed=3,
              spawn_height=(0.3, 2.5))
    a, b = GateRaceEnv(None, **kw), GateRaceEnv(None, **kw)
    a.reset()
    b.reset()
    g = torch.Generator(device=DEV).manual_seed(9)
    passes = crashes = 0
    for t in range(60):
        act = torch.rand(envs, A, 4, device=DEV, generator=g) * 2 - 1
        act[..., 3] = act[..., 3] * 0.25 - 0.72           # around hover and below: some agents sink to the ground
        oa, ra, da, _ = a.step(act, fused=True)
        ob, rb, db, _ = b.step(act, fused=False)
        assert torch.equal(ra, rb) and torch.equal(da, db), t
        assert torch.equal(a._obs, b._obs), t
        crashes += int(b.drone.done.sum())
    torch.cuda.synchronize()
    assert torch.equal(a.drone._state, b.drone._state)
    assert torch.equal(a._progress, b._progress) and torch.equal(a._prev, b._prev)
    assert torch.equal(a.agent_reward, b.agent_reward)
    sa, sb = a.episode_stats(), b.episode_stats()
    for k in sa:
        if k in ("reward_sum", "reward_sq_sum"):
            assert abs(sa[k] - sb[k]) <= 1e-6 * max(1.0, abs(sb[k])), (k, sa[k], sb[k])
        else:
            assert sa[k] == sb[k] or (sa[k] != sa[k] and sb[k] != sb[k]), (k, sa, sb)
    assert crashes > 0


def test_chained_fused_env_steps_are_bit_identical():
    """GateRaceEnv.step(fused=True, chained=True): launches that overlap the end of the previous one (FPV_F_CHAINED; the env's
    per-agent arrays are ordered per 64-agent chunk together with the state) against plain stream order, BASELINE configs[4]
    size, with crashes, restarts and gate passes in the rollout."""
    from fpyv_b200.env import GateRaceEnv
    envs, A = 8192, 32
    kw = dict(num_envs=envs, agents_per_env=A, device=DEV, substeps=8, dt=1e-3, thrust_lut=2049, seed=5, spawn_height=(0.3, 2.5))
    a, b = GateRaceEnv(None, **kw), GateRaceEnv(None, **kw)
    a.reset()
    b.reset()
    g = torch.Generator(device=DEV).manual_seed(11)
    acts = [(torch.rand(envs, A, 4, device=DEV, generator=g) * 2 - 1) for _ in range(6)]
    for x in acts:
        x[..., 3] = x[..., 3] * 0.25 - 0.72
    ra, rb = [], []
    for t in range(40):
        _, r1, d1, _ = a.step(acts[t % 6], fused=True, chained=True)
        ra.append((r1.clone(), d1.clone()))
    for t in range(40):
        _, r2, d2, _ = b.step(acts[t % 6], fused=True)
        rb.append((r2.clone(), d2.clone()))
    torch.cuda.synchronize()
    for t, ((r1, d1), (r2, d2)) in enumerate(zip(ra, rb)):
        assert torch.equal(r1, r2) and torch.equal(d1, d2), t
    assert torch.equal(a.drone._state, b.drone._state) and torch.equal(a._obs, b._obs)
    assert torch.equal(a._progress, b._progress) and torch.equal(a._prev, b._prev)
    assert a.episode_stats()["crashes"] == b.episode_stats()["crashes"] > 0
    assert a.episode_stats()["chain_timeouts"] == 0
