"""TEST INFRASTRUCTURE ONLY.  float64 NumPy model of OUR gate-race reward / termination rules (include/fpv_api.h,
"Multi-agent gate-race environment").  PARITY UNPINNED: the reference has no gate-passing reward;
only the gate plane distance follows it (Gate.calculate_distance, src/utils/components.py:819-822)."""
import numpy as np


def metrics(gates, g, p):
    """gates: [G,7] (c xyz, n xyz, half_size); g: [n] int; p: [n,3] -> (d, r)"""
    c, nrm = gates[g, :3], gates[g, 3:6]
    dp = p - c
    return np.einsum("ij,ij->i", nrm, dp), np.linalg.norm(dp, axis=1)


def reset(gates, p):
    d, r = metrics(gates, np.zeros(len(p), dtype=int), p)
    return np.stack([d, r], 1), np.zeros(len(p), dtype=int), np.zeros(len(p), dtype=int)


def step(gates, p, crashed, prev, g, laps, A, w_gate, w_progress, w_crash, laps_to_finish):
    G = len(gates)
    d, r = metrics(gates, g, p)
    passed = (~crashed) & (prev[:, 0] < 0) & (d >= 0) & ((r * r - d * d) <= gates[g, 6] ** 2)
    reward = np.where(crashed, -w_crash, w_progress * (prev[:, 1] - r) + w_gate * passed)
    g = np.where(passed, g + 1, g)
    wrap = g == G
    laps = laps + wrap
    g = np.where(wrap, 0, g)
    g = np.where(crashed, 0, g)
    laps = np.where(crashed, 0, laps)
    finished = (laps >= laps_to_finish) if laps_to_finish > 0 else np.zeros(len(p), dtype=bool)
    d2, r2 = metrics(gates, g, p)
    rebase = passed | crashed
    prev = np.stack([np.where(rebase, d2, d), np.where(rebase, r2, r)], 1)
    env_reward = reward.reshape(-1, A).sum(1)
    env_done = (crashed | finished).reshape(-1, A).any(1)
    return reward, env_reward, env_done, prev, g, laps
