"""Soak: two-stream closed-loop population + chained rotation + gate race chained, 20k launches each; check error words."""
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from fpyv_b200 import TwoStreamDrones, BatchedDrone
from fpyv_b200.env import GateRaceEnv
dev = "cuda:0"
n = 1 << 20
pop = TwoStreamDrones(None, num_envs=n, device=dev, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
g = torch.Generator(device=dev).manual_seed(0)
pos = torch.randn(n, 3, device=dev, generator=g) * 5; pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
pop.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
acts = [torch.rand(n, 4, device=dev, generator=g) * 2 - 1 for _ in range(4)]
t0 = time.time()
for i in range(10000):
    a = acts[i % 4]
    pop.step(lambda d, j: a[pop.bounds[j]:pop.bounds[j + 1]])
pop.join(); torch.cuda.synchronize()
st = pop.episode_stats()
print(f"two streams: 10,000 steps in {time.time()-t0:.2f} s; env_steps {st['env_steps']:.4g} crashes {st['crashes']:.4g} nonfinite {st['nonfinite']} chain_timeouts {st['chain_timeouts']}")
env = GateRaceEnv(None, num_envs=8192, agents_per_env=32, device=dev, substeps=8, dt=1e-3, thrust_lut=2049)
env.reset()
a = torch.rand(8192, 32, 4, device=dev, generator=g) * 2 - 1
t0 = time.time()
for i in range(20000):
    env.step(a, fused=True, chained=True)
torch.cuda.synchronize()
st = env.episode_stats()
print(f"gate race chained: 20,000 steps in {time.time()-t0:.2f} s; crashes {st['crashes']:.4g} nonfinite {st['nonfinite']} chain_timeouts {st['chain_timeouts']} reward_sum {st['reward_sum']:.6g}")
