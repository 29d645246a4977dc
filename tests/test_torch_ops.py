"""The torch.library binding of the step (fpyv_b200/torch_ops.py): same launch as BatchedDrone.step, declared as a mutating op."""
import pytest
import torch


def test_op_is_registered_and_has_no_cpu_implementation():
    from fpyv_b200 import torch_ops  # noqa: F401
    op = torch.ops.fpyv_b200.drone_step
    schema = str(op.default._schema)
    for mutated in ("state", "done", "acc"):
        assert f"!) {mutated}" in schema          # declared as written in place
    assert "Tensor actions" in schema and "-> ()" in schema
    z = torch.zeros(4, 4, 4)
    with pytest.raises((NotImplementedError, RuntimeError)):
        op(z, torch.zeros(1, 4), torch.zeros(1, dtype=torch.uint8), torch.zeros(1, 4), 12345)


@pytest.mark.gpu
def test_op_matches_the_method_and_traces_without_graph_breaks():
    from fpyv_b200 import BatchedDrone, torch_ops
    n, dev = 5000, "cuda:0"
    g = torch.Generator(device=dev).manual_seed(3)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5 + torch.tensor([0.0, 0.0, 4.0], device=dev)
    vel, rpy = torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30
    a, b, c = (BatchedDrone(None, num_envs=n, device=dev, substeps=8, dt=1e-3, thrust_lut=2049, auto_reset=True) for _ in range(3))
    for d in (a, b, c):
        d.reset(pos, vel, rpy)
    obs = [torch.randn(n, 4, device=dev, generator=g) for _ in range(6)]
    step_b, step_c = torch_ops.bind(b), torch_ops.bind(c)

    def policy_and_step(x, step, state):
        act = torch.tanh(x * 0.7)
        step(act)
        return state[0, :n, 2].sum()          # reads what the op wrote: ordering must survive tracing

    compiled = torch.compile(lambda x: policy_and_step(x, step_c, c._state), backend="aot_eager", fullgraph=True)
    for x in obs:
        a.step(torch.tanh(x * 0.7), return_obs=False)
        zb = policy_and_step(x, step_b, b._state)
        zc = compiled(x)
        assert torch.equal(a._state, b._state) and torch.equal(a._state, c._state)
        assert torch.equal(a.done, b.done) and torch.equal(a.done, c.done)
        assert torch.equal(zb, zc)
    # the op works on the tensors it is given (a functionalising backend passes copies and writes them back)
    st, dn, ac = b._state.clone(), b._done.clone(), b._acc.clone()
    act = torch.tanh(obs[0])
    torch.ops.fpyv_b200.drone_step(st, act, dn, ac, b._op_handle)
    assert torch.equal(b._state, a._state)                       # the drone's own buffers were not touched ...
    a.step(act, return_obs=False)
    assert torch.equal(st, a._state) and torch.equal(dn, a._done) and torch.equal(ac, a._acc)   # ... the copies were stepped
    with pytest.raises(RuntimeError, match="must be a contiguous"):
        torch.ops.fpyv_b200.drone_step(b._state[:, :8], obs[0], b._done, b._acc, b._op_handle)
    torch.library.opcheck(torch.ops.fpyv_b200.drone_step, (b._state, torch.tanh(obs[0]), b._done, b._acc, b._op_handle),
                          test_utils=("test_schema", "test_faketensor"))
