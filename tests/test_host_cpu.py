"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/fpv_api.h declares,
struct layouts agree, config front-end parses the reference's file formats, host logic of objects/sticks."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import CONFIG, GOLDEN, ROOT


@pytest.fixture(scope="module")
def lib():
    from fpyv_b200 import build, _lib
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "fpv_api.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(fpv_[a-z_0-9]+)\s*\(", hdr))
    assert {"fpv_drone_step", "fpv_drone_reset", "fpv_racer_step", "fpv_sticks_to_actions"} <= names
    raw = C.CDLL(os.path.join(ROOT, "fpyv_b200", "libfpyv_b200.so"))
    for n in sorted(names):
        assert hasattr(raw, n), f"{n} declared in fpv_api.h but not exported"
    from fpyv_b200 import _lib
    assert set(_lib.EXPORTS) == names


def test_abi_version_and_struct_sizes(lib):
    from fpyv_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "fpv_api.h")).read()
    assert int(re.search(r"#define FPV_ABI_VERSION (\d+)", hdr).group(1)) == lib.fpv_abi_version() == _lib.ABI_VERSION
    for i, st in enumerate(_lib._STRUCTS):
        assert lib.fpv_sizeof(i) == C.sizeof(st), st.__name__
    assert lib.fpv_sizeof(99) == -1
    for name, val in (("FPV_F_GROUND", _lib.F_GROUND), ("FPV_F_AUTO_RESET", _lib.F_AUTO_RESET),
                      ("FPV_F_FREEZE_DONE", _lib.F_FREEZE_DONE), ("FPV_F_THRUST_LUT", _lib.F_THRUST_LUT),
                      ("FPV_F_SCALAR", _lib.F_SCALAR),
                      ("FPV_MAX_OBJECTS", _lib.MAX_OBJECTS), ("FPV_DRONE_PLANES", _lib.DRONE_PLANES),
                      ("FPV_RACER_PLANES", _lib.RACER_PLANES)):
        assert int(re.search(rf"#define {name} (\d+)", hdr).group(1)) == val, name


def test_argument_validation_needs_no_gpu(lib):
    """EINVAL paths return before any CUDA call."""
    from fpyv_b200 import _lib
    p, io = _lib.DroneParams(), _lib.DroneIO()
    assert lib.fpv_drone_step(None, None, None) == -22
    io.n, io.plane_stride = 8, 4
    assert lib.fpv_drone_step(C.byref(p), C.byref(io), None) == -22
    assert b"stride" in lib.fpv_last_error()
    io.n = 0
    io.plane_stride = 0
    assert lib.fpv_drone_step(C.byref(p), C.byref(io), None) == 0       # empty batch is a no-op
    assert lib.fpv_racer_step(None, None, 0, 0, None, None, None, None) == -22
    assert lib.fpv_sticks_to_actions(None, None, 0, None, None, None) == -22


def test_no_cpu_fallback():
    import torch
    from fpyv_b200 import BatchedDrone
    with pytest.raises(RuntimeError):
        BatchedDrone(None, num_envs=2, device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            BatchedDrone(None, num_envs=2, device="cuda:0")


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "fpyv_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_config_constants_match_reference_golden():
    from fpyv_b200 import config
    g = np.load(os.path.join(GOLDEN, "consts.npz"))
    params = config.load_params()
    c = config.derive_constants(params)
    assert c.dt == float(g["dt"]) and c.mass == float(g["mass"])
    np.testing.assert_allclose(c.thrust_poly.coeffs, g["poly"], rtol=1e-12)
    np.testing.assert_allclose(c.throttle_poly.coeffs, g["inv_poly"], rtol=1e-12)
    np.testing.assert_allclose(c.motors_relative_position, g["motors_relative_position"], atol=1e-16)
    np.testing.assert_allclose(c.cross_section_areas, g["cross_section_areas"], rtol=1e-15)
    np.testing.assert_allclose(c.throttle2thrust(g["t2t_x"]), g["t2t_y"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(c.thrust2throttle(g["inv_x"]), g["inv_y"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose([c.min_throttle_in_force, c.max_throttle_in_force], [g["min_force"], g["max_force"]], rtol=1e-13)
    np.testing.assert_allclose(c.thrust_newton, g["thrust_n"], rtol=1e-15)
    assert c.motor_name == "F80 Pro KV1900" and c.propeller == "5055 Tri-Blade"


def test_motor_report_reader_handles_format_quirks(tmp_path):
    from fpyv_b200 import config
    sweeps = config.read_motor_test_report(os.path.join(CONFIG, "t_motos_f80_motor_test.csv"))
    assert len(sweeps) == 5 and all(len(s.throttle_percent) == 11 for s in sweeps)
    assert sweeps[2].thrust_grams[2] == 993.47          # quoted decimal comma "993,47"
    p = tmp_path / "m.csv"
    p.write_text('Type,Propeller,Throttle,Thrust (g)\nX,P,50%,"10,5"\n,,100%,20\n,,50%,11\n,,100%,21\n')
    s = config.read_motor_test_report(str(p))
    assert len(s) == 2 and s[0].thrust_grams.tolist() == [10.5, 20.0] and s[0].motor == "X"


def test_windows_paths_resolve_next_to_params(tmp_path):
    from fpyv_b200 import config
    assert config.resolve_path(r"C:\Users\omri_\PycharmProjects\FpyV\config\frsky.json").endswith("frsky.json")
    (tmp_path / "frsky.json").write_text("{}")
    assert config.resolve_path(r"C:\x\frsky.json", (str(tmp_path),)) == str(tmp_path / "frsky.json")
    with pytest.raises(FileNotFoundError):
        config.resolve_path(r"C:\x\nope.json")
    params = config.load_params()
    params["drone"]["motor_test_report_path"] = r"C:\Users\omri_\PycharmProjects\FpyV\config\t_motos_f80_motor_test.csv"
    assert config.derive_constants(params).max_throttle_in_force == pytest.approx(81.30229036293663)


def test_thrust_table():
    from fpyv_b200 import config
    c = config.derive_constants(config.load_params())
    t = config.thrust_table(c, 2049)
    x = np.linspace(-1, 1, 20001)
    lin = np.interp(x, np.linspace(-1, 1, 2049), t.astype(np.float64))
    assert np.max(np.abs(lin - c.throttle2thrust(x))) < 6e-6
    b = config.thrust_table(c, 1025, "bench")
    assert b[0] == 0 and b[-1] == pytest.approx(c.thrust_newton[-1], rel=1e-6)
    with pytest.raises(ValueError):
        config.thrust_table(c, 16, "nope")


def test_stick_calibration_json():
    from fpyv_b200 import config
    cal = config.StickCalibration.load(os.path.join(CONFIG, "frsky.json"))
    assert cal.stick_idx == [0, 1, 2, 5] and list(cal.sticks) == ["Throttle", "Roll", "Pitch", "Yaw"]
    assert cal.max_vals[0] == 48371
    with pytest.raises(FileNotFoundError):
        config.StickCalibration.load("/nonexistent.json")


def test_object_list_lowering():
    from fpyv_b200 import objects, _lib
    g, lowered = objects.lower_object_list([objects.Target([1, 2, 3], 0.5), objects.Gate([0, 0, 0], np.eye(3), 2.0),
                                            objects.Cylinder([4, 5, 0], 1.0, 6.0), objects.Ground()])
    assert g and [o.kind for o in lowered] == [_lib.OBJ_SPHERE, _lib.OBJ_CYLINDER]
    assert lowered[1].b == 6.0
    with pytest.raises(ValueError):
        objects.lower_object_list([objects.Ground(), objects.Target([0, 0, 0], 1)])
    with pytest.raises(ValueError):
        objects.lower_object_list([objects.Target([0, 0, 0], 1)] * 17)
    with pytest.raises(AssertionError):
        objects.Cylinder([0, 0, 0], -1, 1)
    gate = objects.Gate([1.0, 0, 0], np.eye(3), 2.0)
    assert gate.calculate_distance(np.array([3.0, 5, 5])) == pytest.approx(2.0)


def test_sincos_polynomial_accuracy():
    """The small-angle kernels in csrc/vec.cuh, evaluated here in float32 with the same coefficients."""
    x = np.linspace(-0.7854, 0.7854, 200001).astype(np.float32)
    x2 = x * x
    ps = np.float32(-1.95152959e-4) * x2 + np.float32(8.33216087e-3)
    ps = ps * x2 + np.float32(-1.66666546e-1)
    s = ps * (x2 * x) + x
    pc = np.float32(2.44331571e-5) * x2 + np.float32(-1.38873163e-3)
    pc = pc * x2 + np.float32(4.16666457e-2)
    pc = pc * x2 + np.float32(-0.5)
    c = pc * x2 + np.float32(1.0)
    assert np.max(np.abs(s - np.sin(x.astype(np.float64)))) < 1.5e-7
    assert np.max(np.abs(c - np.cos(x.astype(np.float64)))) < 1.5e-7


def test_generate_track_matches_reference_layout():
    """generators.generate_track (src/utils/generators.py:7-18) incl. its size quirk."""
    from fpyv_b200.objects import generate_track
    gates = generate_track(count=6, radius=12, gate_size=5, gate_resolution=17)
    assert len(gates) == 6
    th = np.linspace(0, 2 * np.pi, 7)[:-1]
    for i, g in enumerate(gates):
        base = np.array([np.cos(th[i]) * 5, np.sin(th[i]) * 12, 0.0])
        if i % 3 == 1:      # circle
            assert np.allclose(g.position, base + [0, 0, 2.5]) and g.size == 2.5
        else:
            assert np.allclose(g.position, base) and g.size == 17
        yaw = th[i] + np.pi / 2
        assert np.allclose(g.normal, [np.cos(yaw), np.sin(yaw), 0.0])


def test_gate_env_model_basics():
    from oracle import gate_env_oracle as go
    gates = np.array([[2.0, 0, 0, 1, 0, 0, 1.0], [4.0, 0, 0, 1, 0, 0, 1.0]])
    p0 = np.array([[1.0, 0.2, 0.0], [1.0, 3.0, 0.0]])
    prev, g, laps = go.reset(gates, p0)
    p1 = p0 + [1.5, 0, 0]
    r, re, de, prev, g, laps = go.step(gates, p1, np.array([False, False]), prev, g, laps, 2, 10.0, 1.0, 5.0, 0)
    assert g.tolist() == [1, 0] and r[0] > 10 and r[1] < 10 and re[0] == r.sum() and not de[0]


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: the header must compile as C (no C++-isms, no CUDA types)."""
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "fpv_api.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_object_point_clouds_match_the_reference_shapes():
    """Ground / Cylinder / Gate point clouds (components.py:655-667, :697-708, :787-805) against the arrays the
    reference itself produced for the golden world (oracle/make_golden_chase.py)."""
    from fpyv_b200 import Cylinder, Gate, Ground, Target
    from fpyv_b200.objects import icosphere_vertices
    g = np.load(os.path.join(GOLDEN, "chase_camera.npz"))
    np.testing.assert_allclose(Ground(40, 24, random=False).points, g["obj4"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(Cylinder(np.array([6.0, 2.0, 0.0]), 1.5, 8.0, 10, 12).points, g["obj1"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(Cylinder(np.array([-4.0, -7.0, 0.0]), 2.0, 5.0, 8, 9).points, g["obj2"], rtol=0, atol=1e-12)
    yaw = 0.7
    rot = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
    np.testing.assert_allclose(Gate(np.array([3.0, -5.0, 2.5]), rot, 5.0, shape="circle", resolution=17).points, g["obj3"],
                               rtol=0, atol=1e-12)
    for nu in (1, 2, 5):
        v = icosphere_vertices(nu)
        assert v.shape == (10 * nu * nu + 2, 3)
        np.testing.assert_allclose(np.linalg.norm(v, axis=1), 1.0, rtol=0, atol=1e-12)
    t = Target(np.array([1.0, 2.0, 3.0]), 0.5, nu=2)
    np.testing.assert_allclose(np.linalg.norm(t.points - t.position, axis=1), 0.5, rtol=0, atol=1e-12)
    assert t.calculate_distance(np.array([1.0, 2.0, 5.0])) == pytest.approx(1.5)
    box = Ground(40, 24).bbox3d
    assert box.shape == (8, 3) and box[:, 0].min() == -20 and box[:, 0].max() == 20
    rg = Ground(60, 50, random=True, rng=np.random.default_rng(0)).points
    assert rg.shape == (2500, 3) and np.abs(rg[:, :2]).max() <= 60 and np.abs(rg[:, 2]).max() <= 0.2


def test_new_entry_points_validate_without_a_gpu(lib):
    from fpyv_b200 import _lib
    cam, ap, ac = _lib.CameraParams(), _lib.AutopilotParams(), _lib.AcroParams()
    assert lib.fpv_camera_update(None, None, 0, 0, None, None) == -22
    assert lib.fpv_camera_update(C.byref(cam), None, 0, 0, None, None) == -22 and b"resolution" in lib.fpv_last_error()
    cam.width, cam.height, cam.fx, cam.fy = 641, 481, 100.0, 100.0
    dummy = (C.c_double * 64)()
    img = (C.c_uint8 * 64)()
    assert lib.fpv_camera_render(C.byref(cam), dummy, 1, dummy, 1, dummy, 1, None, 10.0, img, img, None) == -22
    assert b"multiple of 4" in lib.fpv_last_error()
    cam.width, cam.height = 640, 480
    assert lib.fpv_camera_render(C.byref(cam), dummy, 1, dummy, 1, dummy, 65, None, 10.0, img, img, None) == -22
    assert lib.fpv_camera_target_pixel(C.byref(cam), dummy, 1, dummy, 1, dummy, 1, None, 0.0, dummy, img, None) == -22
    assert lib.fpv_camera_rays(C.byref(cam), dummy, 1, dummy, 7, dummy, None) == -22 and b"world, drone or camera" in lib.fpv_last_error()
    ap.ref_frame = 5
    ap.dt = 0.01
    assert lib.fpv_autopilot(C.byref(ap), C.byref(cam), dummy, 0, 0, dummy, None, dummy, dummy, dummy, None, None, None, None) == -22
    assert b"Unknown reference frame" in lib.fpv_last_error()
    assert lib.fpv_acro_step(C.byref(ac), dummy, 0, 0, dummy, None, 0, None, None, None, None, None, None) == -22
    assert lib.fpv_acro_reset(None, 0, 0, None, None, None, None, None) == -22


def test_host_and_rollout_entry_points_validate_without_a_gpu(lib):
    from fpyv_b200 import _lib
    p, io = _lib.DroneParams(), _lib.DroneIO()
    buf = (C.c_float * 64)()
    assert lib.fpv_drone_step_host(None, None, None, None, 4, None) == -22
    assert lib.fpv_drone_step_host(C.byref(p), C.byref(io), None, None, 4, None) == -22 and b"host buffer" in lib.fpv_last_error()
    assert lib.fpv_drone_step_host(C.byref(p), C.byref(io), buf, buf, 4, None) == -22 and b"staging" in lib.fpv_last_error()
    assert lib.fpv_drone_rollout(None, None, None, 0, 1, None, 0, None) == -22
    io.n = io.plane_stride = 0
    assert lib.fpv_drone_rollout(C.byref(p), C.byref(io), None, 0, 4, None, 0, None) == 0      # empty batch: no-op


@pytest.mark.gpu
@pytest.mark.parametrize("slots", [1, 2, 3])
def test_cta_slots_do_not_change_results(slots):
    """max_ctas_per_sm only changes how much of each SM a launch occupies: results are bit-identical."""
    import torch
    from fpyv_b200 import BatchedDrone
    n, dev = 1 << 19, "cuda:0"
    g = torch.Generator(device=dev).manual_seed(4)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5
    pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
    vel = torch.randn(n, 3, device=dev, generator=g)
    rpy = (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30
    acts = [torch.rand(n, 4, device=dev, generator=g) * 2 - 1 for _ in range(6)]
    ref = BatchedDrone(None, num_envs=n, device=dev, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
    a = BatchedDrone(None, num_envs=n, device=dev, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049, cta_slots=slots)
    b = BatchedDrone(None, num_envs=n, device=dev, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049, cta_slots=slots)
    for d in (ref, a, b):
        d.reset(pos, vel, rpy)
    assert a.cta_slots == slots
    for t in range(6):      # a and b alternate, chained: their launches run side by side
        ref.step(acts[t], return_obs=False)
        a.step(acts[t], return_obs=False, chained=True)
        b.step(acts[t], return_obs=False, chained=True)
    torch.cuda.synchronize()
    assert torch.equal(a._state, ref._state) and torch.equal(b._state, ref._state)
    assert torch.equal(a.done, ref.done) and a.episode_stats()["crashes"] == ref.episode_stats()["crashes"]


def test_plain_c_consumer_of_the_abi(lib, tmp_path):
    """A C99 program (tests/c/abi_smoke.c) dlopens the library, compares sizeof() of every ABI struct as the C compiler
    lays it out with what the library was built with, and drives the argument validation."""
    import subprocess
    exe = tmp_path / "abi_smoke"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-ldl", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), os.path.join(ROOT, "fpyv_b200", "libfpyv_b200.so")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "10 struct layouts agree" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("slots,K", [(1, 8), (2, 8), (2, 1), (3, 4)])
def test_same_batch_chained_on_partial_grids_stays_correct(slots, K):
    """Chaining consecutive steps of ONE batch on grids that take only some CTA slots lets up to five launches be alive
    at once, each waiting chunk by chunk on its predecessor (slow, but it must be right): 60 chained steps bit-identical
    to plain steps, crossing the uint32 wrap of the epoch counter on the way."""
    import torch
    from fpyv_b200 import BatchedDrone
    n, dev = 1 << 19, "cuda:0"
    g = torch.Generator(device=dev).manual_seed(5)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5
    pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
    vel = torch.randn(n, 3, device=dev, generator=g)
    rpy = (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30
    acts = [torch.rand(n, 4, device=dev, generator=g) * 2 - 1 for _ in range(4)]
    ref = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049, cta_slots=slots)
    ref.reset(pos, vel, rpy)
    d.reset(pos, vel, rpy)
    d._epoch = 0xFFFFFFE0
    d._chunk_epoch.fill_(-32)
    for t in range(60):
        ref.step(acts[t % 4], return_obs=False)
        d.step(acts[t % 4], return_obs=False, chained=True)
    torch.cuda.synchronize()
    assert torch.equal(d._state, ref._state)
    assert d.episode_stats()["crashes"] == ref.episode_stats()["crashes"]
    assert int(d._chunk_epoch[0]) & 0xFFFFFFFF == (0xFFFFFFE0 + 60) & 0xFFFFFFFF


def test_sass_carries_the_blackwell_paths(lib):
    """The built library must really contain what DESIGN.md claims for the hot kernel: packed FP32 math (FFMA2 / FMUL2 /
    FADD2), TMA bulk copies completing on mbarriers (UBLKCP, SYNCS), programmatic dependent launch (ACQBULK / PREEXIT) and
    the release / acquire pair of the chained-launch protocol -- a toolchain or flag regression that silently fell back to
    scalar math or plain loads would pass every numerical test."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "fpyv_b200", "libfpyv_b200.so")
    out = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True).stdout
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", so], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)[1:]
    hot = [f for f in funcs if f.startswith("_ZN3fpv16ring_step_kernelINS_9DroneModeINS_2F2ELi4ELb0ENS_6NoPostENS_7DroneIOELi0E")]
    assert len(hot) == 1
    sass = hot[0]
    count = lambda op: len(re.findall(r"\b" + op + r"\b", sass))
    assert count("FFMA2") >= 100 and count("FMUL2") >= 40 and count("FADD2") >= 6, (count("FFMA2"), count("FMUL2"), count("FADD2"))
    assert count(r"UBLKCP\.S\.G") >= 5
    assert "SYNCS.ARRIVE.TRANS64" in sass and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass
    assert "ACQBULK" in sass and "PREEXIT" in sass
    assert "LDG.E.STRONG.GPU" in sass and "MEMBAR.ALL.GPU" in sass and "FENCE.VIEW.ASYNC" in sass
    # the rollout kernel and every other mode of the ring kernel (obstacle path, gate-race epilogue, Racer, acro) use the
    # packed pipe and the TMA ring as well
    f = [x for x in funcs if x.startswith("_ZN3fpv20drone_rollout_kernelINS_2F2ELi4")]
    assert f and len(re.findall(r"\bFFMA2\b", f[0])) >= 60
    for prefix in ("_ZN3fpv16ring_step_kernelINS_9DroneModeINS_2F2ELi4ELb1ENS_6NoPost", "_ZN3fpv16ring_step_kernelINS_9DroneModeINS_2F2ELi4ELb0ENS_8GatePost",
                   "_ZN3fpv16ring_step_kernelINS_9RacerModeINS_2F2E", "_ZN3fpv16ring_step_kernelINS_8AcroModeINS_2F2E",
                   "_ZN3fpv16ring_step_kernelINS_9DroneModeINS_2F2ELi4ELb0ENS_6NoPostENS_7DroneIOELi1E",       # raw uint16 sticks in the ring
                   "_ZN3fpv16ring_step_kernelINS_9DroneModeINS_2F2ELi4ELb0ENS_6NoPostENS_7DroneIOELi2E"):      # CRSF sticks in the ring
        f = [x for x in funcs if x.startswith(prefix)]
        assert f and len(re.findall(r"\bFFMA2\b", f[0])) >= 60, prefix
        assert len(re.findall(r"UBLKCP\.S\.G", f[0])) >= 5 and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in f[0], prefix


@pytest.mark.parametrize("n", [1, 63, 65536, 262143, 262144, 300_000, 1 << 20, (1 << 20) + 1, 16_777_216])
@pytest.mark.parametrize("slices", [0, 1, 2, 4, 8, 16, 99])
def test_host_step_slices_cover_the_batch(n, slices):
    """The env ranges of the pipelined host step (mirror of host_slice_bounds in fpv_api.cu): contiguous, covering,
    starting on 64-env boundaries, none shorter than 65,536 envs unless it is the whole remainder, at most 16 + 1."""
    from fpyv_b200.drone import host_slice_bounds
    b = host_slice_bounds(n, slices)
    want = min(16, slices) if slices > 0 else 4
    assert b[0][0] == 0 and b[-1][1] == n and 1 <= len(b) <= want
    for (a0, a1), (b0, b1) in zip(b, b[1:]):
        assert a1 == b0
    for a0, a1 in b:
        assert a0 % 64 == 0 and a1 > a0
        assert len(b) == 1 or a1 - a0 >= 65536 - 64 * 16
    if want >= 2 and n >= 4 * 65536:
        assert len(b) >= 2 and 65536 <= b[-1][1] - b[-1][0] < b[0][1] - b[0][0] + 64      # the tail range is the short one
        assert b[-1][1] - b[-1][0] < max(65536, n // 16) + 64
    if n >= 16 * 65536:
        assert len(b) == want
