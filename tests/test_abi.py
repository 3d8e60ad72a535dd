"""CPU-side checks of the native boundary: the library builds, loads, exports every symbol
include/mppi_b200.h declares, and the ctypes mirror of mppi_config_t matches.  No compute calls."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mppi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mppi_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from quadrotor_manipulator_mppi_b200 import _native, build
    path = build.build()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mppi_b200.h but not exported"
    assert sorted(_native.EXPORTS) == declared


def test_config_struct_mirror_and_defaults():
    from quadrotor_manipulator_mppi_b200 import _native
    assert C.sizeof(_native.MppiConfig) == 344
    arm = _native.default_config(_native.MODEL_ARM7)
    assert (arm.n_samples, arm.n_horizon, arm.savgol_window) == (100, 32, 9)      # mppi.py:40-41,149
    assert abs(arm.sigma[0] - 0.1) < 1e-7 and abs(arm.lambda_ - 0.1) < 1e-7 and abs(arm.dt - 0.01) < 1e-9
    assert list(arm.cost_w)[:4] == [50.0, 30.0, 40.0, 30.0]                      # cost_manager.py:30-33
    assert (arm.torque_kp, arm.torque_kd, arm.cost_flags) == (400.0, 40.0, 0)     # kinova.py:184; the torque law is opt-in
    drone = _native.default_config(_native.MODEL_DRONE3)
    assert (drone.n_samples, drone.n_horizon, drone.savgol_window) == (1000, 32, 5)  # drone_mppi.py:16-17,160
    assert drone.sigma[0] == 30.0 and list(drone.drone_target) == pytest.approx([1.0, 2.0, 3.4], rel=1e-6)
    assert _native.algorithmic_flops(_native.MODEL_WB11) == 850.0       # counted: tests/test_flop_count.py
    with pytest.raises(_native.MppiError):
        _native.default_config(17)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (never route through oracle/)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from quadrotor_manipulator_mppi_b200.mppi_solver.mppi import MPPI
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MPPI(verbose=False)
    from quadrotor_manipulator_mppi_b200 import _native
    h = C.c_void_p()
    cfg = _native.default_config(_native.MODEL_ARM7)
    rc = _native.load().mppi_create(C.byref(cfg), C.byref(h))
    assert rc == 3 and b"no CPU fallback" in _native.load().mppi_last_error(None)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "quadrotor_manipulator_mppi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libmppi_oracle" not in src, f


def test_reference_import_lines_work_through_the_compat_shim():
    """kinova.py:23 / drone.py:19 import lines, verbatim, resolve to the B200 classes."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import quadrotor_manipulator_mppi_b200.compat as c; "
            "sys.path.insert(0, c.PATH)\n"
            "from mppi_solver.mppi import MPPI as A\nfrom mppi_solver.drone_mppi import MPPI as D\n"
            "assert A.__module__ == 'quadrotor_manipulator_mppi_b200.mppi_solver.mppi', A.__module__\n"
            "assert D.__module__ == 'quadrotor_manipulator_mppi_b200.mppi_solver.drone_mppi'\n"
            "import inspect; assert all(p.kind == p.KEYWORD_ONLY for p in list(inspect.signature(A.__init__).parameters.values())[1:])\n"
            "print('ok')") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def _build_c_example(tmp_path):
    import subprocess
    from quadrotor_manipulator_mppi_b200 import build
    lib_dir = os.path.dirname(build.build())
    exe = str(tmp_path / "step_host")
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "step_host.c"), "-L", lib_dir, "-lmppi_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """include/mppi_b200.h compiles as C99 -pedantic; a C caller links against the library and, without a GPU,
    fails loudly instead of falling back."""
    import subprocess
    import torch
    exe = _build_c_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("has a GPU (the run is covered by the gpu-marked test)")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_example_runs_on_the_gpu(tmp_path):
    import subprocess
    r = subprocess.run([_build_c_example(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("step ")]
    assert len(lines) == 5 and "tau[1]=" in lines[-1]
    tau1 = float(lines[0].split("tau[1]=")[1])
    assert 5.0 < abs(tau1) < 40.0              # gravity load on the shoulder joint at the home pose


def test_a_stale_library_is_never_loaded_silently(tmp_path, monkeypatch):
    """the .so travels to the GPU box; it may only be used if it was built from the sources on disk.
    build.py stores a content hash of csrc/ + include/ next to the library and rebuilds on any mismatch."""
    from quadrotor_manipulator_mppi_b200 import build
    build.build()
    assert not build.needs_build()
    assert open(build.HASH).read().strip() == build.source_hash()
    srcs = [os.path.basename(p) for p in build.sources()]
    for must in ("mppi_b200.cu", "mppi_model_unit.cu", "mppi_kernels.cuh", "mppi_device.cuh", "mppi_dynamics.cuh", "arm_inertia_gen.cuh",
                 "fk_tables_gen.cuh", "mppi_launch.cuh", "mppi_host.cuh", "mppi_vec.cuh", "mppi_b200.h"):
        assert must in srcs, must
    good = open(build.HASH).read()
    try:
        with open(build.HASH, "w") as f:
            f.write("0" * 64 + "\n")                 # as if a header had been edited after the build
        assert build.needs_build()
    finally:
        with open(build.HASH, "w") as f:
            f.write(good)
    assert not build.needs_build()
