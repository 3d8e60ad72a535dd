"""The three ways a control step can run must agree:

  two_kernels    rollout_cost_kernel -> weight_philox_kernel (+ finalize in its last block)      any size
  fused          step_fused_kernel: ONE cooperative launch, grid-wide barrier                      co-resident grids
  time_parallel  step_tp_kernel: one warp per sample, horizon steps on the lanes (ARM7 / DRONE3)  T <= 64

The reference has a single code path (mppi.py:122-169); these are scheduling choices of the same arithmetic, so the
costs must be bit-identical (fused) or within a few float32 ulps (time-parallel: scans instead of serial sums) and
the update within the stage-isolated tolerance of SURVEY 8(c)."""
import numpy as np
import pytest
import torch

from conftest import rel_inf

pytestmark = pytest.mark.gpu

Q_HOME = [1.57, 1.7, 0.0, 4.4, 0.0, 4.71, 0.0]


def _state(native, model):
    st = np.zeros(native.MODEL_STATE[model], np.float32)
    if model == native.MODEL_ARM7:
        st[:7] = Q_HOME; st[7:14] = [0.05, -0.1, 0.02, 0.3, -0.2, 0.1, -0.05]; st[14:21] = [0.1, -0.2, 2.1, 0, 0, 0, 1]
    elif model == native.MODEL_WB11:
        st[:12] = [0.0, 0.0, 2.1, 0.02, -0.03, 0.1, 0.1, 0.0, -0.05, 0.02, 0.01, -0.03]
        st[12:19] = Q_HOME; st[19:26] = [0.05, -0.1, 0.02, 0.3, -0.2, 0.1, -0.05]
    elif model == native.MODEL_QUAD4:
        st[:12] = [0.1, -0.2, 2.1, 0.05, -0.08, 0.3, 0.2, -0.1, 0.05, 0.1, -0.2, 0.05]
    else:
        st[:6] = [0.1, -0.2, 2.1, 0.3, 0.1, -0.2]
    return st


def _u0(native, model, T):
    u = torch.zeros(T, native.MODEL_NU[model])
    if model == native.MODEL_QUAD4:
        u[:, 0] = 14.7 * 9.81
    if model == native.MODEL_WB11:
        u[:, 0] = 20.2 * 9.81
    return u


@pytest.fixture(scope="module")
def native():
    from quadrotor_manipulator_mppi_b200 import _native
    _native.load()
    return _native


def _run(native, model, K, T, lam, steps=3, **opts):
    from quadrotor_manipulator_mppi_b200.core import NativeSolver
    s = NativeSolver(model, n_samples=K, n_horizon=T, seed=77, lam=lam, **opts)
    s.set_state(_state(native, model))
    s.u_prev = _u0(native, model, T)
    outs = []
    for i in range(steps):
        outs.append(s.step(None).copy())          # warm-started sequence: errors would compound
    return s, outs


MODELS = ["drone", "arm", "quad", "wb"]


def _mid(native, name):
    return {"drone": native.MODEL_DRONE3, "arm": native.MODEL_ARM7, "quad": native.MODEL_QUAD4, "wb": native.MODEL_WB11}[name]


@pytest.mark.parametrize("rounds", [10, 7])
@pytest.mark.parametrize("K,T", [(1, 8), (100, 32), (129, 30), (4097, 16), (32768, 64), (75776, 24), (1000, 128)])
@pytest.mark.parametrize("name", MODELS)
def test_single_launch_step_equals_the_two_kernel_step(name, K, T, rounds, native):
    model = _mid(native, name)
    lam = 0.1 if name in ("arm",) else 50.0                  # several samples alive, so the update is not a single sample's noise
    a, oa = _run(native, model, K, T, lam, fused=1, time_parallel=0, philox_rounds=rounds, steps=1)
    b, ob = _run(native, model, K, T, lam, fused=0, time_parallel=0, philox_rounds=rounds, steps=1)
    assert a.last_path == "fused" and b.last_path == "two_kernels"
    assert torch.equal(a.costs, b.costs)                                       # same rollout arithmetic, bit for bit
    assert rel_inf(a.u_prev.cpu().numpy(), b.u_prev.cpu().numpy()) < 2e-6       # block partition of the sums differs
    # two more warm-started steps: the 1e-7 differences of the updates may not grow
    for s_ in (a, b):
        oa_, ob_ = (oa, ob)
        (oa if s_ is a else ob).extend([s_.step(None).copy(), s_.step(None).copy()])
    assert rel_inf(a.costs.cpu().numpy(), b.costs.cpu().numpy()) < 1e-5
    assert rel_inf(a.u_prev.cpu().numpy(), b.u_prev.cpu().numpy()) < 1e-4
    x, y = oa[0], ob[0]
    assert x[native.MPPI_OUT_RHO] == y[native.MPPI_OUT_RHO]
    assert x[native.MPPI_OUT_ETA] == pytest.approx(y[native.MPPI_OUT_ETA], rel=1e-6)
    assert x[native.MPPI_OUT_ESS] == pytest.approx(y[native.MPPI_OUT_ESS], rel=1e-5)
    assert np.abs(x[:28] - y[:28]).max() <= 2e-6 * max(1.0, np.abs(y[:28]).max())
    for x, y in zip(oa, ob):
        assert x[native.MPPI_OUT_STEP] == y[native.MPPI_OUT_STEP]


def test_single_launch_falls_back_when_the_grid_cannot_be_resident(native):
    """K beyond the resident rows (or T > 256) keeps the two-kernel path; both are the same step."""
    s, _ = _run(native, native.MODEL_WB11, 148 * 4 * 128 + 128, 16, 50.0, steps=1)
    assert s.last_path == "two_kernels"
    s, _ = _run(native, native.MODEL_DRONE3, 1024, 200, 50.0, steps=1, time_parallel=0)
    assert s.last_path == "two_kernels"
    s, _ = _run(native, native.MODEL_DRONE3, 1024, 100, 50.0, steps=1, time_parallel=0, fused=1)
    assert s.last_path == "fused"
    s, _ = _run(native, native.MODEL_DRONE3, 1024, 100, 50.0, steps=1)                          # T <= 128: four steps per lane
    assert s.last_path == "time_parallel"
    s, _ = _run(native, native.MODEL_DRONE3, 1024, 129, 50.0, steps=1)                          # T <= 256: eight steps per lane
    assert s.last_path == "time_parallel"
    s, _ = _run(native, native.MODEL_ARM7, 20000, 32, 0.1, steps=1, fused=1)                   # auto, K beyond the time-parallel range
    assert s.last_path == "fused"
    s, _ = _run(native, native.MODEL_ARM7, 20000, 32, 0.1, steps=1)                            # library defaults
    assert s.last_path == "two_kernels"
    s, _ = _run(native, native.MODEL_ARM7, 2000, 32, 0.1, steps=1)
    assert s.last_path == "time_parallel"
    s, _ = _run(native, native.MODEL_WB11, 2000, 32, 0.1, steps=1)                             # non-linear model: never time-parallel
    assert s.last_path == "two_kernels"


@pytest.mark.parametrize("rounds", [10, 7])
@pytest.mark.parametrize("K,T", [(1, 8), (3, 33), (100, 32), (1000, 30), (1024, 30), (4099, 64), (40000, 20), (517, 65), (1000, 100),
                                 (300, 128), (200, 129), (520, 256)])
@pytest.mark.parametrize("name", ["arm", "drone"])
def test_time_parallel_step_equals_the_thread_per_sample_step(name, K, T, rounds, native, oracle):
    model = _mid(native, name)
    lam = 0.1 if name == "arm" else 2000.0
    a, oa = _run(native, model, K, T, lam, time_parallel=1, philox_rounds=rounds, steps=1)
    b, ob = _run(native, model, K, T, lam, time_parallel=0, philox_rounds=rounds, steps=1)
    assert a.last_path == "time_parallel" and b.last_path in ("fused", "two_kernels")
    Sa, Sb = a.costs.cpu().numpy(), b.costs.cpu().numpy()
    assert rel_inf(Sa, Sb) < 1e-6                             # warp scans / tree sums vs serial sums: a few ulps
    # stage-isolated update: the oracle's weighting on the time-parallel path's OWN costs and the
    # materialised Philox noise of that step
    noise = a.generate_noise(0).cpu().numpy()
    sg = 9 if name == "arm" else 5
    iso = oracle._update(Sa, noise, _u0(native, model, T).numpy(), lam, sg)
    assert rel_inf(a.u_prev.cpu().numpy(), iso["u_new"]) < 1e-5
    assert oa[0][native.MPPI_OUT_RHO] == Sa.min()
    w = np.exp(-(Sa.astype(np.float64) - Sa.min()) / lam)
    assert oa[0][native.MPPI_OUT_ETA] == pytest.approx(w.sum(), rel=1e-5)
    assert oa[0][native.MPPI_OUT_ESS] == pytest.approx(w.sum() ** 2 / (w ** 2).sum(), rel=1e-4)
    # controller outputs come from the same finalize block
    assert np.abs(oa[0][:14] - ob[0][:14]).max() < 1e-5 * max(1.0, np.abs(ob[0][:14]).max()) or lam == 0.1


def test_time_parallel_multi_step_sequence_tracks_the_thread_per_sample_one(native):
    """Ten warm-started steps: the two layouts stay together (no drift from the scan rounding)."""
    a, oa = _run(native, native.MODEL_DRONE3, 1000, 32, 500.0, steps=10, time_parallel=1)
    b, ob = _run(native, native.MODEL_DRONE3, 1000, 32, 500.0, steps=10, time_parallel=0)
    assert rel_inf(a.u_prev.cpu().numpy(), b.u_prev.cpu().numpy()) < 1e-4
    assert rel_inf(oa[-1][:6], ob[-1][:6]) < 1e-5


def test_per_kernel_times_and_path_report(native):
    """SURVEY section 5 tracing hook: CUDA-event times of the step's kernels."""
    from quadrotor_manipulator_mppi_b200.core import NativeSolver
    s = NativeSolver(native.MODEL_WB11, n_samples=200000, n_horizon=32)
    s.set_state(_state(native, native.MODEL_WB11))
    s.u_prev = _u0(native, native.MODEL_WB11, 32)
    with pytest.raises(native.MppiError):
        s.kernel_times()                                      # not enabled yet
    s.profile(True)
    s.step(None)
    t = s.kernel_times()
    assert t["path"] == "two_kernels" and t["rollout_us"] > 20.0 and 0.0 < t["weighting_finalize_us"] < t["rollout_us"]
    s2 = NativeSolver(native.MODEL_ARM7, n_samples=512, n_horizon=30)
    s2.profile(True)
    s2.step(None)
    t2 = s2.kernel_times()
    assert t2["path"] == "time_parallel" and 0.0 < t2["rollout_us"] < 100.0 and t2["weighting_finalize_us"] == 0.0
