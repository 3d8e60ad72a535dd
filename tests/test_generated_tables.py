"""The generated constant tables (FK chain, arm inertias) are current and agree with the oracle's numbers."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

REF_URDF = "/root/reference/src/aerial_manipulation/urdf/full_robot_floating2.urdf"


def _run(script, *args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", script), *args], capture_output=True, text=True)


def test_fk_tables_are_current():
    r = _run("gen_fk_tables.py", "--check")
    assert r.returncode == 0, r.stdout + r.stderr


def test_fk_tables_reproduce_the_oracle_chain(oracle):
    """Folded constants C0 Rz(q1) C1 ... Rz(q7) C7 evaluated in float64 against the oracle's unfolded FK."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import gen_fk_tables as g
    finally:
        sys.path.pop(0)
    Cs = g.fold(g.KINOVA)
    q = np.array([1.2, 2.0, -0.4, 4.0, 0.7, 4.2, -1.0])
    T = np.eye(4)
    for j in range(7):
        c, s = np.cos(q[j]), np.sin(q[j])
        Rz = np.eye(4)
        Rz[:2, :2] = [[c, -s], [s, c]]
        T = T @ Cs[j] @ Rz
    T = T @ Cs[7]
    assert np.allclose(T, oracle.fk(q.astype(np.float32)), atol=2e-6)


@pytest.mark.skipif(not os.path.exists(REF_URDF), reason="the reference URDF is only present in the build container")
def test_arm_inertia_tables_are_current():
    r = _run("gen_arm_inertia.py", "--check")
    assert r.returncode == 0, r.stdout + r.stderr


def test_arm_inertia_header_matches_the_oracle_module():
    from oracle import arm_inertia_gen as gen
    txt = open(os.path.join(ROOT, "quadrotor_manipulator_mppi_b200", "csrc", "arm_inertia_gen.cuh")).read()
    for v in gen.MASS + [x for row in gen.COM for x in row] + [x for row in gen.INERTIA for x in row]:
        assert repr(float(v)) + "f" in txt
