"""`from mppi_solver.mppi import MPPI` (kinova.py:23) -> the B200 arm controller."""
from quadrotor_manipulator_mppi_b200.mppi_solver.mppi import MPPI  # noqa: F401
