"""`from mppi_solver.drone_mppi import MPPI` (drone.py:19) -> the B200 drone controller."""
from quadrotor_manipulator_mppi_b200.mppi_solver.drone_mppi import MPPI  # noqa: F401
