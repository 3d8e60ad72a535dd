"""Import-path shim: put THIS directory on sys.path and the reference's own import lines work unchanged:

    from mppi_solver.mppi import MPPI          # src/mav_mppi/scripts/kinova.py:23
    from mppi_solver.drone_mppi import MPPI    # src/mav_mppi/scripts/drone.py:19

    import quadrotor_manipulator_mppi_b200.compat as compat; sys.path.insert(0, compat.PATH)
"""
import os

PATH = os.path.dirname(os.path.abspath(__file__))
