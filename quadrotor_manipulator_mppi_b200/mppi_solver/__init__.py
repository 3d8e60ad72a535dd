"""Drop-in replacements for the reference's `mppi_solver` package:

    from mppi_solver.mppi import MPPI          (kinova.py:23)   -> quadrotor_manipulator_mppi_b200.mppi_solver.mppi.MPPI
    from mppi_solver.drone_mppi import MPPI    (drone.py:19)    -> quadrotor_manipulator_mppi_b200.mppi_solver.drone_mppi.MPPI

plus the two models the reference only sketches: `quad_mppi.MPPI` (rigid body, nu=4) and
`wholebody_mppi.MPPI` (quadrotor + arm, nu=11).
"""
