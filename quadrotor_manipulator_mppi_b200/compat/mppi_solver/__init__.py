"""Stands in for the reference package `src/mav_mppi/scripts/mppi_solver` (see ../__init__.py)."""
