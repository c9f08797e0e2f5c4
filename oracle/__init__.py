"""Checkers for the hot path.  TEST INFRASTRUCTURE ONLY: import from tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py, never from sgufp_solver_b200/."""
