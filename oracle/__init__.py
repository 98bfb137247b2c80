"""CPU oracle of the ActiveGS hot path.  TEST INFRASTRUCTURE ONLY: importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg -- never from the
product package (active_gs_b200 / diff_gaussian_rasterization_2d)."""
