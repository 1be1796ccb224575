"""CPU oracle for the QR / Cholesky hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (genericlinearalgebra.jl_b200) never does.
Parity status: see the header of gla_oracle.cpp ("parity unpinned" for the two Julia-stdlib
routines; pinned by tests/golden/kat.json and LAPACK cross-checks).
"""
from .oracle import *  # noqa: F401,F403
