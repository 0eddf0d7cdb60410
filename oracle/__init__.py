"""CPU oracle for the Turing.jl HMC/NUTS hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  See oracle/tnuts_oracle.h for the parity status ("parity unpinned").
"""
from .binding import *  # noqa: F401,F403
