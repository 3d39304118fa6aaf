"""CPU oracle for the BigSeqKit hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  PARITY UNPINNED (see oracle/bsk_oracle.h).
"""
from .oracle import *  # noqa: F401,F403
