"""CPU oracle for the radix-sort hot path -- TEST INFRASTRUCTURE, never a product fallback.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  See oracle/radix_oracle.c for the parity status.
"""
from .oracle import *  # noqa: F401,F403
