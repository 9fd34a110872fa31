"""CPU oracle for the mirres-b200 hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  PARITY UNPINNED: the reference ships no golden vectors for this path.
"""
