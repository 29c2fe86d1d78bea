"""CPU oracle for the ALDI++ hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything under oracle/.  The product path (aldi_b200/) never does.
"""
