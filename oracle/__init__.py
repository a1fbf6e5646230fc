"""CPU oracle for dfmir_b200 — TEST INFRASTRUCTURE ONLY (see oracle/dfmir_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  dfmir_b200/ never does.
"""
