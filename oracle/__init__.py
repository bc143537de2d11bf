"""TEST INFRASTRUCTURE ONLY: CPU oracle + compiled reference for the scalar-advection path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product package maestroex_b200 never does.
"""
