"""CPU oracle for the FFVD hot path -- TEST INFRASTRUCTURE, never imported by the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may use it.  See oracle/ffvd_oracle.py for the reference file:line map.
"""
