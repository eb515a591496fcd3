"""CPU oracle for the SimpleTetris step path — TEST INFRASTRUCTURE, not product.

See `st_oracle.c` (the C restatement), `oracle.py` (ctypes front end) and
`ref_shim.py` (loader of the unmodified reference, build container only).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import it.
"""
