"""TEST INFRASTRUCTURE -- CPU checkers for the plonk.c prove/verify path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package.  The product (plonk.c_b200/) never does.

  oracle.ref   ctypes binding of oracle/_ref/libref_oracle.so  = the UNMODIFIED reference headers
               behind a batch driver (kind "reference")
  oracle.port  ctypes binding of oracle/libplonk_port.so       = the C restatement (kind "port")

Both expose the same batch functions over the same byte layouts, so tests can swap them.
"""
