"""TEST INFRASTRUCTURE ONLY -- CPU checkers for the tfluids hot path.

`oracle.Oracle`     : ctypes binding of oracle/liboracle.so (our C restatement,
                      oracle/tfluids_oracle.c).
`oracle.Reference`  : ctypes binding of oracle/_ref/libtfluids_ref.so (the reference's
                      own CPU sources compiled in place, oracle/ref_shim/ref_driver.cc).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (fluidnet_b200) never does.

Both classes expose the same numpy-level API, named after the reference's Lua
operators (torch/tfluids/init.lua): arrays are float32, C-contiguous, 5-D
[b][c][z][y][x].
"""
from .api import (Oracle, Reference, build, have_reference, CellType, ADVECT_METHODS,
                  simulate, model_forward, create_plume_bcs, default_mconf)

__all__ = ["Oracle", "Reference", "build", "have_reference", "CellType", "ADVECT_METHODS",
           "simulate", "model_forward", "create_plume_bcs", "default_mconf"]
