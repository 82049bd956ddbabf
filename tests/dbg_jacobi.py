"""Kernel-tuning script (not a test): 100 Jacobi sweeps at n^3, CUDA-event timed."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fluidnet_b200 import tfluids, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = tfluids.context()
fl = torch.from_numpy(synth.make_flags(n, n, n, True, nb=1, geometry=True)).cuda()
dv = torch.randn(1, 1, n, n, n, device="cuda") * (fl == 1)
pj = torch.zeros_like(fl)
ts = []
for it in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ctx.use_current_stream()
    ctx.check(ctx.lib.tfl_solve_linear_system_jacobi(ctx.h, tfluids._grid(pj), tfluids._grid(fl), tfluids._grid(dv), 1, 0.0, 100, None, None))
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.median(ts[1:]))
print("zchunk %s: Jacobi x100 at %d^3: %.3f ms  (%.2f us / sweep, %.0f GB/s algorithmic)" % (
    os.environ.get("TFL_JACOBI_ZCHUNK", "default"), n, ms, ms * 10, 16.0 * n ** 3 * 100 / ms / 1e6))
