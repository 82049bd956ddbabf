"""Debug / timing driver for the PCG solve (not a test): python tests/dbg_pcg.py [small|big]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import api
import pcg_cases
from fluidnet_b200 import tfluids, synth

mode = sys.argv[1] if len(sys.argv) > 1 else "small"
orc = api.Oracle()
if mode == "small":
    for is3d in (True, False):
        flags, U, div = pcg_cases.make(orc, is3d)
        for precond in ("none", "ic0"):
            want = np.zeros(flags.shape, np.float32)
            rw = orc.solveLinearSystemPCG(want, flags, div, is3d, 1e-5, 1000, precond)
            p = torch.zeros(flags.shape, device="cuda")
            rg = tfluids.solveLinearSystemPCG(p, torch.from_numpy(flags).cuda(), torch.from_numpy(div).cuda(), is3d,
                                              1e-5, 1000, precond)
            got = p.cpu().numpy()
            print(is3d, precond, "oracle res %g it %d | gpu res %g it %d | max diff %g of %g" % (
                rw, orc.last_pcg_iters, rg, tfluids.solveLinearSystemPCG.last_iterations,
                np.abs(got - want).max(), np.abs(want).max()), flush=True)
else:
    for n in (64, 128):
        flags = synth.make_flags(n, n, n, True, nb=1, geometry=True, exotic=False)
        U = synth.make_smooth_velocity(flags, True, amp=2.0)
        f = torch.from_numpy(flags).cuda(); u = torch.from_numpy(U).cuda()
        tfluids.setWallBcsForward(u, f)
        d = torch.zeros_like(f); tfluids.velocityDivergenceForward(u, f, d)
        p = torch.zeros_like(f)
        for precond, gp in (("ic0", 0), ("ic0", 4), ("none", 0)):
            c = tfluids.context()
            c.lib.tfl_debug_pcg_groups(c.h, gp)
            for rep in range(2):
                torch.cuda.synchronize(); t0 = time.time()
                res = tfluids.solveLinearSystemPCG(p, f, d, True, 1e-4, 100, precond)
                torch.cuda.synchronize(); t1 = time.time()
            it = tfluids.solveLinearSystemPCG.last_iterations
            print("n=%d %s gp=%d: %.2f ms, %d iterations (%.3f ms/iter), residual %g" % (
                n, precond, gp, (t1 - t0) * 1e3, it, (t1 - t0) * 1e3 / max(it, 1), res), flush=True)

    # pipeline picture of the LAST sweep of a solve (n = 128, default groups)
    import ctypes as C
    c = tfluids.context()
    c.lib.tfl_debug_pcg_groups(c.h, 0)
    buf = torch.zeros(64 * 4, dtype=torch.int64, device="cuda")
    c.lib.tfl_debug_pcg_timing.argtypes = [C.c_void_p, C.c_void_p]
    c.lib.tfl_debug_pcg_timing(c.h, C.c_void_p(buf.data_ptr()))
    tfluids.solveLinearSystemPCG(p, f, d, True, 1e-4, 8, "ic0")
    torch.cuda.synchronize()
    c.lib.tfl_debug_pcg_timing(c.h, None)
    t = buf.cpu().numpy().reshape(-1, 4)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    print("chunk: fwd start, fwd end, bwd start, bwd end (us)")
    for i, row in enumerate(t):
        print(i, " ".join("%8.1f" % ((v - t0) / 1e3) for v in row))
