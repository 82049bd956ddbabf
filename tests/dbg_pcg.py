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
        for precond in ("ic0", "none"):
            for rep in range(2):
                torch.cuda.synchronize(); t0 = time.time()
                res = tfluids.solveLinearSystemPCG(p, f, d, True, 1e-4, 100, precond)
                torch.cuda.synchronize(); t1 = time.time()
            it = tfluids.solveLinearSystemPCG.last_iterations
            print("n=%d %s: %.2f ms, %d iterations (%.3f ms/iter), residual %g" % (
                n, precond, (t1 - t0) * 1e3, it, (t1 - t0) * 1e3 / max(it, 1), res), flush=True)
