"""tfl_solve_linear_system_pcg (through the C ABI) against the CPU oracle and against the criteria of
the reference's own PCG test (tfluids/test_tfluids.lua:836-905).  Floating point with unspecified
summation order on both sides (cuBLAS dots in the reference), so: same iteration counts within +-2,
pressure within 2e-3 * max|p| of the oracle's (both are tol-converged solutions of the same
system), residual < 2 tol, max|div| after the velocity update < 1e-4."""
import numpy as np
import pytest

import pcg_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from gpu_backend import GpuBackend
    return GpuBackend()


@pytest.mark.parametrize("precond", ["none", "ilu0", "ic0"])
@pytest.mark.parametrize("is3d", [True, False], ids=["3d", "2d"])
def test_pcg_matches_oracle(orc, gpu, is3d, precond):
    flags, U, div = pcg_cases.make(orc, is3d)
    tol = 1e-5
    want = np.zeros(flags.shape, np.float32)
    got = np.random.default_rng(3).random(flags.shape).astype(np.float32)      # must be overwritten
    rw = orc.solveLinearSystemPCG(want, flags, div, is3d, tol, 1000, precond)
    rg = gpu.solveLinearSystemPCG(got, flags, div, is3d, tol, 1000, precond)
    assert rg < 2 * tol and rw < 2 * tol
    assert abs(gpu.last_pcg_iters - orc.last_pcg_iters) <= 2, (gpu.last_pcg_iters, orc.last_pcg_iters)
    assert not np.isnan(got).any()
    assert np.all(got[flags != 1] == 0)
    assert np.abs(got - want).max() <= 2e-3 * np.abs(want).max()
    U2 = U.copy()
    gpu.velocityUpdateForward(U2, flags, got)
    assert np.abs(orc.velocityDivergenceForward(U2, flags)).max() < 1e-4


def test_pcg_small_components(orc, gpu):
    """1-cell components keep p = 0, 2-4 cell components are solved without preconditioner, every
    solved component has zero mean (generic/tfluids.cu:1386-1404, 1735-1740)."""
    flags, U, div = pcg_cases.make(orc, True, nb=1)
    got = np.zeros(flags.shape, np.float32)
    gpu.solveLinearSystemPCG(got, flags, div, True, 1e-6, 1000, "ic0")
    comp, sizes = orc.findConnectedFluidComponents(flags, True, 0)
    assert (sizes == 1).any() and ((sizes > 1) & (sizes < 5)).any() and (sizes >= 5).sum() >= 3
    for ic, size in enumerate(sizes):
        vals = got[0, 0][comp == ic]
        if size == 1:
            assert vals[0] == 0
        else:
            assert abs(vals.astype(np.float64).mean()) < 1e-5
            assert np.abs(vals).max() > 0


def test_pcg_empty_cells_and_max_iter(orc, gpu):
    """Empty (Dirichlet) neighbours enter the diagonal only; `iter <= maxIter` runs maxIter + 1 iterations."""
    flags, U, div = pcg_cases.make(orc, True, nb=1, empty=True)
    want = np.zeros(flags.shape, np.float32)
    got = np.zeros(flags.shape, np.float32)
    orc.solveLinearSystemPCG(want, flags, div, True, 1e-5, 1000, "ic0")
    gpu.solveLinearSystemPCG(got, flags, div, True, 1e-5, 1000, "ic0")
    assert np.abs(got - want).max() <= 2e-3 * np.abs(want).max()
    gpu.solveLinearSystemPCG(got, flags, div, True, 1e-12, 3, "none")
    assert gpu.last_pcg_iters == 4


def test_pcg_errors(orc, gpu):
    import torch
    from fluidnet_b200 import tfluids
    from fluidnet_b200._lib import TflError
    flags, U, div = pcg_cases.make(orc, True, nb=1)
    f = torch.from_numpy(flags).cuda()
    d = torch.from_numpy(div).cuda()
    p = torch.zeros_like(f)
    with pytest.raises(TflError, match="preconType"):
        tfluids.solveLinearSystemPCG(p, f, d, True, 1e-5, 10, "jacobi")
    f2 = f.clone()
    f2[0, 0, 0, 5, 5] = 1
    with pytest.raises(TflError, match="Non fluid cell"):
        tfluids.solveLinearSystemPCG(p, f2, d, True, 1e-5, 10, "ic0")
    nothing = torch.full_like(f, 2.0)                         # no fluid at all: residual stays -inf
    assert tfluids.solveLinearSystemPCG(p, nothing, d, True) == float("-inf")
    assert float(p.abs().max()) == 0.0


def test_pcg_step_matches_oracle(orc, gpu):
    """tfluids.simulate with simMethod = 'pcg' (lib/simulate.lua:280-286), both host drivers."""
    import torch
    import oracle
    from fluidnet_b200 import simulate as fsim
    flags, U, _ = pcg_cases.make(orc, True, nb=1, pockets=False)
    rng = np.random.default_rng(5)
    dens = (rng.random(flags.shape).astype(np.float32)) * (flags == 1)
    mconf = oracle.default_mconf(simMethod="pcg", is3D=True, maxIter=100, buoyancyScale=1.0,
                                 vorticityConfinementAmp=0.0)
    ref = dict(pDiv=np.zeros(flags.shape, np.float32), UDiv=U.copy(), flags=flags.copy(), density=dens.copy())
    oracle.simulate(orc, mconf, ref)
    for fused in (False, True):
        b = {k: torch.from_numpy(v).cuda() for k, v in
             dict(pDiv=np.zeros(flags.shape, np.float32), UDiv=U.copy(), flags=flags.copy(), density=dens.copy()).items()}
        b["div"] = torch.zeros_like(b["pDiv"])
        (fsim.simulate_fused if fused else fsim.simulate)(None, mconf, b, None)
        got = b["UDiv"].cpu().numpy()
        assert np.abs(got - ref["UDiv"]).max() <= 2e-3 * np.abs(ref["UDiv"]).max()
        assert np.abs(orc.velocityDivergenceForward(got, flags)).max() < 1e-3


def test_pcg_full_size(orc, gpu):
    """BASELINE config 4 geometry: 128^3, ic0, tol 1e-4, maxIter 100 -- by property."""
    from fluidnet_b200 import synth
    import torch
    from fluidnet_b200 import tfluids
    n = 128
    flags = synth.make_flags(n, n, n, True, nb=1, geometry=True, exotic=False)
    U = synth.make_smooth_velocity(flags, True, amp=2.0)
    f = torch.from_numpy(flags).cuda()
    u = torch.from_numpy(U).cuda()
    tfluids.setWallBcsForward(u, f)
    d = torch.zeros_like(f)
    tfluids.velocityDivergenceForward(u, f, d)
    p = torch.zeros_like(f)
    res = tfluids.solveLinearSystemPCG(p, f, d, True, 1e-4, 1000, "ic0")
    assert res < 2e-4
    tfluids.velocityUpdateForward(u, f, p)
    d2 = torch.zeros_like(f)
    tfluids.velocityDivergenceForward(u, f, d2)
    assert float(d2.abs().max()) < 1e-3
    assert float(d.abs().max()) > 10 * float(d2.abs().max())
