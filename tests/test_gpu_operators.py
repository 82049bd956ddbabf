"""GPU parity tests proper: every operator of libtfl.so, called through the C ABI (via the
fluidnet_b200.tfluids mirror), against the CPU oracle on the same seeded inputs.
Bit-exact for everything that is integer / flag logic AND for the float stencils (the
kernels are built without FMA contraction precisely so this holds); reductions (Jacobi
residual) carry a stated tolerance."""
import os

import numpy as np
import pytest

import oracle
from cases import CASES, CASE_IDS, build, bits_equal, describe_diff

pytestmark = pytest.mark.gpu
METHODS = list(oracle.ADVECT_METHODS)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gpu():
    from gpu_backend import GpuBackend
    return GpuBackend()


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
@pytest.mark.parametrize("method", METHODS)
def test_advect_scalar(orc, gpu, case, method):
    c = build(case)
    U = c["U"].copy()
    orc.setWallBcsForward(U, c["flags"])
    for outside in (False, True):
        want = orc.advectScalar(0.1, c["density"], U, c["flags"], method, outside, 0.6)
        got = gpu.advectScalar(0.1, c["density"], U, c["flags"], method, outside, 0.6)
        assert bits_equal(got, want), describe_diff(got, want)
    got = gpu.advectScalar(0.1, c["density"], U, c["flags"], method, False, 0.6, in_place=True)
    want = orc.advectScalar(0.1, c["density"], U, c["flags"], method, False, 0.6)
    assert bits_equal(got, want), "in place: " + describe_diff(got, want)
    assert gpu.trace_faults() == 0


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
@pytest.mark.parametrize("method", METHODS)
def test_advect_vel(orc, gpu, case, method):
    c = build(case)
    U = c["U"].copy()
    orc.setWallBcsForward(U, c["flags"])
    want = orc.advectVel(0.1, U, c["flags"], method, 0.6)
    got = gpu.advectVel(0.1, U, c["flags"], method, 0.6)
    assert bits_equal(got, want), describe_diff(got, want)
    got = gpu.advectVel(0.1, U, c["flags"], method, 0.6, in_place=True)
    assert bits_equal(got, want), "in place: " + describe_diff(got, want)
    assert gpu.trace_faults() == 0


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_pointwise_operators(orc, gpu, case):
    c = build(case)
    fl = c["flags"]
    for mask in (False, True):
        a, b = c["U"].copy(), c["U"].copy()
        gpu.setWallBcsForward(a, fl, mask)
        orc.setWallBcsForward(b, fl, mask)
        assert bits_equal(a, b), "setWallBcs " + describe_diff(a, b)
    U = c["U"].copy()
    orc.setWallBcsForward(U, fl)
    a, b = gpu.velocityDivergenceForward(U, fl), orc.velocityDivergenceForward(U, fl)
    assert bits_equal(a, b), "divergence " + describe_diff(a, b)
    a, b = U.copy(), U.copy()
    gpu.velocityUpdateForward(a, fl, c["p"])
    orc.velocityUpdateForward(b, fl, c["p"])
    assert bits_equal(a, b), "velocityUpdate " + describe_diff(a, b)
    g = [0.1, -0.7, 0.3]
    a, b = U.copy(), U.copy()
    gpu.addBuoyancy(a, fl, c["density"], g, 0.1)
    orc.addBuoyancy(b, fl, c["density"], g, 0.1)
    assert bits_equal(a, b), "addBuoyancy " + describe_diff(a, b)
    a, b = U.copy(), U.copy()
    gpu.addGravity(a, fl, g, 0.1)
    orc.addGravity(b, fl, g, 0.1)
    assert bits_equal(a, b), "addGravity " + describe_diff(a, b)
    a, b = U.copy(), U.copy()
    gpu.vorticityConfinement(a, fl, 0.3)
    orc.vorticityConfinement(b, fl, 0.3)
    assert bits_equal(a, b), "vorticityConfinement " + describe_diff(a, b)
    x, inv, bc = c["density"].copy(), (c["p"] > 0).astype(np.float32), c["p"].copy()
    a, b = x.copy(), x.copy()
    gpu.applyBC(a, inv, bc)
    orc.applyBC(b, inv, bc)
    assert bits_equal(a, b), "applyBC " + describe_diff(a, b)
    a, b = (U * 1e6).astype(np.float32), (U * 1e6).astype(np.float32)
    gpu.clamp(a, -1e6, 1e6)
    orc.clamp(b, -1e6, 1e6)
    assert bits_equal(a, b), "clamp " + describe_diff(a, b)


@pytest.mark.parametrize("shape,is3d", [((32, 12, 10), True), ((128, 6, 5), True), ((16, 20, 1), False),
                                        ((160, 8, 1), False)],
                         ids=["32x12x10", "128x6x5", "16x20-2d", "160x8-2d"])
def test_vorticity_quad_kernels(orc, gpu, shape, is3d):
    """Row lengths that select the 4-voxels-per-thread curl / force kernels (nx % 4 == 0 and a block shape
    that tiles): bit-exact against the oracle, geometry and exotic flags included."""
    from fluidnet_b200 import synth
    nx, ny, nz = shape
    fl = synth.make_flags(nx, ny, nz, is3d, nb=2, geometry=True, exotic=True)
    U = synth.make_velocity(fl, is3d, amp=3.0)
    orc.setWallBcsForward(U, fl)
    a, b = U.copy(), U.copy()
    gpu.vorticityConfinement(a, fl, 0.3)
    orc.vorticityConfinement(b, fl, 0.3)
    assert bits_equal(a, b), "vorticityConfinement (quad) " + describe_diff(a, b)


def test_empty_domain_and_occupancy(orc, gpu):
    for is3d, shape in ((True, (2, 1, 7, 9, 11)), (False, (2, 1, 1, 9, 11))):
        for bnd in (1, 2):
            a = np.zeros(shape, np.float32)
            b = np.zeros(shape, np.float32)
            gpu.emptyDomain(a, is3d, bnd)
            orc.emptyDomain(b, is3d, bnd)
            assert bits_equal(a, b)
            assert bits_equal(gpu.flagsToOccupancy(a), orc.flagsToOccupancy(b))
    a[0, 0, 0, 3, 3] = 4.0
    from fluidnet_b200._lib import TflError
    with pytest.raises(TflError):
        gpu.flagsToOccupancy(a)


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_jacobi(orc, gpu, case):
    c = build(case)
    fl = c["flags"]
    U = c["U"].copy()
    orc.setWallBcsForward(U, fl)
    div = orc.velocityDivergenceForward(U, fl)
    for ptol, iters in ((0.0, 1), (0.0, 7), (0.0, 40), (1e-2, 500)):
        a = np.full_like(div, 9.0)
        b = np.full_like(div, 9.0)
        ra = gpu.solveLinearSystemJacobi(a, fl, div, c["is3d"], ptol, iters)
        rb = orc.solveLinearSystemJacobi(b, fl, div, c["is3d"], ptol, iters)
        # pressure: bit-exact (same per-cell expression); residual: a reduction, 1e-5 rel.
        assert bits_equal(a, b), "jacobi p (%g,%d) %s" % (ptol, iters, describe_diff(a, b))
        assert abs(ra - rb) <= 1e-5 * max(abs(rb), 1e-12) + 1e-12, (ra, rb)
        assert gpu.last_jacobi_iters == orc.last_jacobi_iters


@pytest.mark.parametrize("shape", [(128, 16, 12), (256, 8, 9)], ids=["128x16x12", "256x8x9"])
def test_jacobi_marching_kernel(orc, gpu, shape):
    """Shapes that select the 2.5-D z-marching Jacobi kernel (nx % 128 == 0, ny % 8 == 0)."""
    from fluidnet_b200 import synth
    nx, ny, nz = shape
    fl = synth.make_flags(nx, ny, nz, True, nb=2, geometry=True, exotic=True)
    U = synth.make_velocity(fl, True, amp=2.0)
    orc.setWallBcsForward(U, fl)
    div = orc.velocityDivergenceForward(U, fl)
    for iters in (1, 2, 9, 30):
        a = np.full_like(div, 9.0)
        b = np.full_like(div, 9.0)
        ra = gpu.solveLinearSystemJacobi(a, fl, div, True, 0.0, iters)
        rb = orc.solveLinearSystemJacobi(b, fl, div, True, 0.0, iters)
        assert bits_equal(a, b), "jacobi march (%d iters) %s" % (iters, describe_diff(a, b))
        assert abs(ra - rb) <= 1e-5 * max(abs(rb), 1e-12) + 1e-12


@pytest.mark.parametrize("fname", sorted(f for f in os.listdir(GOLD) if f.startswith("ref_") and f.endswith(".npz")))
def test_gpu_matches_reference_golden(gpu, fname):
    """Against outputs of the reference's own CPU code (fixtures committed under
    tests/golden/, generated by tests/golden/make_golden.py)."""
    from golden_util import replay
    z = np.load(os.path.join(GOLD, fname))
    for key, got, want in replay(gpu, z):
        assert bits_equal(got, want), "%s %s: %s" % (fname, key, describe_diff(got, want))


def test_argument_errors(gpu):
    """Error behaviour mirrors the reference's asserts / luaL_error."""
    import torch
    from fluidnet_b200 import tfluids
    from fluidnet_b200._lib import TflError
    flags = torch.ones(1, 1, 4, 5, 6, device="cuda")
    U = torch.zeros(1, 3, 4, 5, 6, device="cuda")
    with pytest.raises(TflError):
        tfluids.advectVel(0.1, U, flags, "notAMethod")
    with pytest.raises(AssertionError):
        tfluids.advectVel(0.1, torch.zeros(1, 3, 4, 5, 7, device="cuda"), flags)
    with pytest.raises(AssertionError):
        tfluids.advectVel(0.1, torch.zeros(1, 2, 4, 5, 6, device="cuda"), flags)   # 2D U, depth 4
    with pytest.raises(TflError):
        tfluids.advectVel(0.1, U.cpu(), flags.cpu())
    with pytest.raises(TflError):
        tfluids.solveLinearSystemJacobi(flags.clone(), flags, flags.clone(), True, 0, 0)
