"""Pins the C restatement (oracle/tfluids_oracle.c) against the reference's own CPU code
compiled in place (oracle/_ref): bit-exact on seeded inputs for every operator and
advection method.  Skipped where /root/reference (hence oracle/_ref) is unavailable."""
import numpy as np
import pytest

import oracle
from cases import CASES, CASE_IDS, build, bits_equal, describe_diff

METHODS = list(oracle.ADVECT_METHODS)


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
@pytest.mark.parametrize("method", METHODS)
def test_advect_scalar(orc, ref, case, method):
    c = build(case)
    U = c["U"].copy()
    orc.setWallBcsForward(U, c["flags"])
    for outside in (False, True):
        a = orc.advectScalar(0.1, c["density"], U, c["flags"], method, outside, 0.6)
        b = ref.advectScalar(0.1, c["density"], U, c["flags"], method, outside, 0.6)
        assert bits_equal(a, b), describe_diff(a, b)
    assert orc.trace_faults() == 0


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
@pytest.mark.parametrize("method", METHODS)
def test_advect_vel(orc, ref, case, method):
    c = build(case)
    U = c["U"].copy()
    orc.setWallBcsForward(U, c["flags"])
    a = orc.advectVel(0.1, U, c["flags"], method, 0.6)
    b = ref.advectVel(0.1, U, c["flags"], method, 0.6)
    assert bits_equal(a, b), describe_diff(a, b)


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_pointwise_operators(orc, ref, case):
    c = build(case)
    fl = c["flags"]
    for mask in (False, True):
        a, b = c["U"].copy(), c["U"].copy()
        orc.setWallBcsForward(a, fl, mask)
        ref.setWallBcsForward(b, fl, mask)
        assert bits_equal(a, b), "setWallBcs " + describe_diff(a, b)
    U = c["U"].copy()
    orc.setWallBcsForward(U, fl)
    a, b = orc.velocityDivergenceForward(U, fl), ref.velocityDivergenceForward(U, fl)
    assert bits_equal(a, b), "divergence " + describe_diff(a, b)
    a, b = U.copy(), U.copy()
    orc.velocityUpdateForward(a, fl, c["p"])
    ref.velocityUpdateForward(b, fl, c["p"])
    assert bits_equal(a, b), "velocityUpdate " + describe_diff(a, b)
    g = [0.1, -0.7, 0.3]
    a, b = U.copy(), U.copy()
    orc.addBuoyancy(a, fl, c["density"], g, 0.1)
    ref.addBuoyancy(b, fl, c["density"], g, 0.1)
    assert bits_equal(a, b), "addBuoyancy " + describe_diff(a, b)
    a, b = U.copy(), U.copy()
    orc.addGravity(a, fl, g, 0.1)
    ref.addGravity(b, fl, g, 0.1)
    assert bits_equal(a, b), "addGravity " + describe_diff(a, b)
    a, b = U.copy(), U.copy()
    orc.vorticityConfinement(a, fl, 0.3)
    ref.vorticityConfinement(b, fl, 0.3)
    assert bits_equal(a, b), "vorticityConfinement " + describe_diff(a, b)


def test_empty_domain_and_occupancy(orc, ref):
    for is3d, shape in ((True, (2, 1, 7, 9, 11)), (False, (2, 1, 1, 9, 11))):
        for bnd in (1, 2):
            a = np.zeros(shape, np.float32)
            b = np.zeros(shape, np.float32)
            orc.emptyDomain(a, is3d, bnd)
            ref.emptyDomain(b, is3d, bnd)
            assert bits_equal(a, b)
            assert bits_equal(orc.flagsToOccupancy(a), ref.flagsToOccupancy(b))


def test_line_trace_random(orc, ref):
    rs = np.random.RandomState(5)
    from fluidnet_b200 import synth
    flags = synth.make_flags(20, 18, 16, True, nb=1, geometry=True)
    n = 0
    for _ in range(4000):
        pos = (rs.rand(3) * [18, 16, 14] + 1).astype(np.float32)
        if (int(flags[0, 0, int(pos[2]), int(pos[1]), int(pos[0])]) & 1) == 0:
            continue
        delta = (rs.randn(3) * rs.choice([0.3, 2.0, 9.0])).astype(np.float32)
        ha, pa = orc.calcLineTrace(pos, delta, flags)
        hb, pb = ref.calcLineTrace(pos, delta, flags)
        assert ha == hb and bits_equal(pa, pb), (pos, delta, pa, pb)
        n += 1
    assert n > 2000
