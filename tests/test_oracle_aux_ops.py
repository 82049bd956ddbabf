"""normalizePressureMean, volumetricUpSamplingNearestForward, rectangularBlur, signedDistanceField:
the oracle's restatements against the reference's own CPU code (oracle/_ref), bit for bit; the mean
removal to float rounding (the reference accumulates it with an OpenMP atomic, order unspecified)."""
import numpy as np
import pytest

from oracle import api
from cases import bits_equal, describe_diff
import pcg_cases


@pytest.fixture(scope="module")
def ref():
    if not api.have_reference():
        pytest.skip("oracle/_ref not built")
    return api.Reference()


@pytest.fixture(scope="module")
def orc():
    return api.Oracle()


def fields(orc, is3d, seed=0):
    flags, _, _ = pcg_cases.make(orc, is3d, nb=2, seed=seed)
    x = np.random.default_rng(seed).standard_normal((2, 3) + flags.shape[2:]).astype(np.float32)
    return flags, x


@pytest.mark.parametrize("is3d", [True, False])
def test_upsampling(orc, ref, is3d):
    _, x = fields(orc, is3d)
    x = np.ascontiguousarray(x[:, :, :5, :6, :7])
    for ratio in (1, 2, 3):
        a, b = orc.volumetricUpSamplingNearestForward(ratio, x), ref.volumetricUpSamplingNearestForward(ratio, x)
        assert bits_equal(a, b)
        assert a.shape[2:] == tuple(s * ratio for s in x.shape[2:])


@pytest.mark.parametrize("is3d", [True, False])
def test_rectangular_blur(orc, ref, is3d):
    _, x = fields(orc, is3d)
    for rad in (1, 2, 5, 40):               # 40 > every extent: the clamped-edge branches
        a, b = orc.rectangularBlur(x, rad, is3d), ref.rectangularBlur(x, rad, is3d)
        assert bits_equal(a, b), describe_diff(a, b)
    const = np.full_like(x, 3.0)
    assert np.allclose(orc.rectangularBlur(const, 3, is3d), 3.0, atol=1e-5)


@pytest.mark.parametrize("is3d", [True, False])
def test_signed_distance_field(orc, ref, is3d):
    flags, _ = fields(orc, is3d)
    for rad in (1, 3):
        a, b = orc.signedDistanceField(flags, rad, is3d), ref.signedDistanceField(flags, rad, is3d)
        assert bits_equal(a, b)
        assert np.all(a[flags == 2] == 0) and a.max() <= rad


@pytest.mark.parametrize("is3d", [True, False])
def test_normalize_pressure_mean(orc, ref, is3d):
    flags, x = fields(orc, is3d)
    flags[0, 0, 0, 3, 3] = 1                 # a fluid cell on the border is legal here
    p1 = np.ascontiguousarray(x[:, :1]).copy()
    p2 = p1.copy()
    orc.normalizePressureMean(p1, flags, is3d)
    ref.normalizePressureMean(p2, flags, is3d)
    assert np.abs(p1 - p2).max() <= 2e-6 * np.abs(p2).max()
    assert bits_equal(p1[flags != 1], x[:, :1][flags != 1])          # non-fluid cells untouched
    comp, sizes = orc.findConnectedFluidComponents(flags, is3d, 0)
    for ic in range(len(sizes)):
        assert abs(p1[0, 0][comp == ic].astype(np.float64).mean()) < 1e-5
