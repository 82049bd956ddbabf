"""The operators of tfluids/init.lua around the step, through the C ABI, against the oracle (which is
pinned bit-exactly on the reference's CPU code, tests/test_oracle_aux_ops.py)."""
import numpy as np
import pytest
import torch

from cases import bits_equal, describe_diff
import pcg_cases

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def fields(orc, is3d, seed=0):
    flags, _, _ = pcg_cases.make(orc, is3d, nb=2, seed=seed)
    x = np.random.default_rng(seed).standard_normal((2, 3) + flags.shape[2:]).astype(np.float32)
    return flags, x


@pytest.mark.parametrize("is3d", [True, False], ids=["3d", "2d"])
def test_upsampling(orc, is3d):
    from fluidnet_b200 import tfluids
    _, x = fields(orc, is3d)
    x = np.ascontiguousarray(x[:, :, :5, :6, :7])
    for ratio in (1, 2, 3):
        want = orc.volumetricUpSamplingNearestForward(ratio, x)
        out = torch.full(want.shape, 7.0, device="cuda")
        tfluids.volumetricUpSamplingNearestForward(ratio, dev(x), out)
        assert bits_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("is3d", [True, False], ids=["3d", "2d"])
def test_rectangular_blur(orc, is3d):
    from fluidnet_b200 import tfluids
    _, x = fields(orc, is3d)
    for rad in (1, 2, 5, 40):
        want = orc.rectangularBlur(x, rad, is3d)
        src = dev(x)
        out = torch.full_like(src, 7.0)
        tfluids.rectangularBlur(src, rad, is3d, out)
        got = out.cpu().numpy()
        assert bits_equal(got, want), describe_diff(got, want)
        assert torch.equal(src, dev(x)), "rectangularBlur modified its input"


@pytest.mark.parametrize("is3d", [True, False], ids=["3d", "2d"])
def test_signed_distance_field(orc, is3d):
    from fluidnet_b200 import tfluids
    flags, _ = fields(orc, is3d)
    for rad in (1, 3):
        want = orc.signedDistanceField(flags, rad, is3d)
        out = torch.full(flags.shape, 7.0, device="cuda")
        tfluids.signedDistanceField(dev(flags), rad, is3d, out)
        assert bits_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("is3d", [True, False], ids=["3d", "2d"])
def test_normalize_pressure_mean(orc, is3d):
    from fluidnet_b200 import tfluids
    flags, x = fields(orc, is3d)
    flags[0, 0, 0, 3, 3] = 1                 # fluid on the border: legal for this operator
    rng = np.random.default_rng(1)
    flags[(rng.random(flags.shape) < 0.2) & (flags == 1)] = 2       # plus many small components
    p = np.ascontiguousarray(x[:, :1])
    want = p.copy()
    orc.normalizePressureMean(want, flags, is3d)
    t = dev(p)
    tfluids.normalizePressureMean(t, dev(flags), is3d)
    got = t.cpu().numpy()
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
    assert bits_equal(got[flags != 1], p[flags != 1])


def test_argument_errors(orc):
    from fluidnet_b200 import tfluids
    from fluidnet_b200._lib import TflError
    flags, x = fields(orc, True)
    f, s = dev(flags), dev(x)
    with pytest.raises(TflError, match="size mismatch"):
        tfluids.volumetricUpSamplingNearestForward(2, s, torch.zeros_like(s))
    with pytest.raises(TflError, match="alias"):
        tfluids.rectangularBlur(s, 2, True, s)
    with pytest.raises(AssertionError, match="blurRad"):
        tfluids.rectangularBlur(s, 0, True, torch.zeros_like(s))
    with pytest.raises(AssertionError, match="searchRad"):
        tfluids.signedDistanceField(f, 0, True, torch.zeros_like(f))


@pytest.mark.parametrize("is3d", [True, False], ids=["3d", "2d"])
def test_backward_operators(orc, is3d):
    from fluidnet_b200 import synth, tfluids
    rng = np.random.default_rng(3)
    nx, ny, nz = (14, 12, 10) if is3d else (22, 18, 1)
    fl = synth.make_flags(nx, ny, nz, is3d, nb=2, geometry=True, exotic=True)
    U = synth.make_velocity(fl, is3d, amp=1.0)
    go = rng.standard_normal(fl.shape).astype(np.float32)
    want = orc.velocityDivergenceBackward(U, fl, go)
    gU = torch.full(U.shape, 7.0, device="cuda")
    tfluids.velocityDivergenceBackward(dev(U), dev(fl), dev(go), gU)
    assert bits_equal(gU.cpu().numpy(), want), describe_diff(gU.cpu().numpy(), want)
    goU = rng.standard_normal(U.shape).astype(np.float32)
    p = rng.standard_normal(fl.shape).astype(np.float32)
    want = orc.velocityUpdateBackward(U, fl, p, goU)
    gp = torch.full(fl.shape, 7.0, device="cuda")
    tfluids.velocityUpdateBackward(dev(U), dev(fl), dev(p), dev(goU), gp)
    assert bits_equal(gp.cpu().numpy(), want), describe_diff(gp.cpu().numpy(), want)    # same fixed order as the oracle
    x = rng.standard_normal((2, 3, 4, 5, 6)).astype(np.float32)
    for ratio in (1, 2, 3):
        g = rng.standard_normal((2, 3, 4 * ratio, 5 * ratio, 6 * ratio)).astype(np.float32)
        gi = torch.full(x.shape, 7.0, device="cuda")
        tfluids.volumetricUpSamplingNearestBackward(ratio, dev(x), dev(g), gi)
        assert bits_equal(gi.cpu().numpy(), orc.volumetricUpSamplingNearestBackward(ratio, x, g))
