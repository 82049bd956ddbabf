"""The oracle's restatement of the single-bank model graphs of torch/lib/model.lua:164-340 ('default', 'tog',
'yang': convolution, ConvolutionUpsample pixel shuffle, non-linearity, pooling) against an independent
evaluation with torch.nn.functional in float64 -- the conv stack has no CPU source in the reference
(cudnn.torch), so this pins the graph semantics the GPU path is then compared with."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from fluidnet_b200 import synth


def torch_stack(x, model):
    is3d = model["is3D"]
    nl = len(model["layers"])
    pool = model.get("pool") or [1] * nl
    up = model.get("up") or [1] * nl
    t = torch.from_numpy(x).double()
    if not is3d:
        t = t[:, :, 0]
    for li, (w, b) in enumerate(model["layers"]):
        wt, bt = torch.from_numpy(w).double(), torch.from_numpy(b).double()
        k = w.shape[-1]
        if is3d:
            t = F.conv3d(t, wt, bt, padding=k // 2)
        else:
            t = F.conv2d(t, wt[:, :, 0], bt, padding=k // 2)
        s = up[li]
        if s > 1:
            if is3d:                      # explicit index form of volumetric_convolution_upsample.lua
                bsz, ct, d, h, w_ = t.shape
                no = ct // s ** 3
                out = torch.empty(bsz, no, d * s, h * s, w_ * s, dtype=t.dtype)
                for st in range(s):
                    for sh in range(s):
                        for sw in range(s):
                            ch = torch.arange(no) * s ** 3 + (st * s + sh) * s + sw
                            out[:, :, st::s, sh::s, sw::s] = t[:, ch]
                t = out
            else:
                t = F.pixel_shuffle(t, s)
        if li < nl - 1:
            t = torch.sigmoid(t) if model.get("nonlinType") == "sigmoid" else F.relu(t)
        if pool[li] > 1:
            fn = {(True, "avg"): F.avg_pool3d, (True, "max"): F.max_pool3d, (False, "avg"): F.avg_pool2d,
                  (False, "max"): F.max_pool2d}[(is3d, model.get("poolType", "avg"))]
            t = fn(t, pool[li])
    if not is3d:
        t = t[:, :, None]
    return t.numpy()


@pytest.mark.parametrize("model_type", ["default", "tog", "yang"])
@pytest.mark.parametrize("is3d", [True, False])
def test_graph_matches_torch(is3d, model_type):
    orc = oracle.Oracle()
    n = 8 if is3d else 16
    flags = synth.make_flags(n, n, n if is3d else 1, is3d, nb=2, geometry=False)
    U = synth.make_smooth_velocity(flags, is3d, amp=1.0)
    orc.setWallBcsForward(U, flags)
    model = synth.make_model(is3d, model_type=model_type)
    if model_type == "tog":
        model["poolType"] = "max" if is3d else "avg"                   # cover both pooling types
    p0 = (synth.make_density(flags, seed=5) - np.float32(0.5)) * np.float32(0.1)
    p, U2, scale = oracle.model_forward(orc, model, p0, U, flags)
    # rebuild the network input exactly as model_forward does and push it through torch
    U1 = U.copy()
    orc.setWallBcsForward(U1, flags, as_mask_multiply=True)
    sc = scale.reshape(-1, 1, 1, 1, 1)
    x = np.concatenate([(p0 / sc).astype(np.float32), (orc.velocityDivergenceForward(U1, flags) / sc).astype(np.float32),
                        orc.flagsToOccupancy(flags)], axis=1)
    want = torch_stack(np.ascontiguousarray(x), model) * sc
    assert p.shape == flags.shape
    assert np.abs(p - want).max() <= 2e-6 * max(np.abs(want).max(), 1e-3)
