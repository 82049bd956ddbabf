"""fluidnet_b200/formats.py: Manta `.bin` frames (torch/lib/load_manta_file.lua) and `.vbox` volumes
(torch/fluid_net_3d_sim.lua) -- byte layouts checked against hand-built files."""
import struct

import numpy as np
import pytest

from fluidnet_b200 import formats


@pytest.mark.parametrize("is3d", [True, False])
def test_manta_bin_layout(tmp_path, is3d):
    nx, ny, nz = 5, 4, (3 if is3d else 1)
    n = nx * ny * nz
    rng = np.random.default_rng(0)
    fields = [rng.standard_normal(n).astype(np.float32) for _ in range(3 if is3d else 2)]
    p = rng.standard_normal(n).astype(np.float32)
    flags = rng.integers(1, 3, n).astype(np.int32)
    dens = rng.random(n).astype(np.float32)
    raw = struct.pack("<5i", 0, nx, ny, nz, 1 if is3d else 0) + b"".join(a.tobytes() for a in fields) + \
        p.tobytes() + flags.tobytes() + dens.tobytes()
    path = tmp_path / "frame.bin"
    path.write_bytes(raw)
    gp, gU, gf, gd, g3 = formats.load_manta_file(str(path))
    assert g3 is is3d and gU.shape == (1, len(fields), nz, ny, nx) and gf.dtype == np.float32
    assert gp[0, 0, nz - 1, 2, 3] == p[((nz - 1) * ny + 2) * nx + 3]          # x fastest
    for c, a in enumerate(fields):
        assert np.array_equal(gU[0, c].ravel(), a)
    assert np.array_equal(gf.ravel(), flags.astype(np.float32)) and np.array_equal(gd.ravel(), dens)
    out = tmp_path / "again.bin"
    formats.save_manta_file(str(out), gp, gU, gf, gd)
    assert out.read_bytes() == raw
    path.write_bytes(raw[:-8])
    with pytest.raises(ValueError, match="truncated"):
        formats.load_manta_file(str(path))


def test_vbox_layout(tmp_path):
    res = 4
    frames = [np.arange(res ** 3, dtype=np.float32).reshape(res, res, res) + 100 * f for f in range(2)]
    path = tmp_path / "d.vbox"
    with formats.VboxWriter(str(path), res, 2) as w:
        w.write(frames[0])
        w.write(np.stack([frames[1], frames[1]])[None])          # [1][c][z][y][x]: channel mean
    raw = path.read_bytes()
    assert struct.unpack("<4i", raw[:16]) == (res, res, res, 2)
    body = np.frombuffer(raw[16:], np.float32).reshape(2, res, res, res)     # [frame][x][y][z]
    assert body[0, 1, 2, 3] == frames[0][3, 2, 1]
    assert np.array_equal(formats.load_vbox(str(path)), np.stack(frames))
