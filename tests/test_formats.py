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


def test_binvox_reader(tmp_path):
    """Run-length volume with the reference parser's quirks (obstacles_import_binvox.lua:62-120)."""
    dims = (3, 4, 2)
    vol = np.zeros(24, np.uint8)
    vol[5:9] = 1
    vol[15:20] = 1
    vol[20] = 1          # quirk: the 4th run writes count + 1 cells and the (dropped) last run never overwrites it
    runs = [(0, 5), (1, 4), (0, 6), (1, 5), (0, 4)]                  # covers all 24 voxels; the last run is all zeros
    raw = b"#binvox 1\ndim 3 4 2\ntranslate 0.5 -1 2\nscale 1.25\ndata\n" + bytes(b for r in runs for b in r)
    path = tmp_path / "o.binvox"
    path.write_bytes(raw)
    out = formats.load_binvox(str(path))
    assert out["dims"] == [3, 4, 2] and out["translation"] == [0.5, -1.0, 2.0] and out["scale"] == 1.25
    assert out["data"].shape == (3, 2, 4) and out["data"].dtype == np.float32
    assert np.array_equal(out["data"], vol.reshape(dims).transpose(0, 2, 1))
    # the reference drops the final run: make it non-empty and it is not applied
    runs2 = [(0, 5), (1, 4), (0, 6), (1, 5), (1, 4)]
    path.write_bytes(raw[:raw.index(b"data\n") + 5] + bytes(b for r in runs2 for b in r))
    assert np.array_equal(formats.load_binvox(str(path))["data"], out["data"])
