"""Deterministic synthetic inputs for parity tests and bench.

Host-side numpy only.  Flags follow the reference's `emptyDomain(bnd=1)`
(torch/tfluids/generic/tfluids.cc:136-172) plus optional solid geometry (a sphere of
radius N/8 centred at (N/2, N/3, N/2) and a 2-cell slab) to exercise line tracing and
wall boundary conditions.
"""
import numpy as np

FLUID, OBSTACLE, EMPTY, OUTFLOW, STICK = 1.0, 2.0, 4.0, 16.0, 128.0


def make_flags(nx, ny, nz, is3d=True, nb=1, geometry=True, exotic=False, seed=7):
    """[nb][1][nz][ny][nx] float32 flags. exotic=True sprinkles Empty / Outflow / Stick
    cells (never on the border) to exercise every branch of the wall/pressure stencils."""
    if not is3d:
        assert nz == 1
    f = np.full((nb, 1, nz, ny, nx), FLUID, np.float32)
    f[..., 0] = OBSTACLE
    f[..., -1] = OBSTACLE
    f[..., 0, :] = OBSTACLE
    f[..., -1, :] = OBSTACLE
    if is3d:
        f[:, :, 0] = OBSTACLE
        f[:, :, -1] = OBSTACLE
    if geometry:
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        n = max(nx, ny, nz)
        cz = nz // 2 if is3d else 0
        r2 = (x - nx // 2) ** 2 + (y - ny // 3) ** 2 + ((z - cz) ** 2 if is3d else 0)
        sphere = r2 <= (n / 8.0) ** 2
        slab = (y >= (2 * ny) // 3) & (y < (2 * ny) // 3 + 2) & (x >= nx // 4) & (x < nx // 2)
        f[:, 0][np.broadcast_to(sphere | slab, f[:, 0].shape)] = OBSTACLE
    if exotic:
        rs = np.random.RandomState(seed)
        inner = np.zeros(f.shape, bool)
        if is3d:
            inner[:, :, 2:-2, 2:-2, 2:-2] = True
        else:
            inner[:, :, :, 2:-2, 2:-2] = True
        r = rs.rand(*f.shape)
        f[inner & (r < 0.03)] = EMPTY
        f[inner & (r >= 0.03) & (r < 0.04)] = EMPTY + OUTFLOW
        f[inner & (r >= 0.04) & (r < 0.05)] = OBSTACLE + STICK
    return np.ascontiguousarray(f)


def make_velocity(flags, is3d=True, amp=2.0, seed=1234):
    """Uniform in [-amp, amp] cells/s per face (|u| dt <= 0.35 cell at amp=2, dt=0.1)."""
    nb, _, nz, ny, nx = flags.shape
    rs = np.random.RandomState(seed)
    nc = 3 if is3d else 2
    U = (rs.rand(nb, nc, nz, ny, nx).astype(np.float32) * 2.0 - 1.0) * np.float32(amp)
    return np.ascontiguousarray(U.astype(np.float32))


def make_smooth_velocity(flags, is3d=True, amp=2.0, seed=1234):
    """Band-limited velocity (a few random Fourier modes): exercises longer coherent
    traces than white noise does."""
    nb, _, nz, ny, nx = flags.shape
    rs = np.random.RandomState(seed)
    nc = 3 if is3d else 2
    z, y, x = np.meshgrid(np.arange(nz) / max(nz, 1), np.arange(ny) / ny, np.arange(nx) / nx,
                          indexing="ij")
    U = np.zeros((nb, nc, nz, ny, nx), np.float64)
    for b in range(nb):
        for c in range(nc):
            for _ in range(4):
                k = rs.randint(1, 4, size=3)
                ph = rs.rand(3) * 2 * np.pi
                U[b, c] += rs.randn() * np.sin(2 * np.pi * k[0] * x + ph[0]) * \
                    np.sin(2 * np.pi * k[1] * y + ph[1]) * np.cos(2 * np.pi * k[2] * z + ph[2])
    U *= amp / max(np.abs(U).max(), 1e-9)
    return np.ascontiguousarray(U.astype(np.float32))


def make_density(flags, seed=1235):
    rs = np.random.RandomState(seed)
    d = rs.rand(*flags.shape).astype(np.float32)
    d[(flags.astype(np.int32) & 1) == 0] = 0.0
    return np.ascontiguousarray(d)


def make_model(is3d=True, seed=4321):
    """Random-init weights of the reference 'default' architecture
    (torch/lib/model.lua:179-186 2-D, :219-226 3-D), Torch `reset` convention
    uniform +-1/sqrt(fan_in). Inputs: pDiv, div, occupancy (lib/default_conf.lua:76-81)."""
    rs = np.random.RandomState(seed)
    if is3d:
        spec = [(3, 8, 3), (8, 8, 3), (8, 8, 3), (8, 8, 1), (8, 1, 1)]
    else:
        spec = [(3, 16, 3), (16, 16, 3), (16, 16, 3), (16, 16, 3), (16, 1, 1)]
    layers = []
    for cin, cout, k in spec:
        kz = k if is3d else 1
        fan_in = cin * kz * k * k
        bound = 1.0 / np.sqrt(fan_in)
        w = ((rs.rand(cout, cin, kz, k, k) * 2 - 1) * bound).astype(np.float32)
        b = ((rs.rand(cout) * 2 - 1) * bound).astype(np.float32)
        layers.append((np.ascontiguousarray(w), np.ascontiguousarray(b)))
    return {"is3D": is3d, "layers": layers}
