"""Seeded PCG inputs shared by the oracle (CPU) and libtfl (GPU) tests: flags with several connected
fluid components (pockets of 1, 2-4 and >= 5 cells), a wall-conditioned random velocity, its
divergence as the right-hand side."""
import numpy as np


def make(oracle, is3d, nb=2, seed=0, n=(18, 16, 14), pockets=True, empty=False):
    rng = np.random.default_rng(seed)
    nx, ny, nz = n if is3d else (n[0] * 2, n[1] * 2, 1)
    shape = (nb, 1, nz, ny, nx)
    flags = np.ones(shape, np.float32)
    oracle.emptyDomain(flags, is3d, 1)
    for b in range(nb):
        f = flags[b, 0]
        # a wall across the domain (two big components) with a few random obstacle cells
        f[:, ny // 2, :] = 2
        m = rng.random(f.shape) < 0.04
        f[m] = 2
        if pockets:
            zc = nz // 2
            def box(z0, z1, y0, y1, x0, x1):                     # obstacle shell around a fluid pocket
                zs = slice(max(z0 - 1, 0), z1 + 1) if is3d else slice(0, 1)
                f[zs, y0 - 1:y1 + 1, x0 - 1:x1 + 1] = 2
                f[(slice(z0, z1) if is3d else slice(0, 1)), y0:y1, x0:x1] = 1
            z0 = zc if is3d else 0
            box(z0, z0 + 1, 3, 4, 3, 4)                           # 1 cell: skipped
            box(z0, z0 + 1, 3, 4, 7, 10)                          # 3 cells: no preconditioner
            box(z0, z0 + (2 if is3d else 1), 3, 6, 12, 15)        # >= 9 cells: preconditioned
        if empty:
            f[(slice(1, nz - 1) if is3d else slice(0, 1)), ny - 3, 2:nx - 2] = 4
    tmp = flags.copy()
    oracle.emptyDomain(tmp, is3d, 1)
    flags[tmp == 2] = 2
    U = rng.standard_normal((nb, 3 if is3d else 2, nz, ny, nx)).astype(np.float32)
    oracle.setWallBcsForward(U, flags)
    div = oracle.velocityDivergenceForward(U, flags)
    return flags, U, div
