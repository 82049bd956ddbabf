"""Replays the operator set stored in a golden fixture on any backend exposing the
oracle's numpy-level API (oracle.Oracle, oracle.Reference, gpu_backend.GpuBackend)."""
import numpy as np

import oracle

G = [0.1, -0.7, 0.3]
DT = 0.1
STRENGTH = 0.6


def run_all(be, flags, U, density, p):
    """Returns {key: ndarray} for every operator / method on the given inputs."""
    out = {}
    for m in oracle.ADVECT_METHODS:
        for outside in (0, 1):
            out["advectScalar/%s/%d" % (m, outside)] = be.advectScalar(DT, density, U, flags, m,
                                                                        bool(outside), STRENGTH)
        out["advectVel/%s" % m] = be.advectVel(DT, U, flags, m, STRENGTH)
    a = U.copy(); be.setWallBcsForward(a, flags); out["setWallBcs"] = a
    a = U.copy(); be.setWallBcsForward(a, flags, True); out["setWallBcsMask"] = a
    out["div"] = be.velocityDivergenceForward(U, flags)
    a = U.copy(); be.velocityUpdateForward(a, flags, p); out["velUpdate"] = a
    a = U.copy(); be.addBuoyancy(a, flags, density, G, DT); out["buoy"] = a
    a = U.copy(); be.addGravity(a, flags, G, DT); out["grav"] = a
    a = U.copy(); be.vorticityConfinement(a, flags, 0.3); out["vort"] = a
    return out


def replay(be, z):
    flags, U, density, p = z["in_flags"], z["in_U"], z["in_density"], z["in_p"]
    got = run_all(be, flags, U, density, p)
    for key in sorted(got):
        yield key, got[key], z["out_" + key.replace("/", "__")]
