"""Generates the golden fixtures in this directory from the reference's OWN CPU code
(oracle/_ref/libtfluids_ref.so, compiled in place from /root/reference -- see
oracle/Makefile).  Run here (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each .npz holds seeded inputs (in_*) and the reference's float32 outputs (out_*) for every
operator and advection method of the step."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from fluidnet_b200 import synth  # noqa: E402
from golden_util import run_all  # noqa: E402

SPECS = [
    ("ref_3d_geom", True, (14, 12, 10), True, False, 4.0),
    ("ref_3d_exotic", True, (12, 12, 10), True, True, 12.0),
    ("ref_2d_geom", False, (22, 18, 1), True, False, 4.0),
    ("ref_2d_exotic", False, (20, 18, 1), True, True, 12.0),
]


def main():
    ref = oracle.Reference()
    for name, is3d, (nx, ny, nz), geom, exotic, amp in SPECS:
        flags = synth.make_flags(nx, ny, nz, is3d, nb=1, geometry=geom, exotic=exotic, seed=21)
        U = synth.make_velocity(flags, is3d, amp=amp, seed=31)
        ref.setWallBcsForward(U, flags)
        density = synth.make_density(flags, seed=41)
        p = np.ascontiguousarray(synth.make_density(flags, seed=51) - np.float32(0.5))
        outs = run_all(ref, flags, U, density, p)
        data = {"in_flags": flags, "in_U": U, "in_density": density, "in_p": p}
        for k, v in outs.items():
            data["out_" + k.replace("/", "__")] = v
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
