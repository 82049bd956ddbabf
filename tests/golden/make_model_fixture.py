"""Writes tests/golden/myModel2D_layers.npz: the convolution weights of the trained 2-D model the
reference ships (/root/reference/data/models/myModel2D, Torch7 binary), extracted with
fluidnet_b200/torch7.py, plus the few mconf keys the projection reads.  Run here (needs /root/reference);
the fixture travels to the GPU box so the parity tests can use real trained weights."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from fluidnet_b200 import torch7  # noqa: E402

SRC = "/root/reference/data/models/myModel2D"


def main():
    ref = torch7.load_reference_model(SRC)
    out = {"is3D": np.array(ref["is3D"]), "n_layers": np.array(len(ref["layers"])),
           "normalizeInputThreshold": np.array(float(ref["mconf"]["normalizeInputThreshold"]))}
    for i, (w, b) in enumerate(ref["layers"]):
        out["w%d" % i] = w
        out["b%d" % i] = b
    np.savez_compressed(os.path.join(HERE, "myModel2D_layers.npz"), **out)
    print("layers:", [w.shape for w, _ in ref["layers"]])


if __name__ == "__main__":
    main()
