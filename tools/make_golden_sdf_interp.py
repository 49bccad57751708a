#!/usr/bin/env python
"""Writes tests/golden/sdf_interp.npz: outputs of the REFERENCE'S OWN getValueInterpolated / getGradientInterpolated
(layers/sdf_matching_loss_kernel.cu:37-86, compiled from the reference source into oracle/_ref/libsdf_ref.so by
oracle/sdf_ref/Makefile) on seeded grid coordinates that cover the interior, the (-0.5, 0.5) truncation band, the
border voxels, half-integer coordinates and out-of-bounds taps.  Grid dims are powers of two so that a GPU test can
feed the coordinates through the operator's (x - min)/(max' - min)*dim map exactly."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sdf_ref_lib  # noqa: E402


def coordinates(dims, n, rng):
    d = np.asarray(dims, np.float64)
    parts = [
        rng.uniform(-2.0, d + 2.0, (n // 4, 3)),                      # everywhere incl. out of bounds
        rng.uniform(1.5, d - 1.5, (n // 4, 3)),                       # interior (all 7 samples in bounds)
        rng.uniform(-0.75, 0.75, (n // 8, 3)) + rng.randint(0, 2, (n // 8, 3)) * (d - 1.0),   # truncation band / far border
        rng.uniform(-0.5, 2.5, (n // 8, 3)),                          # near border: gradient samples mix in 1.0
        d - rng.uniform(-0.5, 2.5, (n // 8, 3)),
    ]
    k = n - sum(p.shape[0] for p in parts)
    half = rng.randint(-1, int(d.max()) + 2, (k, 3)) + 0.5 * rng.randint(0, 2, (k, 3))   # exact (half-)integers
    parts.append(np.minimum(half, d + 1.0))
    return np.concatenate(parts).astype(np.float32)


def main():
    assert sdf_ref_lib.build_ref(), "needs /root/reference"
    rng = np.random.RandomState(20261017)
    dims = (16, 8, 32)
    grid = rng.uniform(-0.3, 0.6, dims).astype(np.float32)
    delta = np.float32(0.0123)
    pg = coordinates(dims, 20000, rng)
    val, grad = sdf_ref_lib.interp(pg, grid, delta)
    out = os.path.join(ROOT, "tests", "golden", "sdf_interp.npz")
    np.savez_compressed(out, grid=grid, delta=delta, pgrid=pg, value=val, grad=grad)
    print("wrote", out, os.path.getsize(out), "bytes;", int((val == 1.0).sum()), "out-of-bounds values")


if __name__ == "__main__":
    main()
