"""ORACLE (test infrastructure): ctypes binding of oracle/sdf_loss_ref.c, the CPU restatement of
omg_cuda.sdf_loss_forward (layers/sdf_matching_loss_kernel.cu:204-262, layers/omg_layers.cpp:24-49)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libomg_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "sdf_loss_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.omg_oracle_sdf_loss.restype = ctypes.c_longlong
        _lib.omg_oracle_sdf_loss.argtypes = [fp] * 8 + [ctypes.c_int, ctypes.c_int] + [fp] * 3
        _lib.omg_oracle_value_interp.argtypes = [fp, ctypes.c_int, fp] + [ctypes.c_int] * 3 + [fp]
        _lib.omg_oracle_grad_interp.argtypes = [fp, ctypes.c_int, fp] + [ctypes.c_int] * 3 + [ctypes.c_float, fp]
    return _lib


def interp(pgrid, grid, delta):
    """The restatement's getValueInterpolated / getGradientInterpolated on grid coordinates pgrid [N,3] of one grid
    [d0,d1,d2] -> (values [N], gradients [N,3])."""
    lib = _load()
    pgrid, grid = _f32(pgrid), _f32(grid)
    n = pgrid.shape[0]
    val, grad = np.empty(n, np.float32), np.empty((n, 3), np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    c = lambda a: a.ctypes.data_as(fp)
    lib.omg_oracle_value_interp(c(pgrid), n, c(grid), grid.shape[0], grid.shape[1], grid.shape[2], c(val))
    lib.omg_oracle_grad_interp(c(pgrid), n, c(grid), grid.shape[0], grid.shape[1], grid.shape[2], float(delta), c(grad))
    return val, grad


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def sdf_loss_forward(pose_init, sdf_grids, sdf_limits, points, epsilons, padding_scales, clearances,
                     disables, return_pin=False):
    """numpy in / numpy out.  Same argument order and shapes as omg_cuda.sdf_loss_forward."""
    lib = _load()
    pose_init, sdf_grids, sdf_limits, points = _f32(pose_init), _f32(sdf_grids), _f32(sdf_limits), _f32(points)
    epsilons, padding_scales, clearances, disables = (_f32(epsilons), _f32(padding_scales),
                                                      _f32(clearances), _f32(disables))
    n, o = points.shape[0], pose_init.shape[0]
    assert points.shape == (n, 3) and pose_init.shape == (o, 4, 4) and sdf_limits.shape == (o, 10)
    assert sdf_grids.shape[0] == o
    for i in range(o):  # the kernel indexes every object's grid with its own (d0,d1,d2) from limits
        assert tuple(int(v) for v in sdf_limits[i, 6:9]) == tuple(sdf_grids.shape[1:])
    pot = np.empty(n, np.float32)
    grad = np.empty((n, 3), np.float32)
    col = np.empty(n, np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    c = lambda a: a.ctypes.data_as(fp)
    pin = lib.omg_oracle_sdf_loss(c(pose_init), c(sdf_grids), c(sdf_limits), c(points), c(epsilons),
                                  c(padding_scales), c(clearances), c(disables), n, o, c(pot), c(grad), c(col))
    if return_pin:
        return pot, grad, col, int(pin)
    return pot, grad, col


def sdf_loss_forward_torch(pose_init, sdf_grids, sdf_limits, points, epsilons, padding_scales, clearances,
                           disables):
    """torch CPU tensors in / out: what tools/ref_harness.py binds as omg_cuda.sdf_loss_forward."""
    import torch

    pot, grad, col = sdf_loss_forward(*[t.detach().cpu().numpy() for t in (
        pose_init, sdf_grids, sdf_limits, points, epsilons, padding_scales, clearances, disables)])
    return [torch.from_numpy(pot), torch.from_numpy(grad), torch.from_numpy(col)]
