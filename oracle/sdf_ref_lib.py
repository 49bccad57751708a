"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_ref/libsdf_ref.so: the reference's OWN CUDA operator source
(layers/sdf_matching_loss_kernel.cu) compiled where it lies under /root/reference by oracle/sdf_ref/Makefile (build
container only; the .so travels to the GPU box).  Only tests/ may import this.

  value_interp / grad_interp   the reference's getValueInterpolated / getGradientInterpolated (kernel.cu:37-86),
                               host-compiled: run on any CPU;
  forward_device               the reference's sdf_loss_cuda_forward (kernel.cu:204-262) with its own kernels on the
                               current CUDA device (torch tensors in / out; needs a GPU)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsdf_ref.so")
_lib = None
_fp = ctypes.POINTER(ctypes.c_float)
_vp = ctypes.c_void_p


def build_ref(reference_root="/root/reference"):
    """Compile the reference's operator source (only where /root/reference exists).  Returns the path or None."""
    if not os.path.isfile(os.path.join(reference_root, "layers", "sdf_matching_loss_kernel.cu")):
        return _SO if os.path.exists(_SO) else None
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "sdf_ref"), "REF=" + reference_root])
    return _SO


def have_ref():
    return os.path.exists(_SO)


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_SO)
        _lib.sdfref_value_interp.argtypes = [_fp, ctypes.c_int, _fp] + [ctypes.c_int] * 3 + [_fp]
        _lib.sdfref_grad_interp.argtypes = [_fp, ctypes.c_int, _fp] + [ctypes.c_int] * 3 + [ctypes.c_float, _fp]
        _lib.sdfref_forward_device.argtypes = [_vp] * 8 + [ctypes.c_int] * 5 + [_vp] * 3
        _lib.sdfref_forward_device.restype = ctypes.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def interp(pgrid, grid, delta):
    """Reference helpers on grid coordinates pgrid [N,3] of one grid [d0,d1,d2] -> (values [N], gradients [N,3])."""
    lib = _load()
    pgrid, grid = _f32(pgrid), _f32(grid)
    n = pgrid.shape[0]
    val, grad = np.empty(n, np.float32), np.empty((n, 3), np.float32)
    c = lambda a: a.ctypes.data_as(_fp)
    lib.sdfref_value_interp(c(pgrid), n, c(grid), grid.shape[0], grid.shape[1], grid.shape[2], c(val))
    lib.sdfref_grad_interp(c(pgrid), n, c(grid), grid.shape[0], grid.shape[1], grid.shape[2], float(delta), c(grad))
    return val, grad


def forward_device(pose_init, sdf_grids, sdf_limits, points, epsilons, padding_scales, clearances, disables):
    """Same contract as omg_cuda.sdf_loss_forward (layers/omg_layers.cpp:24-49): fp32 contiguous CUDA tensors in,
    [potentials [N], potential_grads [N,3], collides [N]] out -- computed by the reference's own kernels."""
    import torch

    lib = _load()
    args = (pose_init, sdf_grids, sdf_limits, points, epsilons, padding_scales, clearances, disables)
    for t in args:
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
    n, o = points.shape[0], pose_init.shape[0]
    pot = torch.empty((n,), dtype=torch.float32, device=points.device)
    grad = torch.empty((n, 3), dtype=torch.float32, device=points.device)
    col = torch.empty((n,), dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()   # (the reference launches on the legacy default stream)
    with torch.cuda.device(points.device):
        rc = lib.sdfref_forward_device(*[_vp(t.data_ptr()) for t in args], n, o, sdf_grids.shape[1], sdf_grids.shape[2],
                                       sdf_grids.shape[3], _vp(pot.data_ptr()), _vp(grad.data_ptr()), _vp(col.data_ptr()))
    if rc != 0:
        raise RuntimeError("sdfref_forward_device: cudaError %d" % rc)
    return [pot, grad, col]
