"""SignedDensityField: the reference's SDF carrier (omg/sdf_tools.py:17-44, 165-199) over device memory.

The reference keeps every object's grid twice (numpy `data` + CUDA `data_torch`) and materialises from_pth's
`permute(1,0,2)` on the host.  Here the raw grid goes to the device ONCE in the layout it has on disk; the permute,
`resize`'s scaling and the fp64->fp32 conversion are folded into the one pass that packs the scene
(`core.combine_sdfs` -> omgb_sdf_pack).  `data` / `data_torch` stay available (built on demand) for code that
reads them.

Reference quirks kept on purpose (they change sdf_limits, hence results):
  * `resize` scales `origin` (== `min_coords`, same array) and `delta` but NOT `max_coords`, which was computed in
    the constructor (sdf_tools.py:30, 37-44).
  * `penalize_constant` (omg/core.py:110) is applied to the numpy copy only, never to `data_torch`, so it does not
    reach the planner's SDF tensor; it is not applied here either.
"""
import pickle

import numpy as np
import torch


class SignedDensityField(object):
    """data[x, y, z]; `origin` is the world position of the corner of voxel (0,0,0), `delta` the voxel size."""

    def __init__(self, data, origin, delta, _raw=None, _layout=0, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("SignedDensityField keeps its grid in device memory; no CUDA device")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if _raw is None:
            host = np.ascontiguousarray(data)
            if host.dtype not in (np.float32, np.float64):
                host = host.astype(np.float64)
            _raw = torch.from_numpy(host).to(dev)
            shape = host.shape
        else:
            shape = tuple(data)   # from_pth passes the logical shape only
        self.raw = _raw                       # DEVICE, stored layout, unscaled
        self.layout = int(_layout)            # 0: [X,Y,Z]; 1: [Y,X,Z] (.pth files)
        self.scale = 1.0                      # pending resize ratio (applied by the pack kernel, fp32 multiply)
        self.nx, self.ny, self.nz = (int(s) for s in shape)
        self.origin = origin
        self.delta = delta
        self.min_coords = origin
        self.max_coords = self.origin + delta * np.array(shape)
        self._data = None
        self._data_torch = None

    # ---- what the reference exposes ------------------------------------------------------------------
    @property
    def shape(self):
        return (self.nx, self.ny, self.nz)

    @property
    def data_torch(self):
        """[X,Y,Z] fp32 CUDA (sdf_tools.py:32), built by the pack kernel on first use."""
        if self._data_torch is None:
            from .core import pack_sdf_grids

            self._data_torch = pack_sdf_grids([self], self.shape)[0]
        return self._data_torch

    @property
    def data(self):
        if self._data is None:
            raw = self.raw.cpu().numpy()
            raw = raw.transpose(1, 0, 2) if self.layout == 1 else raw
            self._data = raw * raw.dtype.type(self.scale) if self.scale != 1.0 else raw.copy()
        return self._data

    def resize(self, ratio):
        """sdf_tools.py:37-44 (max_coords is left as it was, like the reference)."""
        self.scale = float(np.float32(self.scale) * np.float32(ratio))
        self.delta *= ratio
        self.origin *= ratio
        self._data = None
        self._data_torch = None

    # ---- loaders ---------------------------------------------------------------------------------------
    @classmethod
    def from_pth(cls, sdf_file, device=None):
        """sdf_tools.py:187-193: {'min_coords', 'max_coords', 'delta', 'sdf_torch' [1,1,Y,X,Z]}
        (written by real_world/convert_sdf.py:30-77)."""
        sdf = torch.load(sdf_file, map_location="cpu", weights_only=False)
        min_coords = sdf["min_coords"].numpy()
        stored = sdf["sdf_torch"][0, 0]                    # [Y,X,Z]; the permute happens in the pack kernel
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        raw = stored.contiguous().to(dev)
        if raw.dtype not in (torch.float32, torch.float64):
            raw = raw.float()
        shape = (stored.shape[1], stored.shape[0], stored.shape[2])
        return cls(shape, min_coords, sdf["delta"], _raw=raw, _layout=1, device=dev)

    @classmethod
    def from_sdf(cls, sdf_file, device=None):
        """sdf_tools.py:165-185: text format `nx ny nz / x0 y0 z0 / delta / one value per line`, x fastest."""
        with open(sdf_file, "r") as fid:
            nx, ny, nz = map(int, fid.readline().split())
            x0, y0, z0 = map(float, fid.readline().split())
            delta = float(fid.readline().strip())
            vals = np.loadtxt(fid, dtype=np.float64, ndmin=1)
        data = np.zeros([nx, ny, nz])
        k = min(vals.shape[0], nx * ny * nz)
        flat = np.zeros(nx * ny * nz)
        flat[:k] = vals[:k]
        data[...] = flat.reshape(nz, ny, nx).transpose(2, 1, 0)   # i -> (i % nx, (i / nx) % ny, i / (nx ny))
        return cls(data, np.array([x0, y0, z0]), delta, device=device)

    @classmethod
    def from_pkl(cls, pkl_file, device=None):
        with open(pkl_file, "rb") as fid:
            data = pickle.load(fid)
        return cls(data["data"], data["origin"], data["delta"], device=device)
