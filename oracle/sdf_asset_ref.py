"""TEST INFRASTRUCTURE -- CPU restatement of the SDF asset path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

  FieldRef                omg/sdf_tools.py:17-44, 187-193 (SignedDensityField: constructor, resize, from_pth), with
                          its quirk: resize scales origin/min_coords and delta but not max_coords
  combine_sdfs            omg/core.py:366-411
  point_sdf               omg/core.py:426-452 (nearest-point distance; the reference uses scipy cKDTree.query, whose
                          squared-distance loop for 3-vectors is ((dx^2 + dy^2) + dz^2) then sqrt)

Pinned: tests/golden/assets_sdf.npz holds the outputs of the reference's own SignedDensityField.from_pth/.resize,
Env.combine_sdfs and PointEnv.compute_sdf_from_points (tools/make_golden_assets.py); tests/test_oracle_assets.py
replays them bit for bit."""
import numpy as np


class FieldRef(object):
    def __init__(self, data, origin, delta):
        self.data = data
        self.origin = origin
        self.delta = delta
        self.min_coords = origin
        self.max_coords = self.origin + delta * np.array(data.shape)
        self.data32 = self.data.astype(np.float32)                      # data_torch

    def resize(self, ratio):
        self.data *= ratio
        self.data32 = self.data32 * np.float32(ratio)
        self.delta *= ratio
        self.origin *= ratio                                           # (min_coords is the same array)

    @classmethod
    def from_stored(cls, stored_yxz, min_coords, delta):
        """from_pth on the arrays a .pth file holds: sdf_torch[0,0] is [Y,X,Z]."""
        return cls(np.ascontiguousarray(np.transpose(stored_yxz, (1, 0, 2))), np.array(min_coords), delta)


def combine_sdfs(fields):
    max_shape = np.array([f.data.shape for f in fields]).max(axis=0)
    num = len(fields)
    grids = np.ones((num, max_shape[0], max_shape[1], max_shape[2]), dtype=np.float32)
    limits = np.zeros((num, 10), dtype=np.float32)
    for i, f in enumerate(fields):
        size = f.data.shape
        grids[i, :size[0], :size[1], :size[2]] = f.data32
        mn, mx = f.min_coords, f.max_coords
        for k in range(3):
            limits[i, k] = mn[k]
            limits[i, 3 + k] = mn[k] + (mx[k] - mn[k]) * max_shape[k] / size[k]
            limits[i, 6 + k] = max_shape[k]
        limits[i, 9] = f.delta
    return grids, limits


def point_sdf(points, grid_resolution=0.02, margin=0.24, chunk=4096):
    points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    if points.shape[0] == 0:
        points = np.ones((2, 3)) * 3
    bounds = np.stack((points.min(0), points.max(0)), axis=1)
    axes = [np.arange(bounds[k][0] - margin, bounds[k][1] + margin, grid_resolution) for k in range(3)]
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 3)
    out = np.empty(grid.shape[0])
    for c0 in range(0, grid.shape[0], chunk):
        d = grid[c0:c0 + chunk, None, :] - points[None]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        out[c0:c0 + chunk] = np.sqrt(d2.min(1))
    shape = tuple(len(a) for a in axes)
    return out.reshape(shape), bounds[:, 0] - margin, axes
