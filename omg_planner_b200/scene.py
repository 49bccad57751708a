"""Synthetic stand-ins for the reference's scene data: `data/` (YCB SDFs, scene
.mat files, robot point clouds) is a 600 MB download that is absent, so every BASELINE config runs on
analytic signed-distance primitives sampled onto voxel grids with the reference's conventions.

Layout produced == what omg/core.py:366-411 (Env.combine_sdfs) hands to the operator:
  sdf_grids  [O, X, Y, Z] fp32, index x*Y*Z + y*Z + z, padded to the max shape with 1.0
  sdf_limits [O, 10] fp32 = (min xyz, max' xyz, dims xyz, delta), max' stretched to the padded shape
  voxel i is centred at (i + 0.5) * delta + min   (layers/sdf_matching_loss_kernel.cu:39-41)
  pose_mats  [O, 4, 4] fp64 object -> world (inverted per call, omg/cost.py:319)
"""
import numpy as np

PAD_VOXELS = 20  # real_world/gen_sdf.py:44-55 pads the object extent by a voxel margin

START_CONF = np.array([0.0, -1.285, 0, -2.356, 0.0, 1.571, 0.785, 0.04, 0.04])  # omg/core.py:38


def _sd_sphere(p, r):
    return np.sqrt((p * p).sum(-1)) - r


def _sd_box(p, h):
    q = np.abs(p) - h
    outside = np.sqrt((np.maximum(q, 0.0) ** 2).sum(-1))
    return outside + np.minimum(q.max(-1), 0.0)


def _sd_capsule(p, r, half_len):
    z = np.clip(p[..., 2], -half_len, half_len)
    d = p.copy()
    d[..., 2] -= z
    return np.sqrt((d * d).sum(-1)) - r


def _sample_grid(kind, params, dims, delta, origin):
    """fp64 analytic distance at voxel centres -> fp32 grid [X,Y,Z] (chunked over x)."""
    X, Y, Z = dims
    ys = (np.arange(Y) + 0.5) * delta + origin[1]
    zs = (np.arange(Z) + 0.5) * delta + origin[2]
    out = np.empty((X, Y, Z), np.float32)
    step = max(1, (1 << 21) // (Y * Z))
    for x0 in range(0, X, step):
        xs = (np.arange(x0, min(X, x0 + step)) + 0.5) * delta + origin[0]
        p = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), -1)
        if kind == "sphere":
            v = _sd_sphere(p, params[0])
        elif kind == "box":
            v = _sd_box(p, np.asarray(params))
        else:
            v = _sd_capsule(p, params[0], params[1])
        out[x0:x0 + xs.shape[0]] = v.astype(np.float32)
    return out


def _sample_grid_torch(kind, params, dims, delta, origin, device):
    """_sample_grid on a CUDA device (torch, fp64, the same IEEE operations in the same order; elementwise torch
    kernels do not fuse multiply-adds) -> fp32 torch tensor [X,Y,Z].  256^3 grids take milliseconds instead of
    seconds of numpy; tests/test_gpu_fullsize.py checks it is bit-identical to the numpy generator."""
    import torch

    X, Y, Z = dims
    f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)
    ys = f64((np.arange(Y) + 0.5) * delta + origin[1])[None, :, None]
    zs = f64((np.arange(Z) + 0.5) * delta + origin[2])[None, None, :]
    out = torch.empty((X, Y, Z), dtype=torch.float32, device=device)
    step = max(1, (1 << 24) // (Y * Z))
    for x0 in range(0, X, step):
        xs = f64((np.arange(x0, min(X, x0 + step)) + 0.5) * delta + origin[0])[:, None, None]
        shape = (xs.shape[0], Y, Z)
        px, py, pz = xs.expand(shape), ys.expand(shape), zs.expand(shape)
        if kind == "sphere":
            v = torch.sqrt(px * px + py * py + pz * pz) - params[0]
        elif kind == "box":
            qx, qy, qz = px.abs() - float(params[0]), py.abs() - float(params[1]), pz.abs() - float(params[2])
            cx, cy, cz = qx.clamp_min(0.0), qy.clamp_min(0.0), qz.clamp_min(0.0)
            outside = torch.sqrt(cx * cx + cy * cy + cz * cz)
            v = outside + torch.maximum(torch.maximum(qx, qy), qz).clamp_max(0.0)
        else:
            dz = pz - pz.clamp(-params[1], params[1])
            v = torch.sqrt(px * px + py * py + dz * dz) - params[0]
        out[x0:x0 + xs.shape[0]] = v.to(torch.float32)
    return out


def _yaw(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def make_scene(num_objects=10, grid=128, seed=0, table=True, grid_choices=None, device=None):
    """One table-top scene.  Object 0 is the grasp target; the last object is the table.
    grid_choices: optional list of per-object cubic grid sizes to draw from (mixed sizes exercise the
    pad-to-max path of combine_sdfs); default: every object is grid^3.
    device: a CUDA device -> the fields are sampled there and "sdf_grids" is a torch tensor on it (same values)."""
    sample = _sample_grid if device is None else (lambda *a: _sample_grid_torch(*a, device))
    rng = np.random.RandomState(1000 + seed)
    names, poses, grids, origins, deltas, dims = [], [], [], [], [], []
    n_free = num_objects - (1 if table else 0)
    for o in range(n_free):
        kind = ("sphere", "box", "capsule")[o % 3]
        if kind == "sphere":
            params = (rng.uniform(0.04, 0.10),)
            ext = 2 * params[0] * np.ones(3)
        elif kind == "box":
            params = tuple(rng.uniform(0.03, 0.10, 3))
            ext = 2 * np.asarray(params)
        else:
            params = (rng.uniform(0.03, 0.05), rng.uniform(0.04, 0.10))
            ext = np.array([2 * params[0], 2 * params[0], 2 * (params[0] + params[1])])
        g = int(grid if grid_choices is None else grid_choices[rng.randint(len(grid_choices))])
        delta = float(ext.max() / (g - 2 * min(PAD_VOXELS, g // 4)))
        origin = -0.5 * g * delta * np.ones(3)
        pose = np.eye(4)
        pose[:3, :3] = _yaw(rng.uniform(-np.pi, np.pi))
        pose[:3, 3] = [rng.uniform(0.3, 0.8), rng.uniform(-0.4, 0.4), rng.uniform(0.0, 0.5)]
        names.append("%03d_%s" % (o, kind)); poses.append(pose)
        grids.append(sample(kind, params, (g, g, g), delta, origin))
        origins.append(origin); deltas.append(delta); dims.append((g, g, g))
    if table:
        # bullet/panda_scene.py:582: table at (0.55, 0, -0.17), model y-up (quat 0.707,0.707,0,0 = Rx(90deg))
        half = np.array([0.5, 0.17, 0.8])
        g = int(grid if grid_choices is None else max(grid_choices))
        delta = float(2 * half.max() / (g - 2 * min(PAD_VOXELS, g // 4)))
        origin = -0.5 * g * delta * np.ones(3)
        pose = np.eye(4)
        pose[:3, :3] = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0.0]])
        pose[:3, 3] = [0.55, 0.0, -0.17]
        names.append("table"); poses.append(pose)
        grids.append(sample("box", tuple(half), (g, g, g), delta, origin))
        origins.append(origin); deltas.append(delta); dims.append((g, g, g))
    return pack_scene(names, poses, grids, origins, deltas, target_idx=0)


def pack_scene(names, poses, grids, origins, deltas, target_idx=0):
    """Same packing as omg/core.py:366-411: pad to the max shape with 1.0, stretch max' accordingly."""
    num = len(names)
    shapes = np.array([tuple(g.shape) for g in grids])
    mx = shapes.max(0)
    if isinstance(grids[0], np.ndarray):
        sdf = np.ones((num, mx[0], mx[1], mx[2]), np.float32)
    else:
        import torch
        sdf = torch.ones((num, int(mx[0]), int(mx[1]), int(mx[2])), dtype=torch.float32, device=grids[0].device)
    lim = np.zeros((num, 10), np.float32)
    for i in range(num):
        s = tuple(grids[i].shape)
        sdf[i, :s[0], :s[1], :s[2]] = grids[i]
        mn = np.asarray(origins[i], dtype=np.float64)
        mxc = mn + deltas[i] * np.array(s)
        lim[i, 0:3] = mn
        lim[i, 3:6] = mn + (mxc - mn) * mx / np.array(s)
        lim[i, 6:9] = mx
        lim[i, 9] = deltas[i]
    return {"names": list(names), "pose_mats": np.array(poses, dtype=np.float64), "sdf_grids": sdf,
            "sdf_limits": lim, "target_idx": int(target_idx), "attached": False}


def clamped_cubic(start, end, n):
    """omg/util.py:238-258 with two knots: CubicSpline(bc_type='clamped') through (0,start),(1,end)
    is the Hermite blend start + (end-start)(3t^2 - 2t^3), sampled at the n interior points of
    linspace(0,1,n+2)."""
    t = np.linspace(0, 1, n + 2)[1:-1, None]
    return start[None] + (end - start)[None] * (3 * t ** 2 - 2 * t ** 3)


def make_trajectories(batch, n, lower, upper, seed=0, tail=5):
    """start fixed (omg/core.py:38); per-trajectory goal uniform inside the padded joint
    limits (seed = trajectory index); clamped-cubic initialisation; goal-set tail = `tail` rows
    approaching the goal along the start->goal direction (stand-in for reach_grasps[g], [c,9])."""
    lower = np.asarray(lower, dtype=np.float64).reshape(-1)
    upper = np.asarray(upper, dtype=np.float64).reshape(-1)
    xi = np.zeros((batch, n, 9)); ends = np.zeros((batch, 9)); tails = np.zeros((batch, tail, 9))
    for b in range(batch):
        rng = np.random.RandomState(7919 * seed + b)
        end = START_CONF.copy()
        end[:7] = rng.uniform(lower[:7] + 0.05, upper[:7] - 0.05)
        u = end - START_CONF
        u = u / (np.linalg.norm(u) + 1e-12)
        for k in range(tail):
            row = end - (tail - 1 - k) / max(tail - 1, 1) * 0.15 * u
            row[:7] = np.clip(row[:7], lower[:7] + 0.01, upper[:7] - 0.01)
            row[7:] = 0.04
            tails[b, k] = row
        tails[b, -1] = end
        ends[b] = end
        xi[b] = clamped_cubic(START_CONF, end, n)
    starts = np.tile(START_CONF, (batch, 1))
    return xi, starts, ends, tails


def make_goal_sets(batch, num_goals, lower, upper, seed=0, tail=5, spread=0.6):
    """Synthetic goal sets (stand-in for the IK solutions of the target's grasps, omg/planner.py:296-455):
    per trajectory G goal configurations scattered around a seed goal inside the padded limits, plus for each the
    `tail` standoff rows approaching it (reach_grasps[g], [c,9]).  Returns goals [B,G,9], reach [B,G,tail,9]."""
    lower = np.asarray(lower, dtype=np.float64).reshape(-1)
    upper = np.asarray(upper, dtype=np.float64).reshape(-1)
    goals = np.zeros((batch, num_goals, 9)); reach = np.zeros((batch, num_goals, tail, 9))
    for b in range(batch):
        rng = np.random.RandomState(104729 * seed + b)
        centre = rng.uniform(lower[:7] + 0.3, upper[:7] - 0.3)
        for g in range(num_goals):
            end = START_CONF.copy()
            end[:7] = np.clip(centre + rng.uniform(-spread, spread, 7), lower[:7] + 0.05, upper[:7] - 0.05)
            u = end - START_CONF
            u = u / (np.linalg.norm(u) + 1e-12)
            for k in range(tail):
                row = end - (tail - 1 - k) / max(tail - 1, 1) * 0.15 * u
                row[:7] = np.clip(row[:7], lower[:7] + 0.01, upper[:7] - 0.01)
                row[7:] = 0.04
                reach[b, g, k] = row
            reach[b, g, -1] = end
            goals[b, g] = end
    return goals, reach
