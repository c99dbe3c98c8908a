"""Seeded synthetic inputs of the reference's benchmark scenes (numpy, host side).

Follows the recipe of the reference's scene generator, not its random stream (the reference
seeds from std::random_device, random_map_generator.cpp:62,98):
  * cuboids  — RandomPCGenerator::generataRandomCaseAux (random_map_generator.cpp:342-443)
               with params/map_cuboids.yaml (80 ground + 80 floating boxes)
  * tables   — RandomPCGenerator::generateDeskCase (random_map_generator.cpp:207-325)
               with params/map_tables.yaml (40 desks + 80 wall boxes)
Points are float32 xyz at 0.05 m pitch, exactly the input format of GridMap::regenerateMap
(grid_map.cpp:733-747). Also generates candidate waypoint paths (x, y, yaw, q1..q7) in the
format MomaTrajOpt::optimizeTraj consumes (moma_traj_opt.cpp:142).
"""
import numpy as np

QMAX = np.array([3.1, 2.26, 3.1, 2.355, 3.1, 2.23, 6.28])  # moma_param.h:116


def _box_points(pos, size, res):
    """Box::generatePCL (random_map_generator.cpp:6-31) for theta = 0; float32 like pcl::PointXYZ."""
    nx, ny, nz = (int(np.ceil(s / res)) for s in size)
    i = (np.arange(nx) * res).astype(np.float32).astype(np.float64)
    j = (np.arange(ny) * res).astype(np.float32).astype(np.float64)
    k = (np.arange(nz) * res).astype(np.float32).astype(np.float64)
    X, Y, Z = np.meshgrid(i + pos[0], j + pos[1], k + pos[2], indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)


def _walls(size_x, size_y, res):
    """Perimeter wall, 1 m high, two cells thick (random_map_generator.cpp:350-369)."""
    pts = []
    b = _box_points((0, 0, 0), (size_x, res * 2.0, 1.0), res).astype(np.float64)
    for sy in (+size_y / 2.0, -size_y / 2.0):
        p = b.copy()
        p[:, 0] += -size_x / 2.0 - res
        p[:, 1] += sy - res
        pts.append(p.astype(np.float32))
    b = _box_points((0, 0, 0), (res * 2.0, size_y, 1.0), res).astype(np.float64)
    for sx in (+size_x / 2.0, -size_x / 2.0):
        p = b.copy()
        p[:, 0] += sx - res
        p[:, 1] += -size_y / 2.0 - res
        pts.append(p.astype(np.float32))
    return pts


def _overlap(a_pos, a_size, b_pos, b_size, with_z=True):
    """Box::overlap / overlap2d (random_map.hpp:56-85) for axis-aligned boxes."""
    for d in (0, 1):
        if a_pos[d] + a_size[d] < b_pos[d] or b_pos[d] + b_size[d] < a_pos[d]:
            return False
    if with_z:
        return a_pos[2] + a_size[2] > b_pos[2] and a_pos[2] < b_pos[2] + b_size[2]
    return True


def cuboids_scene(seed, size_x=20.0, size_y=20.0, res=0.05, obs_num=(80, 80), wall_size=(0.2, 0.8),
                  wall_height=(0.4, 1.5), float_size=(0.3, 0.6), float_height=(0.4, 0.8), scale=1.0):
    """Returns (points float32 [n,3], boxes [m,6] = pos,size). `scale` multiplies the obstacle counts
    by scale^2 so a 40 m map keeps the 20 m map's density."""
    rng = np.random.default_rng(seed)
    pts = _walls(size_x, size_y, res)
    spawn = (np.array([-0.5, -0.5, -0.5]), np.array([1.0, 1.0, 1.0]))
    boxes = []
    counts = [int(round(n * scale * scale)) for n in obs_num]
    for k in range(2):
        j = 0
        while j < counts[k]:
            x = np.floor(rng.uniform(-size_x / 2, size_x / 2) / res) * res + res / 2.0
            y = np.floor(rng.uniform(-size_y / 2, size_y / 2) / res) * res + res / 2.0
            if k == 0:
                size = np.array([rng.uniform(*wall_size), rng.uniform(*wall_size), rng.uniform(*wall_height)])
                h = 0.0
            else:
                size = rng.uniform(float_size[0], float_size[1], 3)
                h = rng.uniform(*float_height)
            pos = np.array([x, y, h])
            if any(_overlap(pos, size, bp, bs) for bp, bs in boxes) or _overlap(pos, size, *spawn, with_z=False):
                continue
            boxes.append((pos, size))
            p = _box_points(pos, size, res)
            fr = np.float32(0.5)
            keep = ~((p[:, 0] > -fr) & (p[:, 0] < fr) & (p[:, 1] > -fr) & (p[:, 1] < fr))
            pts.append(p[keep])
            j += 1
    return np.concatenate(pts, axis=0), np.array([np.concatenate(b) for b in boxes])


def tables_scene(seed, spawn_xy=((0.0, 0.0),), size_x=20.0, size_y=20.0, res=0.05, obs_num=(40, 80),
                 wall_size=(0.2, 0.8), wall_height=(0.4, 1.5), desk_len=(0.75, 1.25), desk_wid=(0.75, 1.25),
                 desk_h=(0.5, 1.0), arrangement=(1, 2)):
    rng = np.random.default_rng(seed)
    pts = _walls(size_x, size_y, res)
    boxes = [(np.array([p[0] - 0.5, p[1] - 0.5, 0.0]), np.array([1.0, 1.0, 1.0])) for p in spawn_xy]
    leg, top = 0.05, 0.05
    i = 0
    while i < obs_num[0]:
        x = np.floor(rng.uniform(-size_x / 2, size_x / 2) / res) * res + res / 2.0
        y = np.floor(rng.uniform(-size_y / 2, size_y / 2) / res) * res + res / 2.0
        sx, sy, h = rng.uniform(*desk_wid), rng.uniform(*desk_len), rng.uniform(*desk_h)
        r, c = rng.integers(arrangement[0], arrangement[1] + 1, 2)
        pos, size = np.array([x, y, 0.0]), np.array([sx * r, sy * c, h])
        if any(_overlap(pos, size, bp, bs) for bp, bs in boxes):
            continue
        boxes.append((pos, size))
        for rr in range(r):
            for cc in range(c):
                p0 = np.array([x + rr * sx, y + cc * sy, 0.0])
                for ox, oy in ((0, 0), (sx - leg, 0), (0, sy - leg), (sx - leg, sy - leg)):
                    pts.append(_box_points(p0 + np.array([ox, oy, 0.0]), (leg, leg, h), res))
                pts.append(_box_points(np.array([p0[0], p0[1], h]), (sx, sy, top), res))
        i += 1
    i = 0
    while i < obs_num[1]:
        x = np.floor(rng.uniform(-size_x / 2, size_x / 2) / res) * res + res / 2.0
        y = np.floor(rng.uniform(-size_y / 2, size_y / 2) / res) * res + res / 2.0
        size = np.array([rng.uniform(*wall_size), rng.uniform(*wall_size), rng.uniform(*wall_height)])
        pos = np.array([x, y, 0.0])
        if any(_overlap(pos, size, bp, bs) for bp, bs in boxes):
            continue
        boxes.append((pos, size))
        pts.append(_box_points(pos, size, res))
        i += 1
    return np.concatenate(pts, axis=0), np.array([np.concatenate(b) for b in boxes])


def random_joints(rng):
    """planner.cpp:533-537: uniform in the joint limits."""
    return (2.0 * QMAX) * rng.uniform(0.0, 1.0, 7) - QMAX


def scurve_candidate(rng, start_xy, goal_xy, n_waypoints=64, lateral=1.5, q_scale=0.5):
    """A smooth S-curve guide path of n_waypoints 10-D waypoints:
    straight start->goal line plus a lateral offset a1*sin(pi s) + a2*sin(2 pi s), yaw along the
    tangent, joints linearly interpolated between two random configurations."""
    s = np.linspace(0.0, 1.0, n_waypoints)
    d = np.asarray(goal_xy, float) - np.asarray(start_xy, float)
    L = np.linalg.norm(d)
    t, nrm = d / L, np.array([-d[1], d[0]]) / L
    a1, a2 = rng.uniform(-lateral, lateral, 2)
    off = a1 * np.sin(np.pi * s) + a2 * np.sin(2 * np.pi * s)
    xy = np.asarray(start_xy)[None, :] + s[:, None] * d[None, :] + off[:, None] * nrm[None, :]
    doff = a1 * np.pi * np.cos(np.pi * s) + 2 * a2 * np.pi * np.cos(2 * np.pi * s)
    tang = d[None, :] + doff[:, None] * nrm[None, :]
    yaw = np.unwrap(np.arctan2(tang[:, 1], tang[:, 0]))
    q0, q1 = random_joints(rng) * q_scale, random_joints(rng) * q_scale
    q = q0[None, :] + s[:, None] * (q1 - q0)[None, :]
    return np.concatenate([xy, yaw[:, None], q], axis=1)


def synthetic_batch(n_cand, seed=1234, n_waypoints=64, span=8.0, lateral=1.5):
    """n_cand S-curve candidates crossing the 20 m map; zero boundary velocity/acceleration
    (planner.cpp:873-875 with start_v = 0). Returns (paths list, bvel [n,10,2], bacc [n,10,2])."""
    paths = []
    for c in range(n_cand):
        rng = np.random.default_rng(seed + c)
        ang = rng.uniform(-np.pi, np.pi)
        ctr = rng.uniform(-1.0, 1.0, 2)
        half = 0.5 * rng.uniform(0.75, 1.0) * 2 * span
        a = ctr - half * np.array([np.cos(ang), np.sin(ang)])
        b = ctr + half * np.array([np.cos(ang), np.sin(ang)])
        paths.append(scurve_candidate(rng, np.clip(a, -span, span), np.clip(b, -span, span), n_waypoints, lateral))
    return paths, np.zeros((n_cand, 10, 2)), np.zeros((n_cand, 10, 2))


def short_candidates(n_cand, seed, dist_range=(3.0, 8.0), n_waypoints=8, lateral=0.6, span=8.0):
    """Candidates at the reference's default scale (start/goal 3-8 m apart, planner.cpp:494-512;
    N ~ 3-10 pieces at sample_interval 1.5 s): the parity-test and latency workload."""
    paths = []
    rng = np.random.default_rng(seed)
    while len(paths) < n_cand:
        a, b = rng.uniform(-span, span, 2), rng.uniform(-span, span, 2)
        if not (dist_range[0] <= np.linalg.norm(a - b) <= dist_range[1]):
            continue
        paths.append(scurve_candidate(rng, a, b, n_waypoints, lateral))
    return paths, np.zeros((n_cand, 10, 2)), np.zeros((n_cand, 10, 2))
