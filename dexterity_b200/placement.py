"""Episode-start prop placement (reference manipulation/tasks/reorient.py:143-151,182-188).

The reference places the cube with dm_control's `initializers.PropPlacer(position=Uniform(box),
quaternion=UniformQuaternion(), settle_physics=False)`: poses are redrawn while the prop is in
contact with anything (dist <= 0), up to 20 draws, then it raises.  The device code for that loop
is `dex_place_env` (csrc/dex_task.cuh); each rejection round costs a pass of the collision kernels
over the envs being reset.

`spawn_clearance` decides at construction time whether the loop can ever fire.  With the hand at
its reset pose (`qpos0`, reference shadow_hand_e.py initialises nothing else for reorient) every
other collidable geom G is a fixed convex set, and the prop, whatever its orientation, lies inside
the ball of its bounding radius r around its centre p, p in the spawn box B.  A contact needs
dist(G, p) <= r, hence dist(G, B) <= r.  For any unit direction n,
    dist(G, B) >= min_{b in B} n.b - max_{g in G} n.g      (separating-axis bound),
so the largest such gap over a fixed set of directions minus r is a LOWER bound of the clearance.
A positive value proves that no draw is ever rejected, and the env configures
`place_max_attempts = 0` (no rounds launched); otherwise the device loop runs with PropPlacer's 20.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from dexterity_b200.model import compile as compile_lib
from dexterity_b200.model import mjcf

PROP_PLACER_MAX_ATTEMPTS = 20   # dm_control initializers.PropPlacer default max_attempts_per_prop


def _directions(n: int = 1024) -> np.ndarray:
    """Fibonacci sphere plus the coordinate axes (deterministic)."""
    i = np.arange(n) + 0.5
    z = 1 - 2 * i / n
    phi = np.pi * (1 + 5 ** 0.5) * i
    r = np.sqrt(1 - z * z)
    d = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    return np.concatenate([d, np.eye(3), -np.eye(3)])


def _geom_world_points(a, kin, g: int) -> np.ndarray:
    """World-frame points whose convex hull contains geom `g` (hull vertices, box corners, or the
    corners of the bounding cube of radius rbound for other primitives)."""
    b = int(a["GEOM_BODY"][g])
    rg = mjcf.quat_to_mat(a["GEOM_QUAT"][g])
    t = int(a["GEOM_TYPE"][g])
    if t == compile_lib.GEOM_MESH:
        va, vn = int(a["GEOM_VERTADR"][g]), int(a["GEOM_VERTNUM"][g])
        local = a["VERT"][va:va + vn, :3]
    else:
        h = a["GEOM_SIZE"][g][:3] if t == compile_lib.GEOM_BOX else np.full(3, a["GEOM_RBOUND"][g])
        local = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float) * h
    in_body = a["GEOM_POS"][g] + local @ rg.T
    return kin["xpos"][b] + in_body @ kin["xmat"][b].T


def spawn_clearance(compiled, prop_body: int, spawn_lo: Sequence[float], spawn_hi: Sequence[float],
                    qpos: np.ndarray | None = None) -> float:
    """Lower bound (metres) of the gap between the prop, anywhere in the spawn box at any
    orientation, and every other collidable geom at `qpos` (default qpos0); the ground plane included."""
    a = compiled.arrays
    qpos = a["QPOS0"] if qpos is None else qpos
    kin = compile_lib.kinematics(a, qpos)
    lo, hi = np.asarray(spawn_lo, float), np.asarray(spawn_hi, float)
    corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    prop_geoms = [g for g in range(len(a["GEOM_BODY"])) if a["GEOM_BODY"][g] == prop_body]
    if not prop_geoms:
        raise ValueError("prop body has no geoms")
    # bounding radius of the whole prop about the body origin (the free joint's position)
    r_prop = max(float(np.linalg.norm(a["GEOM_POS"][g]) + a["GEOM_RBOUND"][g]) for g in prop_geoms)
    dirs = _directions()
    box_min = (corners @ dirs.T).min(axis=0)
    best = np.inf
    for g in range(len(a["GEOM_BODY"])):
        if g in prop_geoms or (a["GEOM_CONTYPE"][g] == 0 and a["GEOM_CONAFFINITY"][g] == 0):
            continue
        if not any((a["GEOM_CONTYPE"][g] & a["GEOM_CONAFFINITY"][p]) or (a["GEOM_CONTYPE"][p] & a["GEOM_CONAFFINITY"][g])
                   for p in prop_geoms):
            continue
        if a["GEOM_TYPE"][g] == compile_lib.GEOM_PLANE:
            b = int(a["GEOM_BODY"][g])
            n = kin["xmat"][b] @ mjcf.quat_to_mat(a["GEOM_QUAT"][g])[:, 2]
            p0 = kin["xpos"][b] + kin["xmat"][b] @ a["GEOM_POS"][g]
            gap = float(((corners - p0) @ n).min())
        else:
            w = _geom_world_points(a, kin, g)
            gap = float((box_min - (w @ dirs.T).max(axis=0)).max())
        best = min(best, gap - r_prop)
    return best


def place_max_attempts(compiled, prop_body: int, spawn_lo, spawn_hi, margin: float = 1e-4) -> int:
    """0 when the spawn region is provably contact-free (no rejection rounds needed), else PropPlacer's 20."""
    return 0 if spawn_clearance(compiled, prop_body, spawn_lo, spawn_hi) > margin else PROP_PLACER_MAX_ATTEMPTS
