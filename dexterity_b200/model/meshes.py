"""STL loading and convex-hull preprocessing for collision meshes.

MuJoCo collides mesh geoms as their convex hulls (qhull) with the geom frame
re-centred at the mesh centre of mass [upstream mjCMesh::Process; SURVEY.md §8c
checklist item 15].  We keep the hull vertices only (the support function of a
mesh equals the support function of its hull), re-centre them at the hull's
volume centroid, and build the hull edge graph for hill-climbing support queries.
Unlike MuJoCo we do NOT rotate the mesh into its principal inertia axes: the
collision geometry is rotation-frame independent, only the centre matters (it is
the MPR interior point).  "Parity vs MuJoCo unpinned" — see DESIGN.md.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np
from scipy.spatial import ConvexHull


def load_stl(path: str) -> np.ndarray:
    """Returns the (ntri, 3, 3) float64 triangle array of a binary or ASCII STL."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) >= 84:
        ntri = struct.unpack_from("<I", data, 80)[0]
        if 84 + 50 * ntri == len(data):
            rec = np.frombuffer(data, dtype=np.dtype([
                ("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=ntri, offset=84)
            return rec["v"].astype(np.float64)
    # ASCII fallback
    verts: List[List[float]] = []
    for line in data.decode("ascii", errors="ignore").splitlines():
        t = line.split()
        if len(t) == 4 and t[0] == "vertex":
            verts.append([float(t[1]), float(t[2]), float(t[3])])
    return np.array(verts, dtype=np.float64).reshape(-1, 3, 3)


def convex_hull(points: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (hull vertices (n,3), triangles (m,3) indexing them, outward oriented)."""
    pts = np.unique(points.reshape(-1, 3), axis=0)
    hull = ConvexHull(pts, qhull_options="Qt")
    vid = hull.vertices
    remap = -np.ones(len(pts), dtype=np.int64)
    remap[vid] = np.arange(len(vid))
    tris = remap[hull.simplices]
    v = pts[vid]
    # orient outward using qhull's facet normals
    n = np.cross(v[tris[:, 1]] - v[tris[:, 0]], v[tris[:, 2]] - v[tris[:, 0]])
    flip = np.einsum("ij,ij->i", n, hull.equations[:, :3]) < 0
    tris[flip] = tris[flip][:, ::-1]
    return v, tris


def volume_centroid(v: np.ndarray, tris: np.ndarray) -> Tuple[float, np.ndarray]:
    a, b, c = v[tris[:, 0]], v[tris[:, 1]], v[tris[:, 2]]
    ref = v.mean(axis=0)
    a, b, c = a - ref, b - ref, c - ref
    vol = np.einsum("ij,ij->i", a, np.cross(b, c)) / 6.0
    cen = (a + b + c) / 4.0
    total = vol.sum()
    return float(total), ref + (vol[:, None] * cen).sum(axis=0) / total


def hull_graph(nvert: int, tris: np.ndarray) -> List[List[int]]:
    adj: List[set] = [set() for _ in range(nvert)]
    for t in tris:
        for i in range(3):
            a, b = int(t[i]), int(t[(i + 1) % 3])
            adj[a].add(b)
            adj[b].add(a)
    return [sorted(s) for s in adj]


class HullCache:
    """file+scale → processed hull (shared between geoms that reuse a mesh)."""

    def __init__(self) -> None:
        self._cache: Dict[Tuple[str, Tuple[float, ...]], dict] = {}

    def get(self, path: str, scale: np.ndarray) -> dict:
        key = (path, tuple(float(s) for s in scale))
        if key not in self._cache:
            tri = load_stl(path) * scale[None, None, :]
            v, t = convex_hull(tri)
            vol, cen = volume_centroid(v, t)
            vc = v - cen
            self._cache[key] = dict(
                verts=vc, tris=t, center=cen, volume=vol,
                rbound=float(np.linalg.norm(vc, axis=1).max()),
                adj=hull_graph(len(vc), t),
            )
        return self._cache[key]
