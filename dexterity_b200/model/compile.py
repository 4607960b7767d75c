"""Flattens a composed `mjcf.Model` into the SoA arrays of include/dex_model_blob.h.

This is the build's replacement for MuJoCo's `mj_compile` + `mj_setConst` on the
scenes the suite builds: body tree in depth-first order,
joint/dof tables, inertias, fixed tendons, affine actuators, convex hulls of the
collision meshes, the statically filtered candidate geom-pair list with mixed
contact parameters, and the constants MuJoCo derives at qpos0 (`dof_invweight0`,
`body_invweight0`, `tendon_invweight0`, `stat.meaninertia`).
"""
from __future__ import annotations

import dataclasses
import os
import re
from typing import Dict, List, Optional, Tuple

import numpy as np

from dexterity_b200.model import meshes as mesh_lib
from dexterity_b200.model import mjcf

JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX, GEOM_MESH = 0, 2, 3, 6, 7
_GEOM_TYPES = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE,
               "box": GEOM_BOX, "mesh": GEOM_MESH}
_MINVAL = 1e-15

_HEADER = os.path.join(os.path.dirname(__file__), "..", "..", "include", "dex_model_blob.h")


def _parse_header() -> Tuple[List[Tuple[str, str]], Dict[str, int], Dict[str, int]]:
    src = open(_HEADER).read()
    fields = re.findall(r"X\((\w+),\s*DEX_BLOB_(I32|F64),", src)

    def enum(name: str) -> Dict[str, int]:
        body = re.search(r"enum %s \{(.*?)\};" % name, src, re.S).group(1)
        names = [t.strip().split("=")[0].strip() for t in body.replace("\n", " ").split(",") if t.strip()]
        return {n: i for i, n in enumerate(names)}

    return fields, enum("DexDim"), enum("DexOpt")


FIELDS, DIM_IDS, OPT_IDS = _parse_header()
FIELD_IDS = {n: i for i, (n, _) in enumerate(FIELDS)}
BLOB_MAGIC = 0x44455842
BLOB_VERSION = 3


@dataclasses.dataclass
class CompiledModel:
    """Arrays (by blob field name) + the name tables Python needs for `bind`-style lookups."""

    arrays: Dict[str, np.ndarray]
    body_names: List[str]
    joint_names: List[str]
    geom_names: List[str]
    site_names: List[str]
    actuator_names: List[str]
    tendon_names: List[str]

    def dim(self, name: str) -> int:
        return int(self.arrays["DIMS"][DIM_IDS["DEXD_" + name]])

    def opt(self, name: str) -> float:
        return float(self.arrays["OPT"][OPT_IDS["DEXO_" + name]])

    def __getattr__(self, item: str):
        arrays = object.__getattribute__(self, "arrays")
        key = item.upper()
        if key in arrays:
            return arrays[key]
        raise AttributeError(item)

    @property
    def nq(self): return self.dim("NQ")
    @property
    def nv(self): return self.dim("NV")
    @property
    def nu(self): return self.dim("NU")
    @property
    def nbody(self): return self.dim("NBODY")
    @property
    def ngeom(self): return self.dim("NGEOM")
    @property
    def nsite(self): return self.dim("NSITE")
    @property
    def npair(self): return self.dim("NPAIR")

    # ---- serialisation ------------------------------------------------- #
    def to_blob(self) -> bytes:
        nf = len(FIELDS)
        head = 24
        dirsz = 24 * nf
        off = head + dirsz
        entries = []
        chunks = []
        for name, dt in FIELDS:
            arr = self.arrays[name]
            arr = np.ascontiguousarray(arr, dtype=np.int32 if dt == "I32" else np.float64).ravel()
            raw = arr.tobytes()
            pad = (-len(raw)) % 8
            entries.append((0 if dt == "I32" else 1, off, arr.size))
            chunks.append(raw + b"\0" * pad)
            off += len(raw) + pad
        out = bytearray()
        out += np.array([BLOB_MAGIC, BLOB_VERSION, nf, 0], dtype=np.uint32).tobytes()
        out += np.array([off], dtype=np.uint64).tobytes()
        for dt, o, c in entries:
            out += np.array([dt, 0], dtype=np.uint32).tobytes()
            out += np.array([o, c], dtype=np.uint64).tobytes()
        for c in chunks:
            out += c
        assert len(out) == off
        return bytes(out)

    def save(self, path: str) -> None:
        names = dict(body=self.body_names, joint=self.joint_names, geom=self.geom_names,
                     site=self.site_names, actuator=self.actuator_names, tendon=self.tendon_names)
        np.savez_compressed(path, **self.arrays, **{"names_" + k: np.array(v) for k, v in names.items()})

    @staticmethod
    def load(path: str) -> "CompiledModel":
        z = np.load(path, allow_pickle=False)
        arrays = {n: z[n] for n, _ in FIELDS}
        g = lambda k: [str(s) for s in z["names_" + k]]
        return CompiledModel(arrays, g("body"), g("joint"), g("geom"), g("site"), g("actuator"), g("tendon"))


# --------------------------------------------------------------------------- #
# numpy reference kinematics (used for the qpos0 constants and in tests)
# --------------------------------------------------------------------------- #
def _axisangle_quat(axis: np.ndarray, ang: float) -> np.ndarray:
    return np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * axis])


def _rot(q: np.ndarray, v: np.ndarray) -> np.ndarray:
    return mjcf.quat_to_mat(q) @ v


def kinematics(a: Dict[str, np.ndarray], qpos: np.ndarray) -> Dict[str, np.ndarray]:
    """Body poses for `qpos` (textbook FK; restates what mj_kinematics produces)."""
    nbody = len(a["BODY_PARENT"])
    xpos = np.zeros((nbody, 3))
    xquat = np.zeros((nbody, 4))
    xquat[0, 0] = 1
    njnt = len(a["JNT_TYPE"])
    xanchor = np.zeros((njnt, 3))
    xaxis = np.zeros((njnt, 3))
    for b in range(1, nbody):
        p = a["BODY_PARENT"][b]
        pos = xpos[p] + _rot(xquat[p], a["BODY_POS"][b])
        quat = mjcf.quat_mul(xquat[p], a["BODY_QUAT"][b])
        for j in range(a["BODY_JNTADR"][b], a["BODY_JNTADR"][b] + a["BODY_JNTNUM"][b]):
            qa = a["JNT_QPOSADR"][j]
            if a["JNT_TYPE"][j] == JNT_FREE:
                pos = qpos[qa:qa + 3].copy()
                quat = qpos[qa + 3:qa + 7] / np.linalg.norm(qpos[qa + 3:qa + 7])
                xanchor[j] = pos
                xaxis[j] = _rot(quat, a["JNT_AXIS"][j])
            elif a["JNT_TYPE"][j] == JNT_HINGE:
                xanchor[j] = pos + _rot(quat, a["JNT_POS"][j])
                xaxis[j] = _rot(quat, a["JNT_AXIS"][j])
                quat = mjcf.quat_mul(quat, _axisangle_quat(a["JNT_AXIS"][j], qpos[qa] - a["QPOS0"][qa]))
                pos = xanchor[j] - _rot(quat, a["JNT_POS"][j])
            else:
                raise NotImplementedError
        xpos[b], xquat[b] = pos, quat / np.linalg.norm(quat)
    xmat = np.stack([mjcf.quat_to_mat(q) for q in xquat])
    xipos = xpos + np.einsum("bij,bj->bi", xmat, a["BODY_IPOS"])
    ximat = np.stack([xmat[b] @ mjcf.quat_to_mat(a["BODY_IQUAT"][b]) for b in range(nbody)])
    return dict(xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, ximat=ximat, xanchor=xanchor, xaxis=xaxis)


def point_jacobian(a: Dict[str, np.ndarray], kin: Dict[str, np.ndarray], body: int, point: np.ndarray):
    """(jacp, jacr) (3,nv) of a world point rigidly attached to `body`."""
    nv = len(a["DOF_BODY"])
    jp = np.zeros((3, nv))
    jr = np.zeros((3, nv))
    b = body
    while b > 0:
        for j in range(a["BODY_JNTADR"][b], a["BODY_JNTADR"][b] + a["BODY_JNTNUM"][b]):
            d = a["JNT_DOFADR"][j]
            if a["JNT_TYPE"][j] == JNT_FREE:
                jp[:, d:d + 3] = np.eye(3)
                R = kin["xmat"][b]
                for k in range(3):
                    jr[:, d + 3 + k] = R[:, k]
                    jp[:, d + 3 + k] = np.cross(R[:, k], point - kin["xpos"][b])
            else:
                jr[:, d] = kin["xaxis"][j]
                jp[:, d] = np.cross(kin["xaxis"][j], point - kin["xanchor"][j])
        b = a["BODY_PARENT"][b]
    return jp, jr


def mass_matrix(a: Dict[str, np.ndarray], kin: Dict[str, np.ndarray]) -> np.ndarray:
    """M(q) = Σ_b m Jpᵀ Jp + Jrᵀ I_world Jr + diag(armature) (definition, not CRB)."""
    nv = len(a["DOF_BODY"])
    M = np.diag(a["DOF_ARMATURE"].astype(np.float64))
    for b in range(1, len(a["BODY_PARENT"])):
        m = a["BODY_MASS"][b]
        if m == 0 and not a["BODY_INERTIA"][b].any():
            continue
        jp, jr = point_jacobian(a, kin, b, kin["xipos"][b])
        Iw = kin["ximat"][b] @ np.diag(a["BODY_INERTIA"][b]) @ kin["ximat"][b].T
        M += m * jp.T @ jp + jr.T @ Iw @ jr
    return M


# --------------------------------------------------------------------------- #
# Compilation
# --------------------------------------------------------------------------- #
def _geom_inertia(g: mjcf.Geom) -> Tuple[float, np.ndarray]:
    s = g.size
    if g.type == "box":
        vol = 8 * s[0] * s[1] * s[2]
        m = g.mass if g.mass is not None else g.density * vol
        I = m / 3.0 * np.array([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2])
    elif g.type == "sphere":
        vol = 4.0 / 3.0 * np.pi * s[0] ** 3
        m = g.mass if g.mass is not None else g.density * vol
        I = np.full(3, 0.4 * m * s[0] ** 2)
    elif g.type == "capsule":
        r, h = s[0], 2 * s[1]
        vc, vs = np.pi * r * r * h, 4.0 / 3.0 * np.pi * r ** 3
        vol = vc + vs
        m = g.mass if g.mass is not None else g.density * vol
        mc, ms = m * vc / vol, m * vs / vol
        iz = 0.5 * mc * r * r + 0.4 * ms * r * r
        ix = mc * (r * r / 4 + h * h / 12) + ms * (0.4 * r * r + 3.0 / 8.0 * r * h + h * h / 4)
        I = np.array([ix, ix, iz])
    else:
        raise NotImplementedError(f"inertia from geom type {g.type}")
    return float(m), I


def compile_model(model: mjcf.Model, timestep: float, nconmax: int = 64,
                  xfrc_gravcomp_prefix: Optional[str] = None) -> CompiledModel:
    """mj_compile + mj_setConst for the supported MJCF subset.

    `xfrc_gravcomp_prefix`: bodies whose name starts with it receive the constant
    `xfrc_applied = -gravity * mass` the reference hands apply every episode
    (reference utils/mujoco_utils.py:91-99, models/hands/shadow_hand_e.py:35-41).
    """
    bodies = model.bodies()
    nbody = len(bodies)
    bid = {id(b): i for i, b in enumerate(bodies)}
    body_names = [b.name for b in bodies]
    parent = np.zeros(nbody, np.int32)
    for b in bodies:
        for c in b.children:
            parent[bid[id(c)]] = bid[id(b)]

    opt = model.option
    gravity = mjcf._floats(opt.get("gravity", "0 0 -9.81"))

    A: Dict[str, np.ndarray] = {}
    A["BODY_PARENT"] = parent
    A["BODY_POS"] = np.array([b.pos for b in bodies])
    A["BODY_QUAT"] = np.array([b.quat for b in bodies])

    # joints / dofs
    jt, jb, jq, jd, jl = [], [], [], [], []
    jpos, jaxis, jrange, jmargin, jsr, jsi, jnames = [], [], [], [], [], [], []
    qpos0: List[float] = []
    dbody, djnt, darm, ddamp, dfl, dsr, dsi = [], [], [], [], [], [], []
    bjadr = -np.ones(nbody, np.int32)
    bjnum = np.zeros(nbody, np.int32)
    bdadr = -np.ones(nbody, np.int32)
    bdnum = np.zeros(nbody, np.int32)
    for i, b in enumerate(bodies):
        for j in b.joints:
            if bjadr[i] < 0:
                bjadr[i] = len(jt)
                bdadr[i] = len(dbody)
            bjnum[i] += 1
            jid = len(jt)
            jnames.append(j.name)
            jb.append(i)
            jq.append(len(qpos0))
            jd.append(len(dbody))
            jpos.append(j.pos)
            jaxis.append(j.axis)
            jrange.append(j.range)
            jmargin.append(j.margin)
            jsr.append(j.solreflimit)
            jsi.append(j.solimplimit)
            jl.append(int(j.limited))
            if j.type == "free":
                jt.append(JNT_FREE)
                qpos0 += list(b.pos) + list(b.quat)
                nd = 6
            elif j.type == "hinge":
                jt.append(JNT_HINGE)
                qpos0.append(j.ref)
                nd = 1
            else:
                raise NotImplementedError(j.type)
            for _ in range(nd):
                dbody.append(i)
                djnt.append(jid)
                darm.append(j.armature)
                ddamp.append(j.damping)
                dfl.append(j.frictionloss)
                dsr.append(j.solreffriction)
                dsi.append(j.solimpfriction)
            bdnum[i] += nd
    njnt, nq, nv = len(jt), len(qpos0), len(dbody)
    A["BODY_JNTADR"], A["BODY_JNTNUM"], A["BODY_DOFADR"], A["BODY_DOFNUM"] = bjadr, bjnum, bdadr, bdnum
    A["JNT_TYPE"] = np.array(jt, np.int32)
    A["JNT_BODY"] = np.array(jb, np.int32)
    A["JNT_QPOSADR"] = np.array(jq, np.int32)
    A["JNT_DOFADR"] = np.array(jd, np.int32)
    A["JNT_LIMITED"] = np.array(jl, np.int32)
    A["JNT_POS"] = np.array(jpos).reshape(njnt, 3)
    A["JNT_AXIS"] = np.array(jaxis).reshape(njnt, 3)
    A["JNT_RANGE"] = np.array(jrange).reshape(njnt, 2)
    A["JNT_MARGIN"] = np.array(jmargin, np.float64)
    A["JNT_SOLREF"] = np.array(jsr).reshape(njnt, 2)
    A["JNT_SOLIMP"] = np.array(jsi).reshape(njnt, 5)
    A["QPOS0"] = np.array(qpos0, np.float64)
    A["DOF_BODY"] = np.array(dbody, np.int32)
    A["DOF_JNT"] = np.array(djnt, np.int32)
    A["DOF_ARMATURE"] = np.array(darm, np.float64)
    A["DOF_DAMPING"] = np.array(ddamp, np.float64)
    A["DOF_FRICTIONLOSS"] = np.array(dfl, np.float64)
    A["DOF_SOLREF"] = np.array(dsr).reshape(nv, 2)
    A["DOF_SOLIMP"] = np.array(dsi).reshape(nv, 5)

    # weld / root ids, dof parents
    weld = np.zeros(nbody, np.int32)
    root = np.zeros(nbody, np.int32)
    for i in range(1, nbody):
        weld[i] = i if bjnum[i] > 0 else weld[parent[i]]
        root[i] = i if parent[i] == 0 else root[parent[i]]
    A["BODY_WELD"], A["BODY_ROOT"] = weld, root
    dparent = -np.ones(nv, np.int32)
    for d in range(nv):
        b = dbody[d]
        if d > bdadr[b]:
            dparent[d] = d - 1
        else:
            p = parent[b]
            while p > 0 and bdnum[p] == 0:
                p = parent[p]
            if p > 0:
                dparent[d] = bdadr[p] + bdnum[p] - 1
    A["DOF_PARENT"] = dparent

    # geoms (collidable ones, or referenced by explicit pairs), sites, inertias
    hulls = mesh_lib.HullCache()
    pair_geoms = {n for p in model.pairs for n in (p.geom1, p.geom2)}
    G: List[Tuple[int, mjcf.Geom]] = []
    for i, b in enumerate(bodies):
        for g in b.geoms:
            if g.contype or g.conaffinity or g.name in pair_geoms:
                G.append((i, g))
    ngeom = len(G)
    geom_names = [g.name for _, g in G]
    gtype = np.array([_GEOM_TYPES[g.type] for _, g in G], np.int32)
    gpos = np.zeros((ngeom, 3))
    gquat = np.zeros((ngeom, 4))
    gsize = np.zeros((ngeom, 3))
    grb = np.zeros(ngeom)
    gva = -np.ones(ngeom, np.int32)
    gvn = np.zeros(ngeom, np.int32)
    verts: List[np.ndarray] = []
    vadr: List[int] = []
    vnum: List[int] = []
    adj: List[int] = []
    mesh_slot: Dict[int, Tuple[int, int]] = {}
    nvert = 0
    for k, (i, g) in enumerate(G):
        gpos[k], gquat[k], gsize[k] = g.pos, g.quat, g.size
        if g.type == "mesh":
            asset = model.meshes[g.mesh]
            h = hulls.get(asset.file, asset.scale)
            gpos[k] = g.pos + mjcf.quat_to_mat(g.quat) @ h["center"]
            grb[k] = h["rbound"]
            if id(h) not in mesh_slot:
                mesh_slot[id(h)] = (nvert, len(h["verts"]))
                verts.append(h["verts"])
                for nb in h["adj"]:
                    vadr.append(len(adj))
                    vnum.append(len(nb))
                    adj += nb
                nvert += len(h["verts"])
            gva[k], gvn[k] = mesh_slot[id(h)]
        elif g.type == "box":
            grb[k] = np.linalg.norm(g.size)
        elif g.type == "sphere":
            grb[k] = g.size[0]
        elif g.type == "capsule":
            grb[k] = g.size[0] + g.size[1]
        elif g.type == "plane":
            grb[k] = 0.0
    A["GEOM_TYPE"], A["GEOM_BODY"] = gtype, np.array([i for i, _ in G], np.int32)
    A["GEOM_VERTADR"], A["GEOM_VERTNUM"] = gva, gvn
    A["GEOM_CONTYPE"] = np.array([g.contype for _, g in G], np.int32)
    A["GEOM_CONAFFINITY"] = np.array([g.conaffinity for _, g in G], np.int32)
    A["GEOM_POS"], A["GEOM_QUAT"], A["GEOM_SIZE"], A["GEOM_RBOUND"] = gpos, gquat, gsize, grb
    A["VERT"] = np.concatenate(verts) if verts else np.zeros((0, 3))
    A["VERT_ADJADR"] = np.array(vadr, np.int32)
    A["VERT_ADJNUM"] = np.array(vnum, np.int32)
    A["ADJ"] = np.array(adj, np.int32)

    site_names, sbody, spos, squat = [], [], [], []
    for i, b in enumerate(bodies):
        for s in b.sites:
            site_names.append(s.name)
            sbody.append(i)
            spos.append(s.pos)
            squat.append(s.quat)
    A["SITE_BODY"] = np.array(sbody, np.int32)
    A["SITE_POS"] = np.array(spos, np.float64).reshape(-1, 3)
    A["SITE_QUAT"] = np.array(squat, np.float64).reshape(-1, 4)

    ipos = np.zeros((nbody, 3))
    iquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
    mass = np.zeros(nbody)
    inertia = np.zeros((nbody, 3))
    for i, b in enumerate(bodies):
        if b.inertial is not None:
            ipos[i], iquat[i], mass[i], inertia[i] = b.inertial
        else:
            solid = [g for g in b.geoms if g.type != "plane" and i > 0 and (g.group < 3 or True)]
            solid = [g for g in solid if g.type in ("box", "sphere", "capsule")]
            if len(solid) == 1:
                mass[i], inertia[i] = _geom_inertia(solid[0])
                ipos[i], iquat[i] = solid[0].pos, solid[0].quat
            elif len(solid) > 1 and weld[i] != 0:
                raise NotImplementedError(f"composite inertia for body {b.name}")
    A["BODY_IPOS"], A["BODY_IQUAT"], A["BODY_MASS"], A["BODY_INERTIA"] = ipos, iquat, mass, inertia
    xfrc = np.zeros((nbody, 6))
    if xfrc_gravcomp_prefix is not None:
        for i, n in enumerate(body_names):
            if n.startswith(xfrc_gravcomp_prefix):
                xfrc[i, :3] = -gravity * mass[i]
    A["BODY_XFRC"] = xfrc

    # tendons / actuators
    jid_of = {n: i for i, n in enumerate(jnames)}
    # fixed tendons that are neither limited nor the transmission of an actuator only produce `ten_length`, which
    # nothing on the path reads (the Adroit hand has 44 of them): they are not compiled
    used = {a.tendon for a in model.actuators if a.tendon}
    tendons = [t for t in model.tendons if t.limited or t.name in used]
    tnames = [t.name for t in tendons]
    tadr, tnum, wj, wc = [], [], [], []
    for t in tendons:
        tadr.append(len(wj))
        tnum.append(len(t.joints))
        for jn, c in t.joints:
            wj.append(jid_of[jn])
            wc.append(c)
    nt = len(tnames)
    A["TEN_ADR"], A["TEN_NUM"] = np.array(tadr, np.int32), np.array(tnum, np.int32)
    A["TEN_LIMITED"] = np.array([int(t.limited) for t in tendons], np.int32)
    A["TEN_RANGE"] = np.array([t.range for t in tendons], np.float64).reshape(nt, 2)
    A["TEN_MARGIN"] = np.array([t.margin for t in tendons], np.float64)
    A["TEN_SOLREF"] = np.array([t.solreflimit for t in tendons], np.float64).reshape(nt, 2)
    A["TEN_SOLIMP"] = np.array([t.solimplimit for t in tendons], np.float64).reshape(nt, 5)
    A["WRAP_JNT"], A["WRAP_COEF"] = np.array(wj, np.int32), np.array(wc, np.float64)
    tid_of = {n: i for i, n in enumerate(tnames)}
    acts = model.actuators
    nu = len(acts)
    A["ACT_TRNTYPE"] = np.array([0 if a.joint else 1 for a in acts], np.int32)
    A["ACT_TRNID"] = np.array([jid_of[a.joint] if a.joint else tid_of[a.tendon] for a in acts], np.int32)
    A["ACT_CTRLLIMITED"] = np.array([int(a.ctrllimited) for a in acts], np.int32)
    A["ACT_FORCELIMITED"] = np.array([int(a.forcelimited) for a in acts], np.int32)
    A["ACT_GEAR"] = np.array([a.gear for a in acts], np.float64)
    A["ACT_GAINPRM"] = np.array([a.gainprm for a in acts], np.float64).reshape(nu, 3)
    A["ACT_BIASPRM"] = np.array([a.biasprm for a in acts], np.float64).reshape(nu, 3)
    A["ACT_CTRLRANGE"] = np.array([a.ctrlrange for a in acts], np.float64).reshape(nu, 2)
    A["ACT_FORCERANGE"] = np.array([a.forcerange for a in acts], np.float64).reshape(nu, 2)

    # ---- candidate geom pairs (static part of mj_collision's filtering) ----
    gid_of = {n: i for i, n in enumerate(geom_names)}
    bname_id = {n: i for i, n in enumerate(body_names)}
    excl = set()
    for a_, b_ in model.excludes:
        i, j = bname_id[a_], bname_id[b_]
        excl.add((min(i, j), max(i, j)))
    body_geoms: Dict[int, List[int]] = {}
    for k, (i, _) in enumerate(G):
        body_geoms.setdefault(i, []).append(k)
    P: List[dict] = []
    explicit = set()

    def mixed(g1: mjcf.Geom, g2: mjcf.Geom) -> dict:
        if g1.priority != g2.priority:
            gp = g1 if g1.priority > g2.priority else g2
            fr, sr, si, cd = gp.friction, gp.solref, gp.solimp, gp.condim
        else:
            cd = max(g1.condim, g2.condim)
            fr = np.maximum(g1.friction, g2.friction)
            s1, s2 = g1.solmix, g2.solmix
            if s1 >= _MINVAL and s2 >= _MINVAL:
                mix = s1 / (s1 + s2)
            elif s1 < _MINVAL and s2 < _MINVAL:
                mix = 0.5
            else:
                mix = 0.0 if s1 < _MINVAL else 1.0
            if g1.solref[0] > 0 and g2.solref[0] > 0:
                sr = mix * g1.solref + (1 - mix) * g2.solref
            else:
                sr = np.minimum(g1.solref, g2.solref)
            si = mix * g1.solimp + (1 - mix) * g2.solimp
        return dict(condim=cd, friction=np.array([fr[0], fr[0], fr[1], fr[2], fr[2]]), solref=sr,
                    solimp=si, margin=max(g1.margin, g2.margin), gap=max(g1.gap, g2.gap))

    def ordered(k1: int, k2: int) -> Tuple[int, int]:
        return (k2, k1) if gtype[k1] > gtype[k2] else (k1, k2)

    for p in model.pairs:
        k1, k2 = gid_of[p.geom1], gid_of[p.geom2]
        key = (min(k1, k2), max(k1, k2))
        if key in explicit:
            continue  # duplicated <pair> (adroit_hand.xml:247-248): MuJoCo keeps one signature
        explicit.add(key)
        prm = mixed(G[k1][1], G[k2][1])
        prm["condim"] = p.condim
        if p.friction is not None:
            f = list(p.friction) + list(prm["friction"][len(p.friction):])
            prm["friction"] = np.array(f)
        for key_, val in (("solref", p.solref), ("solimp", p.solimp), ("margin", p.margin), ("gap", p.gap)):
            if val is not None:
                prm[key_] = val if np.isscalar(val) else np.array(list(val) + list(prm[key_][len(val):]))
        k1, k2 = ordered(k1, k2)
        P.append(dict(g1=k1, g2=k2, **prm))

    wparent = weld[parent[weld]]
    for b1 in range(nbody):
        if b1 not in body_geoms:
            continue
        for b2 in range(b1 + 1, nbody):
            if b2 not in body_geoms:
                continue
            w1, w2 = weld[b1], weld[b2]
            if w1 == w2:
                continue
            if w1 != 0 and w2 != 0 and (w1 == wparent[w2] or w2 == wparent[w1]):
                continue
            if (b1, b2) in excl:
                continue
            for k1 in body_geoms[b1]:
                for k2 in body_geoms[b2]:
                    g1, g2 = G[k1][1], G[k2][1]
                    if not ((g1.contype & g2.conaffinity) or (g2.contype & g1.conaffinity)):
                        continue
                    if (min(k1, k2), max(k1, k2)) in explicit:
                        continue
                    o1, o2 = ordered(k1, k2)
                    P.append(dict(g1=o1, g2=o2, **mixed(G[o1][1], G[o2][1])))
    npair = len(P)
    A["PAIR_GEOM1"] = np.array([p["g1"] for p in P], np.int32)
    A["PAIR_GEOM2"] = np.array([p["g2"] for p in P], np.int32)
    A["PAIR_CONDIM"] = np.array([p["condim"] for p in P], np.int32)
    A["PAIR_FRICTION"] = np.array([p["friction"] for p in P], np.float64).reshape(npair, 5)
    A["PAIR_SOLREF"] = np.array([p["solref"] for p in P], np.float64).reshape(npair, 2)
    A["PAIR_SOLIMP"] = np.array([p["solimp"] for p in P], np.float64).reshape(npair, 5)
    A["PAIR_MARGIN"] = np.array([p["margin"] for p in P], np.float64)
    A["PAIR_GAP"] = np.array([p["gap"] for p in P], np.float64)

    # ---- constants at qpos0 (mj_setConst) ----
    kin = kinematics(A, A["QPOS0"])
    M = mass_matrix(A, kin)
    Minv = np.linalg.inv(M)
    dinv = np.zeros(nv)
    for j in range(njnt):
        d = jd[j]
        if jt[j] == JNT_FREE:
            dinv[d:d + 3] = np.mean(np.diag(Minv)[d:d + 3])
            dinv[d + 3:d + 6] = np.mean(np.diag(Minv)[d + 3:d + 6])
        else:
            dinv[d] = Minv[d, d]
    A["DOF_INVWEIGHT0"] = dinv
    binv = np.zeros((nbody, 2))
    for b in range(1, nbody):
        if weld[b] == 0:
            continue
        jp, jr = point_jacobian(A, kin, b, kin["xipos"][b])
        binv[b, 0] = max(np.trace(jp @ Minv @ jp.T) / 3, _MINVAL)
        binv[b, 1] = max(np.trace(jr @ Minv @ jr.T) / 3, _MINVAL)
    A["BODY_INVWEIGHT0"] = binv
    tinv = np.zeros(nt)
    for t in range(nt):
        J = np.zeros(nv)
        for w in range(tadr[t], tadr[t] + tnum[t]):
            J[jd[wj[w]]] = wc[w]
        tinv[t] = J @ Minv @ J
    A["TEN_INVWEIGHT0"] = tinv

    dims = np.zeros(len(DIM_IDS) - 1, np.int32)
    for name, val in dict(
        NBODY=nbody, NJNT=njnt, NQ=nq, NV=nv, NU=nu, NTENDON=nt, NWRAP=len(wj), NGEOM=ngeom,
        NVERT=nvert, NADJ=len(adj), NPAIR=npair, NSITE=len(site_names), NCONMAX=nconmax,
        SOLVER_ITERATIONS=int(opt.get("iterations", 100)),
        MPR_ITERATIONS=int(opt.get("mpr_iterations", 50)),
        LS_ITERATIONS=int(opt.get("ls_iterations", 50)),
    ).items():
        dims[DIM_IDS["DEXD_" + name]] = val
    A["DIMS"] = dims
    o = np.zeros(len(OPT_IDS) - 1)
    for name, val in dict(
        TIMESTEP=timestep, GRAVITY_X=gravity[0], GRAVITY_Y=gravity[1], GRAVITY_Z=gravity[2],
        TOLERANCE=float(opt.get("tolerance", 1e-8)), MPR_TOLERANCE=float(opt.get("mpr_tolerance", 1e-6)),
        IMPRATIO=float(opt.get("impratio", 1.0)), MEANINERTIA=float(np.mean(np.diag(M))),
        LS_TOLERANCE=float(opt.get("ls_tolerance", 0.01)),
    ).items():
        o[OPT_IDS["DEXO_" + name]] = val
    A["OPT"] = o

    return CompiledModel(A, body_names, jnames, geom_names, site_names, [a.name for a in acts], tnames)
