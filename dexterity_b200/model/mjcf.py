"""Minimal MJCF reader + scene composition (host side, runs once).

The reference assembles its scenes with dm_control's PyMJCF (`mjcf.from_path`,
`arena.attach`, `add_free_entity`; reference dexterity/models/arenas/arena.py:40-56,
manipulation/tasks/reorient.py:113-168) and lets MuJoCo compile them.  Neither
package is available here, so this module re-implements the subset of MJCF the
suite's hand models use: `<include>`, nested `<default>` classes with
`childclass` inheritance, `<body>/<inertial>/<joint>/<geom>/<site>`,
`<contact><exclude>/<pair>`, fixed `<tendon>`s and `<position>/<general>/<motor>`
actuators.  Attaching a model under a scene nests its default classes below the
scene's root default, exactly as PyMJCF's `attach` does (so arena-level geom
defaults propagate into attached hands and props; SURVEY.md §8 model facts).

Everything is resolved to plain Python dataclasses with numeric attributes; the
flattening to arrays lives in compile.py.
"""
from __future__ import annotations

import dataclasses
import math
import os
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------- #
# MuJoCo built-in attribute defaults for the element kinds we support.
# --------------------------------------------------------------------------- #
_DEFAULT_SOLREF = (0.02, 1.0)
_DEFAULT_SOLIMP = (0.9, 0.95, 0.001, 0.5, 2.0)

BUILTIN: Dict[str, Dict[str, str]] = {
    "joint": {
        "type": "hinge", "pos": "0 0 0", "axis": "0 0 1", "limited": "false",
        "range": "0 0", "armature": "0", "damping": "0", "frictionloss": "0",
        "margin": "0", "stiffness": "0", "ref": "0", "springref": "0",
    },
    "geom": {
        "type": "sphere", "contype": "1", "conaffinity": "1", "condim": "3",
        "group": "0", "priority": "0", "size": "0 0 0", "friction": "1 0.005 0.0001",
        "solmix": "1", "margin": "0", "gap": "0", "density": "1000",
    },
    "site": {"pos": "0 0 0", "size": "0.005 0.005 0.005", "type": "sphere"},
    "mesh": {"scale": "1 1 1"},
    "actuator": {
        "ctrllimited": "false", "forcelimited": "false", "gear": "1",
        "ctrlrange": "0 0", "forcerange": "0 0",
    },
    "tendon": {"limited": "false", "range": "0 0", "margin": "0", "frictionloss": "0",
               "stiffness": "0", "damping": "0"},
}

_ACTUATOR_TAGS = ("general", "motor", "position", "velocity")


def _floats(s: str) -> np.ndarray:
    return np.array([float(x) for x in s.split()], dtype=np.float64)


def _bool(s: str) -> bool:
    return s.strip().lower() == "true"


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_to_mat(q: np.ndarray) -> np.ndarray:
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat_to_quat(m: np.ndarray) -> np.ndarray:
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
    q = np.array(q)
    return q / np.linalg.norm(q)


def euler_to_quat(e: np.ndarray, seq: str = "xyz") -> np.ndarray:
    """MuJoCo `eulerseq` semantics: lower-case = intrinsic (rotating) axes."""
    q = np.array([1.0, 0, 0, 0])
    for ang, ax in zip(e, seq):
        h = 0.5 * ang
        r = np.array([math.cos(h), 0, 0, 0])
        r["xyz".index(ax.lower()) + 1] = math.sin(h)
        q = quat_mul(q, r) if ax.islower() else quat_mul(r, q)
    return q


# --------------------------------------------------------------------------- #
# Resolved element dataclasses
# --------------------------------------------------------------------------- #
@dataclasses.dataclass
class Joint:
    name: str
    type: str  # free | hinge | slide
    pos: np.ndarray
    axis: np.ndarray
    limited: bool
    range: np.ndarray
    armature: float
    damping: float
    frictionloss: float
    margin: float
    solreflimit: np.ndarray
    solimplimit: np.ndarray
    solreffriction: np.ndarray
    solimpfriction: np.ndarray
    ref: float = 0.0


@dataclasses.dataclass
class Geom:
    name: str
    type: str  # plane | sphere | capsule | box | mesh
    pos: np.ndarray
    quat: np.ndarray
    size: np.ndarray
    contype: int
    conaffinity: int
    condim: int
    priority: int
    friction: np.ndarray
    solref: np.ndarray
    solimp: np.ndarray
    solmix: float
    margin: float
    gap: float
    density: float
    mass: Optional[float]
    mesh: Optional[str]
    group: int = 0


@dataclasses.dataclass
class Site:
    name: str
    pos: np.ndarray
    quat: np.ndarray


@dataclasses.dataclass
class Body:
    name: str
    pos: np.ndarray
    quat: np.ndarray
    inertial: Optional[Tuple[np.ndarray, np.ndarray, float, np.ndarray]]  # ipos, iquat, mass, diag
    joints: List[Joint] = dataclasses.field(default_factory=list)
    geoms: List[Geom] = dataclasses.field(default_factory=list)
    sites: List[Site] = dataclasses.field(default_factory=list)
    children: List["Body"] = dataclasses.field(default_factory=list)

    def walk(self):
        yield self
        for c in self.children:
            yield from c.walk()


@dataclasses.dataclass
class Tendon:
    name: str
    joints: List[Tuple[str, float]]
    limited: bool
    range: np.ndarray
    margin: float
    solreflimit: np.ndarray
    solimplimit: np.ndarray


@dataclasses.dataclass
class Actuator:
    name: str
    joint: Optional[str]
    tendon: Optional[str]
    gear: float
    gainprm: np.ndarray
    biasprm: np.ndarray
    ctrllimited: bool
    ctrlrange: np.ndarray
    forcelimited: bool
    forcerange: np.ndarray


@dataclasses.dataclass
class ContactPair:
    geom1: str
    geom2: str
    condim: int
    friction: Optional[np.ndarray]
    solref: Optional[np.ndarray]
    solimp: Optional[np.ndarray]
    margin: Optional[float]
    gap: Optional[float]


@dataclasses.dataclass
class MeshAsset:
    name: str
    file: str
    scale: np.ndarray


@dataclasses.dataclass
class Model:
    """A resolved MJCF model (one XML file) or a composed scene."""

    name: str
    world: Body
    meshes: Dict[str, MeshAsset]
    tendons: List[Tendon]
    actuators: List[Actuator]
    excludes: List[Tuple[str, str]]
    pairs: List[ContactPair]
    option: Dict[str, str]
    root_defaults: Dict[str, Dict[str, str]]

    def bodies(self):
        return list(self.world.walk())

    def find_body(self, name: str) -> Body:
        for b in self.world.walk():
            if b.name == name:
                return b
        raise KeyError(name)


# --------------------------------------------------------------------------- #
# Parsing
# --------------------------------------------------------------------------- #
def _expand_includes(elem: ET.Element, base_dir: str) -> None:
    i = 0
    while i < len(elem):
        child = elem[i]
        if child.tag == "include":
            sub = ET.parse(os.path.join(base_dir, child.attrib["file"])).getroot()
            _expand_includes(sub, base_dir)
            elem.remove(child)
            for k, sc in enumerate(list(sub)):
                elem.insert(i + k, sc)
            i += len(sub)
        else:
            _expand_includes(child, base_dir)
            i += 1


class _Defaults:
    """Default-class tree.  `classes[name][tag]` is the fully inherited attr dict."""

    def __init__(self, base: Optional[Dict[str, Dict[str, str]]] = None):
        self.classes: Dict[str, Dict[str, Dict[str, str]]] = {}
        self.base = {k: dict(v) for k, v in (base or {}).items()}

    @staticmethod
    def _tagkey(tag: str) -> str:
        return "actuator" if tag in _ACTUATOR_TAGS else tag

    def load(self, default_elem: Optional[ET.Element]) -> None:
        self.classes["main"] = {k: dict(v) for k, v in self.base.items()}
        if default_elem is not None:
            self._load(default_elem, "main", top=True)

    def _load(self, elem: ET.Element, cls: str, top: bool = False) -> None:
        cur = self.classes[cls]
        for child in elem:
            if child.tag == "default":
                name = child.attrib["class"]
                self.classes[name] = {k: dict(v) for k, v in cur.items()}
                self._load(child, name)
            else:
                key = self._tagkey(child.tag)
                d = cur.setdefault(key, {})
                d.update(child.attrib)
                if child.tag == "position" and "kp" in child.attrib:
                    d["kp"] = child.attrib["kp"]

    def resolve(self, tag: str, attrib: Dict[str, str], childclass: Optional[str]) -> Dict[str, str]:
        key = self._tagkey(tag)
        cls = attrib.get("class", childclass or "main")
        out = dict(BUILTIN.get(key, {}))
        out.update(self.classes.get(cls, self.classes["main"]).get(key, {}))
        out.update({k: v for k, v in attrib.items() if k != "class"})
        return out


def _orientation(a: Dict[str, str], compiler: Dict[str, str]) -> np.ndarray:
    if "quat" in a:
        q = _floats(a["quat"])
        return q / np.linalg.norm(q)
    if "euler" in a:
        e = _floats(a["euler"])
        if compiler.get("angle", "degree") == "degree":
            e = np.deg2rad(e)
        return euler_to_quat(e, compiler.get("eulerseq", "xyz"))
    if "axisangle" in a:
        v = _floats(a["axisangle"])
        ang = v[3] if compiler.get("angle", "degree") == "radian" else math.radians(v[3])
        ax = v[:3] / np.linalg.norm(v[:3])
        return np.concatenate([[math.cos(ang / 2)], math.sin(ang / 2) * ax])
    return np.array([1.0, 0, 0, 0])


def _sol(a: Dict[str, str], key: str, default: Sequence[float]) -> np.ndarray:
    if key in a:
        v = list(_floats(a[key]))
        return np.array(v + list(default[len(v):]))
    return np.array(default, dtype=np.float64)


def load(path: str, name: Optional[str] = None,
         base_defaults: Optional[Dict[str, Dict[str, str]]] = None) -> Model:
    """Parses one MJCF file into a resolved `Model` (names are NOT yet prefixed)."""
    path = str(path)
    root = ET.parse(path).getroot()
    base_dir = os.path.dirname(path)
    _expand_includes(root, base_dir)

    compiler: Dict[str, str] = {}
    for c in root.findall("compiler"):
        compiler.update(c.attrib)
    angle_scale = 1.0 if compiler.get("angle", "degree") == "radian" else math.pi / 180.0
    meshdir = os.path.normpath(os.path.join(base_dir, compiler.get("meshdir", "")))

    defaults = _Defaults(base_defaults)
    dels = root.findall("default")
    merged = ET.Element("default")
    for d in dels:
        merged.extend(list(d))
    defaults.load(merged)

    option: Dict[str, str] = {}
    for o in root.findall("option"):
        option.update(o.attrib)

    meshes: Dict[str, MeshAsset] = {}
    for asset in root.findall("asset"):
        for m in asset.findall("mesh"):
            a = defaults.resolve("mesh", m.attrib, None)
            mname = a.get("name") or os.path.splitext(os.path.basename(a["file"]))[0]
            meshes[mname] = MeshAsset(mname, os.path.join(meshdir, a["file"]), _floats(a["scale"]))

    anon = [0]

    def parse_body(elem: ET.Element, childclass: Optional[str], is_world: bool = False) -> Body:
        cc = elem.attrib.get("childclass", childclass)
        inertial = None
        ie = elem.find("inertial")
        if ie is not None:
            ia = ie.attrib
            if "fullinertia" in ia:
                raise NotImplementedError("fullinertia")
            inertial = (
                _floats(ia.get("pos", "0 0 0")), _orientation(ia, compiler),
                float(ia["mass"]), _floats(ia.get("diaginertia", "0 0 0")),
            )
        body = Body(
            name="world" if is_world else elem.attrib.get("name", ""),
            pos=_floats(elem.attrib.get("pos", "0 0 0")),
            quat=_orientation(elem.attrib, compiler),
            inertial=inertial,
        )
        for child in elem:
            if child.tag in ("joint", "freejoint"):
                a = defaults.resolve("joint", child.attrib, cc)
                jtype = "free" if child.tag == "freejoint" else a["type"]
                rng = _floats(a["range"])
                if jtype in ("hinge",):
                    rng = rng * angle_scale
                ax = _floats(a["axis"])
                ax = ax / np.linalg.norm(ax)
                if jtype == "free":
                    a = dict(a, limited="false", armature=a.get("armature", "0"))
                body.joints.append(Joint(
                    name=a.get("name", ""), type=jtype, pos=_floats(a["pos"]), axis=ax,
                    limited=_bool(a["limited"]), range=rng, armature=float(a["armature"]),
                    damping=float(a["damping"]), frictionloss=float(a["frictionloss"]),
                    margin=float(a["margin"]),
                    solreflimit=_sol(a, "solreflimit", _DEFAULT_SOLREF),
                    solimplimit=_sol(a, "solimplimit", _DEFAULT_SOLIMP),
                    solreffriction=_sol(a, "solreffriction", _DEFAULT_SOLREF),
                    solimpfriction=_sol(a, "solimpfriction", _DEFAULT_SOLIMP),
                    ref=float(a["ref"]),
                ))
            elif child.tag == "geom":
                a = defaults.resolve("geom", child.attrib, cc)
                gtype = a["type"]
                if "mesh" in a and "type" not in child.attrib and a["type"] == "sphere":
                    gtype = "mesh"
                if "fromto" in a:
                    ft = _floats(a["fromto"])
                    p0, p1 = ft[:3], ft[3:]
                    pos = 0.5 * (p0 + p1)
                    z = p1 - p0
                    length = np.linalg.norm(z)
                    z = z / length
                    # minimal rotation taking +z to the segment direction (MuJoCo quatZ2Vec)
                    c = np.cross([0, 0, 1.0], z)
                    s = np.linalg.norm(c)
                    if s < 1e-10:
                        quat = np.array([1.0, 0, 0, 0]) if z[2] > 0 else np.array([0.0, 1, 0, 0])
                    else:
                        ang = math.atan2(s, z[2])
                        quat = np.concatenate([[math.cos(ang / 2)], math.sin(ang / 2) * c / s])
                    sz = list(_floats(a["size"]))
                    size = np.array([sz[0], length / 2, 0.0])
                else:
                    pos = _floats(a.get("pos", "0 0 0"))
                    quat = _orientation(a, compiler)
                    sz = list(_floats(a["size"]))
                    size = np.array((sz + [0.0, 0.0, 0.0])[:3])
                gname = a.get("name")
                if not gname:
                    gname = f"//unnamed_geom_{anon[0]}"
                    anon[0] += 1
                body.geoms.append(Geom(
                    name=gname, type=gtype, pos=pos, quat=quat, size=size,
                    contype=int(a["contype"]), conaffinity=int(a["conaffinity"]),
                    condim=int(a["condim"]), priority=int(a["priority"]),
                    friction=_sol(a, "friction", (1.0, 0.005, 0.0001)),
                    solref=_sol(a, "solref", _DEFAULT_SOLREF),
                    solimp=_sol(a, "solimp", _DEFAULT_SOLIMP),
                    solmix=float(a["solmix"]), margin=float(a["margin"]), gap=float(a["gap"]),
                    density=float(a["density"]),
                    mass=float(a["mass"]) if "mass" in a else None,
                    mesh=a.get("mesh"), group=int(a["group"]),
                ))
            elif child.tag == "site":
                a = defaults.resolve("site", child.attrib, cc)
                body.sites.append(Site(a.get("name", ""), _floats(a["pos"]), _orientation(a, compiler)))
            elif child.tag == "body":
                body.children.append(parse_body(child, cc))
        return body

    wb = ET.Element("worldbody")
    for w in root.findall("worldbody"):
        for k, v in w.attrib.items():
            wb.set(k, v)
        wb.extend(list(w))
    world = parse_body(wb, None, is_world=True)

    tendons: List[Tendon] = []
    for te in root.findall("tendon"):
        for f in te.findall("fixed"):
            a = defaults.resolve("tendon", f.attrib, None)
            tendons.append(Tendon(
                name=a["name"],
                joints=[(j.attrib["joint"], float(j.attrib["coef"])) for j in f.findall("joint")],
                limited=_bool(a["limited"]), range=_floats(a["range"]), margin=float(a["margin"]),
                solreflimit=_sol(a, "solreflimit", _DEFAULT_SOLREF),
                solimplimit=_sol(a, "solimplimit", _DEFAULT_SOLIMP),
            ))

    actuators: List[Actuator] = []
    for ae in root.findall("actuator"):
        for act in ae:
            if act.tag not in _ACTUATOR_TAGS:
                raise NotImplementedError(f"actuator <{act.tag}>")
            a = defaults.resolve(act.tag, act.attrib, None)
            gain = np.zeros(3)
            bias = np.zeros(3)
            if act.tag == "position":
                kp = float(a.get("kp", "1"))
                gain[0] = kp
                bias[1] = -kp
            elif act.tag == "velocity":
                kv = float(a.get("kv", "1"))
                gain[0] = kv
                bias[2] = -kv
            elif act.tag == "motor":
                gain[0] = 1.0
            else:
                g = list(_floats(a.get("gainprm", "1 0 0")))
                gain[: min(3, len(g))] = g[:3]
                b = list(_floats(a.get("biasprm", "0 0 0")))
                if a.get("biastype", "none") == "affine":
                    bias[: min(3, len(b))] = b[:3]
                if a.get("gaintype", "fixed") != "fixed" or a.get("dyntype", "none") != "none":
                    raise NotImplementedError("actuator gain/dyn type")
            actuators.append(Actuator(
                name=a["name"], joint=a.get("joint"), tendon=a.get("tendon"),
                gear=_floats(a["gear"])[0], gainprm=gain, biasprm=bias,
                ctrllimited=_bool(a["ctrllimited"]), ctrlrange=_floats(a["ctrlrange"]),
                forcelimited=_bool(a["forcelimited"]), forcerange=_floats(a["forcerange"]),
            ))

    excludes: List[Tuple[str, str]] = []
    pairs: List[ContactPair] = []
    for ce in root.findall("contact"):
        for e in ce.findall("exclude"):
            excludes.append((e.attrib["body1"], e.attrib["body2"]))
        for p in ce.findall("pair"):
            a = p.attrib
            pairs.append(ContactPair(
                geom1=a["geom1"], geom2=a["geom2"], condim=int(a.get("condim", "3")),
                friction=_floats(a["friction"]) if "friction" in a else None,
                solref=_floats(a["solref"]) if "solref" in a else None,
                solimp=_floats(a["solimp"]) if "solimp" in a else None,
                margin=float(a["margin"]) if "margin" in a else None,
                gap=float(a["gap"]) if "gap" in a else None,
            ))

    return Model(
        name=name or root.attrib.get("model", ""), world=world, meshes=meshes,
        tendons=tendons, actuators=actuators, excludes=excludes, pairs=pairs,
        option=option, root_defaults=defaults.classes["main"],
    )


# --------------------------------------------------------------------------- #
# Scene composition (the PyMJCF `attach` subset the tasks use)
# --------------------------------------------------------------------------- #
def _prefix_model(m: Model, prefix: str) -> None:
    def fix(n: str) -> str:
        return n if n.startswith("//") else prefix + n

    for b in m.world.walk():
        if b is not m.world:
            b.name = prefix + b.name
        for j in b.joints:
            j.name = fix(j.name)
        for g in b.geoms:
            g.name = prefix + g.name[2:] if g.name.startswith("//") else prefix + g.name
            if g.mesh:
                g.mesh = prefix + g.mesh
        for s in b.sites:
            s.name = fix(s.name)
    m.meshes = {prefix + k: dataclasses.replace(v, name=prefix + k) for k, v in m.meshes.items()}
    for t in m.tendons:
        t.name = prefix + t.name
        t.joints = [(prefix + j, c) for j, c in t.joints]
    for a in m.actuators:
        a.name = prefix + a.name
        a.joint = prefix + a.joint if a.joint else None
        a.tendon = prefix + a.tendon if a.tendon else None
    m.excludes = [(prefix + a, prefix + b) for a, b in m.excludes]
    for p in m.pairs:
        p.geom1 = prefix + p.geom1
        p.geom2 = prefix + p.geom2


class Scene:
    """A world to which models are attached (mirror of reference Arena.attach_offset /
    add_free_entity, dexterity/models/arenas/arena.py:40-56)."""

    def __init__(self, arena: Model):
        self.model = arena

    @property
    def root_defaults(self):
        return self.model.root_defaults

    def attach(self, child: Model, pos=(0, 0, 0), quat=(1, 0, 0, 0), free: bool = False) -> Body:
        """Wraps `child`'s worldbody contents in a frame body named '<model>/'."""
        prefix = child.name + "/"
        _prefix_model(child, prefix)
        frame = Body(name=prefix, pos=np.array(pos, dtype=np.float64),
                     quat=np.array(quat, dtype=np.float64), inertial=None)
        frame.children = child.world.children
        frame.geoms = child.world.geoms
        frame.sites = child.world.sites
        if free:
            frame.joints.append(Joint(
                name=prefix, type="free", pos=np.zeros(3), axis=np.array([0, 0, 1.0]),
                limited=False, range=np.zeros(2), armature=0.0, damping=0.0, frictionloss=0.0,
                margin=0.0, solreflimit=np.array(_DEFAULT_SOLREF), solimplimit=np.array(_DEFAULT_SOLIMP),
                solreffriction=np.array(_DEFAULT_SOLREF), solimpfriction=np.array(_DEFAULT_SOLIMP),
            ))
        self.model.world.children.append(frame)
        self.model.meshes.update(child.meshes)
        self.model.tendons += child.tendons
        self.model.actuators += child.actuators
        self.model.excludes += child.excludes
        self.model.pairs += child.pairs
        return frame


def primitive_model(name: str, geom_type: str, size: Sequence[float],
                    base_defaults: Dict[str, Dict[str, str]],
                    contype: int = 1, conaffinity: int = 1) -> Model:
    """dm_control `primitive.Primitive`: one geom named 'geom' in the model's worldbody."""
    a = dict(BUILTIN["geom"])
    a.update(base_defaults.get("geom", {}))
    g = Geom(
        name="geom", type=geom_type, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
        size=np.array((list(size) + [0, 0, 0])[:3], dtype=np.float64),
        contype=contype, conaffinity=conaffinity, condim=int(a["condim"]),
        priority=int(a["priority"]), friction=_sol(a, "friction", (1.0, 0.005, 0.0001)),
        solref=_sol(a, "solref", _DEFAULT_SOLREF), solimp=_sol(a, "solimp", _DEFAULT_SOLIMP),
        solmix=float(a["solmix"]), margin=float(a["margin"]), gap=float(a["gap"]),
        density=float(a["density"]), mass=None, mesh=None,
    )
    world = Body(name="world", pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), inertial=None, geoms=[g])
    return Model(name=name, world=world, meshes={}, tendons=[], actuators=[], excludes=[], pairs=[],
                 option={}, root_defaults=base_defaults)
