"""MJCF export of an assembled scene (reference scripts/export_task.py:31-46 exports the task's
MJCF plus its assets with `mjcf.export_with_assets`; this is the same artefact for the scenes this
build assembles itself).

The scene tree (`mjcf.Model`, every default already resolved into the elements) is written as ONE
flat MJCF file with explicit attributes and no `<default>` section, next to the mesh files it names.
Two uses: (1) the zip loads in stock MuJoCo, so a machine that has it can step exactly the scene
this build compiled and compare (tools/mujoco_golden.py covers the reference's own assembly; this
covers ours); (2) the round trip export -> `mjcf.load` -> `compile_model` must reproduce the
committed device blob bit for bit, which pins the exporter and the parser against each other
(tests/test_export_cpu.py).

Not expressible in MJCF and therefore not exported: the constant `xfrc_applied` gravity
compensation the reference hands apply at episode start (`compile_model(xfrc_gravcomp_prefix=)`).
"""
from __future__ import annotations

import os
import zipfile
from typing import Dict, Optional
from xml.sax.saxutils import quoteattr

import numpy as np

from dexterity_b200.model import mjcf


def _f(x) -> str:
    """Shortest decimal strings that parse back to the same float64 values."""
    a = np.atleast_1d(np.asarray(x, dtype=np.float64))
    return " ".join(repr(float(v)) for v in a)


def _b(x) -> str:
    return "true" if x else "false"


def _attrs(**kw) -> str:
    return "".join(f" {k}={quoteattr(str(v))}" for k, v in kw.items() if v is not None)


def _geom_xml(g: mjcf.Geom) -> str:
    kw = dict(name=g.name, type=g.type, pos=_f(g.pos), quat=_f(g.quat), contype=g.contype, conaffinity=g.conaffinity,
              condim=g.condim, priority=g.priority, friction=_f(g.friction), solref=_f(g.solref), solimp=_f(g.solimp),
              solmix=_f(g.solmix), margin=_f(g.margin), gap=_f(g.gap), group=g.group)
    if g.type == "mesh":
        kw["mesh"] = g.mesh
    else:
        kw["size"] = _f(g.size)
    if g.mass is not None:
        kw["mass"] = _f(g.mass)
    else:
        kw["density"] = _f(g.density)
    return f"<geom{_attrs(**kw)}/>"


def _joint_xml(j: mjcf.Joint) -> str:
    if j.type == "free":
        return f"<joint{_attrs(name=j.name, type='free')}/>"
    return "<joint" + _attrs(
        name=j.name, type=j.type, pos=_f(j.pos), axis=_f(j.axis), limited=_b(j.limited), range=_f(j.range),
        armature=_f(j.armature), damping=_f(j.damping), frictionloss=_f(j.frictionloss), margin=_f(j.margin),
        solreflimit=_f(j.solreflimit), solimplimit=_f(j.solimplimit), solreffriction=_f(j.solreffriction),
        solimpfriction=_f(j.solimpfriction), ref=_f(j.ref)) + "/>"


def _body_xml(b: mjcf.Body, depth: int, out: list) -> None:
    pad = "  " * depth
    out.append(f"{pad}<body{_attrs(name=b.name, pos=_f(b.pos), quat=_f(b.quat))}>")
    if b.inertial is not None:
        ipos, iquat, mass, diag = b.inertial
        out.append(f"{pad}  <inertial{_attrs(pos=_f(ipos), quat=_f(iquat), mass=_f(mass), diaginertia=_f(diag))}/>")
    for j in b.joints:
        out.append(pad + "  " + _joint_xml(j))
    for g in b.geoms:
        out.append(pad + "  " + _geom_xml(g))
    for s in b.sites:
        out.append(f"{pad}  <site{_attrs(name=s.name, pos=_f(s.pos), quat=_f(s.quat))}/>")
    for c in b.children:
        _body_xml(c, depth + 1, out)
    out.append(f"{pad}</body>")


def to_xml(model: mjcf.Model, timestep: Optional[float] = None, mesh_files: Optional[Dict[str, str]] = None) -> str:
    """Serialises `model`.  `mesh_files` maps mesh names to the file names written into the XML
    (default: the base name of each asset's source file)."""
    opt = dict(model.option)
    if timestep is not None:
        opt["timestep"] = repr(float(timestep))
    out = [f"<mujoco{_attrs(model=model.name)}>",
           '  <compiler angle="radian" meshdir="assets"/>',
           f"  <option{_attrs(**opt)}/>"]
    used = {g.mesh for b in model.world.walk() for g in b.geoms if g.type == "mesh"}
    if used:
        out.append("  <asset>")
        for name, m in model.meshes.items():
            if name in used:
                fn = (mesh_files or {}).get(name, os.path.basename(m.file))
                out.append(f"    <mesh{_attrs(name=name, file=fn, scale=_f(m.scale))}/>")
        out.append("  </asset>")
    out.append("  <worldbody>")
    w = model.world
    for g in w.geoms:
        out.append("    " + _geom_xml(g))
    for s in w.sites:
        out.append(f"    <site{_attrs(name=s.name, pos=_f(s.pos), quat=_f(s.quat))}/>")
    for c in w.children:
        _body_xml(c, 2, out)
    out.append("  </worldbody>")
    if model.excludes or model.pairs:
        out.append("  <contact>")
        for a, b in model.excludes:
            out.append(f"    <exclude{_attrs(body1=a, body2=b)}/>")
        for p in model.pairs:
            out.append("    <pair" + _attrs(
                geom1=p.geom1, geom2=p.geom2, condim=p.condim,
                friction=None if p.friction is None else _f(p.friction), solref=None if p.solref is None else _f(p.solref),
                solimp=None if p.solimp is None else _f(p.solimp), margin=None if p.margin is None else _f(p.margin),
                gap=None if p.gap is None else _f(p.gap)) + "/>")
        out.append("  </contact>")
    if model.tendons:
        out.append("  <tendon>")
        for t in model.tendons:
            out.append("    <fixed" + _attrs(name=t.name, limited=_b(t.limited), range=_f(t.range), margin=_f(t.margin),
                                            solreflimit=_f(t.solreflimit), solimplimit=_f(t.solimplimit)) + ">")
            for jn, c in t.joints:
                out.append(f"      <joint{_attrs(joint=jn, coef=_f(c))}/>")
            out.append("    </fixed>")
        out.append("  </tendon>")
    if model.actuators:
        out.append("  <actuator>")
        for a in model.actuators:
            out.append("    <general" + _attrs(
                name=a.name, joint=a.joint, tendon=a.tendon, gear=_f(a.gear), gaintype="fixed", biastype="affine",
                gainprm=_f(a.gainprm), biasprm=_f(a.biasprm), ctrllimited=_b(a.ctrllimited), ctrlrange=_f(a.ctrlrange),
                forcelimited=_b(a.forcelimited), forcerange=_f(a.forcerange)) + "/>")
        out.append("  </actuator>")
    out.append("</mujoco>")
    return "\n".join(out) + "\n"


def export_with_assets(model: mjcf.Model, out_dir: str, timestep: Optional[float] = None, zip_path: Optional[str] = None) -> str:
    """Writes `<out_dir>/<model.name>.xml` plus `<out_dir>/assets/<mesh files>`; returns the XML path.  With `zip_path`
    the same files are also packed into a zip (the layout of reference scripts/export_task.py:40-46)."""
    os.makedirs(os.path.join(out_dir, "assets"), exist_ok=True)
    used = {g.mesh for b in model.world.walk() for g in b.geoms if g.type == "mesh"}
    names: Dict[str, str] = {}
    taken = set()
    for name, m in model.meshes.items():
        if name not in used:
            continue
        fn = os.path.basename(m.file)
        if fn in taken:                       # same base name in different directories
            fn = name.replace("/", "_") + os.path.splitext(fn)[1]
        taken.add(fn)
        names[name] = fn
        with open(m.file, "rb") as src, open(os.path.join(out_dir, "assets", fn), "wb") as dst:
            dst.write(src.read())
    xml_path = os.path.join(out_dir, model.name.replace("/", "_") + ".xml")
    with open(xml_path, "w") as f:
        f.write(to_xml(model, timestep, names))
    if zip_path:
        with zipfile.ZipFile(zip_path, "w", zipfile.ZIP_DEFLATED) as z:
            z.write(xml_path, os.path.basename(xml_path))
            for fn in sorted(taken):
                z.write(os.path.join(out_dir, "assets", fn), os.path.join("assets", fn))
    return xml_path
