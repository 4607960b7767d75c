"""Scene builders: the MJCF assembly the reference tasks perform, done once on the host.

* reorient: reference manipulation/tasks/reorient.py:324-364 (`reorient_task`) +
  ReOrient.__init__ :113-168 — Standard arena, Shadow hand attached at its
  palm-upright pose, free OpenAI cube (half-size 0.02), contact-less hint cube.
* reach (Shadow): reference manipulation/tasks/reach.py:91-139 with the Shadow hand
  as BASELINE.json configs[0..1] state; ground collisions disabled (:131-132).
* hand_only: the model `IKSolver` builds its private physics from
  (reference inverse_kinematics/ik_solver.py:46).

The vendored MJCF/STL model data is read from the reference checkout
(`DEXTERITY_REFERENCE_ROOT`, default /root/reference) at *compile* time only; the
compiled blobs are committed under dexterity_b200/assets/ (tools/build_assets.py)
so nothing at run time needs the reference tree.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from dexterity_b200.model import compile as compile_lib
from dexterity_b200.model import mjcf

ASSET_DIR = os.path.join(os.path.dirname(__file__), "..", "assets")

# reference models/hands/shadow_hand_e.py:13-14
SHADOW_PALM_UPRIGHT_POS = (0.0, 0.2, 0.1)
SHADOW_PALM_UPRIGHT_QUAT = (0.0, 0.0, 0.707106781186, -0.707106781186)
SHADOW_FINGERTIPS = ("fftip", "mftip", "rftip", "lftip", "thtip")

# reference manipulation/tasks/reorient.py:41-78
REORIENT_PROP_SIZE = 0.02
REORIENT_HINT_POS = (0.12, 0.0, 0.15)
REORIENT_PHYSICS_TIMESTEP = 0.005
REORIENT_CONTROL_TIMESTEP = 0.025
# reference manipulation/tasks/reach.py:54,59
REACH_PHYSICS_TIMESTEP = 0.02
REACH_CONTROL_TIMESTEP = 0.02


def reference_root() -> str:
    return os.environ.get("DEXTERITY_REFERENCE_ROOT", "/root/reference")


def _paths():
    root = os.path.join(reference_root(), "dexterity", "models")
    return dict(
        arena=os.path.join(root, "arenas", "arena.xml"),
        shadow=os.path.join(root, "vendor", "shadow_robot", "shadow_hand_description", "mjcf",
                            "shadow_hand_series_e.xml"),
        adroit=os.path.join(root, "vendor", "adroit", "adroit_hand_description", "mjcf", "adroit_hand.xml"),
    )


def _standard_arena(ground_collides: bool = True) -> mjcf.Scene:
    """reference models/arenas/arena.py:15-20 + standard.py:13-22."""
    arena = mjcf.load(_paths()["arena"], name="arena")
    ground = mjcf.Geom(
        name="ground", type="plane", pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
        size=np.array([1.0, 1.0, 0.1]),
        contype=1 if ground_collides else 0, conaffinity=1 if ground_collides else 0,
        condim=3, priority=0, friction=np.array([0.4, 0.005, 0.0001]),
        solref=np.array([0.002, 1.0]), solimp=np.array([0.95, 0.99, 0.001, 0.5, 2.0]),
        solmix=1.0, margin=0.0, gap=0.0, density=1000.0, mass=None, mesh=None,
    )
    arena.world.geoms.append(ground)
    return mjcf.Scene(arena)


def _shadow_hand(scene: mjcf.Scene, name: str = "shadow_hand_e") -> mjcf.Model:
    hand = mjcf.load(_paths()["shadow"], name=name, base_defaults=scene.root_defaults)
    # fingertip sites, reference models/hands/shadow_hand_e.py:151-169
    for tip in SHADOW_FINGERTIPS:
        hand.find_body(tip).sites.append(
            mjcf.Site(name=tip + "_site", pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0])))
    return hand


def reorient_assembly():
    """The assembled (not yet compiled) reorient scene and the arguments of its compile step."""
    scene = _standard_arena()
    hand = _shadow_hand(scene)
    scene.attach(hand, pos=SHADOW_PALM_UPRIGHT_POS, quat=SHADOW_PALM_UPRIGHT_QUAT)
    prop = mjcf.primitive_model("prop", "box", [REORIENT_PROP_SIZE] * 3, scene.root_defaults)
    scene.attach(prop, free=True)
    # _hintify (reorient.py:294-321): contype = conaffinity = 0, fixed to the world
    hint = mjcf.primitive_model("target_prop", "box", [REORIENT_PROP_SIZE] * 3, scene.root_defaults,
                                contype=0, conaffinity=0)
    scene.attach(hint, pos=REORIENT_HINT_POS)
    return scene.model, dict(timestep=REORIENT_PHYSICS_TIMESTEP, nconmax=48, xfrc_gravcomp_prefix="shadow_hand_e/")


def reorient_scene() -> compile_lib.CompiledModel:
    model, kw = reorient_assembly()
    return compile_lib.compile_model(model, **kw)


def reach_shadow_assembly():
    scene = _standard_arena(ground_collides=False)
    hand = _shadow_hand(scene)
    scene.attach(hand, pos=SHADOW_PALM_UPRIGHT_POS, quat=SHADOW_PALM_UPRIGHT_QUAT)
    return scene.model, dict(timestep=REACH_PHYSICS_TIMESTEP, nconmax=48, xfrc_gravcomp_prefix="shadow_hand_e/")


def reach_shadow_scene() -> compile_lib.CompiledModel:
    model, kw = reach_shadow_assembly()
    return compile_lib.compile_model(model, **kw)


def reach_adroit_assembly():
    """The REGISTERED reach task's scene (reference manipulation/tasks/reach.py:223-269 with `AdroitHand`,
    models/hands/adroit_hand.py:12-29): Standard arena with the ground's collisions disabled, the Adroit hand attached at
    its palm-upright pose.  The hand's collision geoms have conaffinity 0 (adroit_hand.xml:29-31), so the only contacts are
    its 18 distinct explicit `<contact><pair condim="1">` elements (:231-251): 17 capsule-capsule, 1 capsule-box."""
    scene = _standard_arena(ground_collides=False)
    hand = mjcf.load(_paths()["adroit"], name="adroit_hand", base_defaults=scene.root_defaults)
    scene.attach(hand, pos=SHADOW_PALM_UPRIGHT_POS, quat=SHADOW_PALM_UPRIGHT_QUAT)   # adroit_hand.py:12-13: the same pose
    return scene.model, dict(timestep=REACH_PHYSICS_TIMESTEP, nconmax=48, xfrc_gravcomp_prefix="adroit_hand/")


def reach_adroit_scene() -> compile_lib.CompiledModel:
    model, kw = reach_adroit_assembly()
    return compile_lib.compile_model(model, **kw)


def shadow_hand_only() -> compile_lib.CompiledModel:
    """The hand's own root model (no arena, no attachment offset): IKSolver's physics."""
    scene = mjcf.Scene(mjcf.Model("ik", mjcf.Body("world", np.zeros(3), np.array([1.0, 0, 0, 0]), None),
                                  {}, [], [], [], [], {}, {}))
    hand = mjcf.load(_paths()["shadow"], name="shadow_hand_e")
    for tip in SHADOW_FINGERTIPS:
        hand.find_body(tip).sites.append(
            mjcf.Site(name=tip + "_site", pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0])))
    # the IK physics is the un-attached model: bodies live directly in the world (names keep the
    # hand prefix so that site / joint lookups are uniform across scenes)
    mjcf._prefix_model(hand, "shadow_hand_e/")
    scene.model.world.children = hand.world.children
    scene.model.meshes = hand.meshes
    scene.model.tendons, scene.model.actuators = hand.tendons, hand.actuators
    scene.model.excludes = hand.excludes
    scene.model.option = hand.option
    return compile_lib.compile_model(scene.model, timestep=0.002, nconmax=48)


# reference models/hands/adroit_hand_constants.py:86-92 (sites that exist in adroit_hand.xml)
ADROIT_FINGERTIP_SITES = ("S_fftip", "S_mftip", "S_rftip", "S_lftip", "S_thtip")


def adroit_hand_only() -> compile_lib.CompiledModel:
    """The Adroit hand's own root model (reference models/hands/adroit_hand.py:17-29): the physics of
    `IKSolver(AdroitHand())` and of the golden fingertip test reference models/hands/hands_test.py:195-228.
    KINEMATICS ONLY: the 19 explicit `<contact><pair>` elements of adroit_hand.xml:231-251 (capsule-capsule and
    box-capsule, condim 1) are not compiled (no such narrowphase in this build yet), which IK does not need; the 44
    fixed tendons are neither limited nor actuated (the `limited` default sits in a class no tendon uses) and are dropped."""
    scene = mjcf.Scene(mjcf.Model("ik", mjcf.Body("world", np.zeros(3), np.array([1.0, 0, 0, 0]), None),
                                  {}, [], [], [], [], {}, {}))
    hand = mjcf.load(_paths()["adroit"], name="adroit_hand")
    mjcf._prefix_model(hand, "adroit_hand/")
    scene.model.world.children = hand.world.children
    scene.model.meshes = hand.meshes
    scene.model.tendons, scene.model.actuators = hand.tendons, hand.actuators
    scene.model.excludes = hand.excludes
    scene.model.option = hand.option
    return compile_lib.compile_model(scene.model, timestep=0.002, nconmax=48)


_BUILDERS = {
    "reorient_shadow": reorient_scene,
    "reach_shadow": reach_shadow_scene,
    "reach_adroit": reach_adroit_scene,
    "shadow_hand_only": shadow_hand_only,
    "adroit_hand_only": adroit_hand_only,
}

# scenes whose assembled MJCF can be exported (dexterity_b200/model/export.py, tools/export_task.py)
ASSEMBLIES = {"reorient_shadow": reorient_assembly, "reach_shadow": reach_shadow_assembly, "reach_adroit": reach_adroit_assembly}

# hand name -> (hand-only compiled model, name prefix, fingertip site names)
HANDS = {
    "shadow_hand_e": ("shadow_hand_only", "shadow_hand_e/", tuple(t + "_site" for t in SHADOW_FINGERTIPS)),
    "adroit_hand": ("adroit_hand_only", "adroit_hand/", ADROIT_FINGERTIP_SITES),
}


def load_compiled(name: str, rebuild: bool = False) -> compile_lib.CompiledModel:
    """Loads the committed compiled blob; recompiles from the reference MJCF when asked
    (or when the asset is missing and the reference checkout is present)."""
    path = os.path.join(ASSET_DIR, name + ".npz")
    if not rebuild and os.path.exists(path):
        return compile_lib.CompiledModel.load(path)
    if not os.path.isdir(reference_root()):
        raise FileNotFoundError(
            f"compiled model {path} missing and reference MJCF not found at {reference_root()}")
    model = _BUILDERS[name]()
    os.makedirs(ASSET_DIR, exist_ok=True)
    model.save(path)
    return model
