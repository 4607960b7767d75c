"""Batched, composer-compatible environments (leading env dimension on every array).

Mirrors the reference's public surface for the hot path:
  * `load(domain_name, task_name, seed, ..., num_envs=...)`   ← dexterity/manipulation/__init__.py:56-86
  * `GoalEnvironment.reset() / step(action)`                  ← dexterity/environment.py:14-34
  * `action_spec()` (bounds = ctrlrange where ctrllimited)    ← effectors/mujoco_actuation.py:48-67
  * `observation_spec()` / observation dict keys              ← composer observables enabled by
    shared/observations.py:96-103, reorient.py:81-86, task.py:206-216

dm_env semantics are kept per env: the TimeStep after a terminal step has step_type LAST; the
next `step` call ignores that env's action and returns the FIRST TimeStep of a new episode.
"""
from __future__ import annotations

import collections
import ctypes
import dataclasses
import math
from typing import Dict, NamedTuple, Optional

import numpy as np
import torch

from dexterity_b200 import backend
from dexterity_b200 import placement
from dexterity_b200.model import scenes

FIRST, MID, LAST = 0, 1, 2
OBS_STRIDE, GOAL_STRIDE = 128, 16

# reference manipulation/tasks/reorient.py:41-78
REORIENT = dict(
    success_threshold=0.1, successes_needed=1, steps_before_changing_goal=5,
    max_time_per_goal=300 * 0.025, control_timestep=0.025, physics_timestep=0.005,
    spawn_lo=(-0.025, -0.13 - 0.025, 0.16), spawn_hi=(0.025, -0.13 + 0.025, 0.16),
    w_orientation=1.0, w_success_bonus=800.0, w_action_smoothing=-0.1, orientation_eps=0.1, bonus_threshold=0.1)
# reference manipulation/tasks/reach.py:47-66
REACH = dict(
    success_threshold=0.01, successes_needed=50, steps_before_changing_goal=5,
    max_time_per_goal=150 * 0.02, control_timestep=0.02, physics_timestep=0.02,
    # reach.py:196-205: rewards.tanh_squared(d, margin=0.1) with the default loss_at_margin=0.95 (shared/rewards.py:18-28)
    reward_margin=0.1, reward_loss_at_margin=0.95)


class TaskParams(ctypes.Structure):
    """struct DexTaskParams (include/dexterity_b200.h)."""
    _fields_ = [
        ("task", ctypes.c_int), ("successes_needed", ctypes.c_int), ("steps_before_changing_goal", ctypes.c_int),
        ("fall_termination", ctypes.c_int), ("has_max_time", ctypes.c_int),
        ("prop_body", ctypes.c_int), ("prop_qposadr", ctypes.c_int), ("prop_dofadr", ctypes.c_int),
        ("prop_geom", ctypes.c_int), ("ground_geom", ctypes.c_int),
        ("nhand", ctypes.c_int), ("nsite", ctypes.c_int), ("site_id", ctypes.c_int * 5),
        ("dense_reward", ctypes.c_int), ("place_max_attempts", ctypes.c_int),
        ("success_threshold", ctypes.c_double), ("max_time_per_goal", ctypes.c_double), ("time_limit", ctypes.c_double),
        ("spawn_lo", ctypes.c_double * 3), ("spawn_hi", ctypes.c_double * 3),
        ("w_orientation", ctypes.c_double), ("w_success_bonus", ctypes.c_double), ("w_action_smoothing", ctypes.c_double),
        ("orientation_eps", ctypes.c_double), ("bonus_threshold", ctypes.c_double),
        ("reach_tanh_weight", ctypes.c_double),
        ("seed", ctypes.c_ulonglong),
    ]


@dataclasses.dataclass(frozen=True)
class BoundedArraySpec:
    """Stand-in for dm_env.specs.BoundedArray (dm_env is not a dependency of this build)."""
    shape: tuple
    dtype: np.dtype
    minimum: np.ndarray
    maximum: np.ndarray
    name: str


@dataclasses.dataclass(frozen=True)
class ArraySpec:
    shape: tuple
    dtype: np.dtype
    name: str


class TimeStep(NamedTuple):
    step_type: torch.Tensor   # int32 [num_envs]: 0 FIRST, 1 MID, 2 LAST
    reward: torch.Tensor      # [num_envs] (0 where step_type == FIRST; dm_env uses None)
    discount: torch.Tensor    # [num_envs] (1 where step_type == FIRST)
    observation: Dict[str, torch.Tensor]

    def first(self): return self.step_type == FIRST
    def mid(self): return self.step_type == MID
    def last(self): return self.step_type == LAST


def _reorient_layout(prefix: str, nhand: int, nsite: int) -> "collections.OrderedDict[str, slice]":
    o = collections.OrderedDict()
    o[f"{prefix}/joint_positions_sin_cos"] = slice(0, 2 * nhand)
    o[f"{prefix}/joint_velocities"] = slice(2 * nhand, 3 * nhand)
    a = 3 * nhand
    o[f"{prefix}/fingertip_positions"] = slice(a, a + 3 * nsite)
    o[f"{prefix}/fingertip_linear_velocities"] = slice(a + 3 * nsite, a + 6 * nsite)
    return o


class BatchedGoalEnvironment:
    """`num_envs` independent copies of a GoalTask environment on one GPU."""

    def __init__(self, task_name: str, num_envs: int, seed: Optional[int] = None, device: int = 0,
                 dtype: torch.dtype = torch.float32, time_limit: Optional[float] = None,
                 fall_termination: bool = True, dense_reward: bool = True, prop_spawn_bbox=None,
                 place_max_attempts: Optional[int] = None, hand: str = "shadow_hand_e") -> None:
        """`hand`: "shadow_hand_e" or (reach only) "adroit_hand", the hand of the reference's registered reach task.
        `prop_spawn_bbox` = (lower, upper) overrides the reference's spawn box (reorient.py:72-78).
        `place_max_attempts`: draws of the collision-checked prop placement (PropPlacer); None = decide from a static
        clearance proof of the spawn box (dexterity_b200/placement.py): 0 when no draw can collide, else 20."""
        if task_name not in ("reorient", "reach"):
            raise ValueError(f"Unknown task: {task_name}")
        if hand not in scenes.HANDS or (task_name == "reorient" and hand != "shadow_hand_e"):
            raise ValueError(f"Unknown hand for {task_name}: {hand}")
        self.task_name, self.hand_name = task_name, hand
        self.compiled = scenes.load_compiled({("reorient", "shadow_hand_e"): "reorient_shadow", ("reach", "shadow_hand_e"): "reach_shadow",
                                              ("reach", "adroit_hand"): "reach_adroit"}[(task_name, hand)])
        c = self.compiled
        self.model = backend.PhysicsModel(c, device)
        self.physics = backend.PhysicsBatch(self.model, num_envs, dtype)
        self.num_envs, self.dtype, self.device = num_envs, dtype, self.physics.device
        cfg = REORIENT if task_name == "reorient" else REACH
        self.control_timestep, self.physics_timestep = cfg["control_timestep"], cfg["physics_timestep"]
        self.n_sub_steps = int(round(self.control_timestep / self.physics_timestep))
        seed = int(seed if seed is not None else np.random.randint(0, 2 ** 31 - 1))
        self.seed = seed

        n, dev = num_envs, self.device
        z = lambda *s, dt=dtype: torch.zeros(*s, dtype=dt, device=dev)
        self._obs = z(n, OBS_STRIDE)
        self._reward, self._discount = z(n), z(n)
        self._goal = z(n, GOAL_STRIDE)
        self._task_i = z(n, 8, dt=torch.int32)
        self._task_f = z(n, 8)
        self._rng = z(n, 4, dt=torch.int32)
        self._rng[:, 0] = seed & 0x7FFFFFFF
        self._rng[:, 1] = (seed >> 31) & 0x7FFFFFFF
        self._action = z(n, backend.PhysicsBatch.ROW)
        self._skip = z(n, dt=torch.uint8)
        self._reset_mask = z(n, dt=torch.uint8)
        self.episode_stats = z(8, dt=torch.float64)
        for name, t in (("OBS", self._obs), ("REWARD", self._reward), ("DISCOUNT", self._discount), ("GOAL", self._goal),
                        ("TASK_I", self._task_i), ("TASK_F", self._task_f), ("RNG", self._rng), ("ACTION", self._action),
                        ("EPISODE_STATS", self.episode_stats), ("RESET_MASK", self._reset_mask)):
            self.physics.bind(name, t)
        backend._check(backend.lib().dex_bind(self.physics.ptr, 26, ctypes.c_void_p(self._skip.data_ptr()), 1))

        # task description
        hand_prefix = hand
        sites = [c.site_names.index(f"{hand_prefix}/{t}") for t in scenes.HANDS[hand][2]]
        nhand = sum(1 for j in c.joint_names if j.startswith(hand_prefix + "/"))
        p = TaskParams()
        p.task = 0 if task_name == "reorient" else 1
        p.successes_needed = cfg["successes_needed"]
        p.steps_before_changing_goal = cfg["steps_before_changing_goal"]
        p.fall_termination = int(fall_termination)
        p.has_max_time = 1
        p.nhand, p.nsite = nhand, len(sites)
        for i, s in enumerate(sites):
            p.site_id[i] = s
        p.dense_reward = int(dense_reward)
        if task_name == "reach":
            p.reach_tanh_weight = math.atanh(math.sqrt(cfg["reward_loss_at_margin"])) / cfg["reward_margin"]
        p.success_threshold = cfg["success_threshold"]
        p.max_time_per_goal = cfg["max_time_per_goal"]
        p.time_limit = float(time_limit) if time_limit else math.inf
        p.seed = seed
        layout = _reorient_layout(hand_prefix, nhand, len(sites))
        o = 3 * nhand + 6 * len(sites)
        if task_name == "reorient":
            pb = c.body_names.index("prop/")
            j = c.joint_names.index("prop/")
            p.prop_body, p.prop_qposadr, p.prop_dofadr = pb, int(c.JNT_QPOSADR[j]), int(c.JNT_DOFADR[j])
            p.prop_geom, p.ground_geom = c.geom_names.index("prop/geom"), c.geom_names.index("ground")
            spawn_lo, spawn_hi = prop_spawn_bbox if prop_spawn_bbox is not None else (cfg["spawn_lo"], cfg["spawn_hi"])
            for k in range(3):
                p.spawn_lo[k], p.spawn_hi[k] = spawn_lo[k], spawn_hi[k]
            if place_max_attempts is None:
                place_max_attempts = placement.place_max_attempts(c, pb, spawn_lo, spawn_hi)
            p.place_max_attempts = int(place_max_attempts)
            self.place_max_attempts = p.place_max_attempts
            for k in ("w_orientation", "w_success_bonus", "w_action_smoothing", "orientation_eps", "bonus_threshold"):
                setattr(p, k, cfg[k])
            layout["prop/position"] = slice(o, o + 3)
            layout["prop/orientation"] = slice(o + 3, o + 7)
            layout["prop/linear_velocity"] = slice(o + 7, o + 10)
            layout["prop/angular_velocity"] = slice(o + 10, o + 13)
            layout["target_prop/orientation"] = slice(o + 13, o + 17)
            layout["goal_state"] = slice(o + 17, o + 21)
            self.obs_dim = o + 21
        else:
            layout["goal_state"] = slice(o, o + 3 * len(sites))
            self.obs_dim = o + 3 * len(sites)
        self._params = p
        self.obs_layout = layout
        lib = backend.lib()
        lib.dex_task_configure.argtypes = [ctypes.c_void_p, ctypes.POINTER(TaskParams)]
        lib.dex_env_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.dex_env_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        backend._check(lib.dex_task_configure(self.physics.ptr, ctypes.byref(p)))

        is_lim = c.ACT_CTRLLIMITED.astype(bool)
        lo = np.full(c.nu, -np.inf, np.float32)
        hi = np.full(c.nu, np.inf, np.float32)
        lo[is_lim], hi[is_lim] = c.ACT_CTRLRANGE[is_lim, 0], c.ACT_CTRLRANGE[is_lim, 1]
        self._action_spec = BoundedArraySpec((c.nu,), np.dtype(np.float32), lo, hi,
                                             "\t".join(f"{hand_prefix}_joint{i}" for i in range(c.nu)))

    # ------------------------------------------------------------------ specs
    def action_spec(self) -> BoundedArraySpec:
        return self._action_spec

    def observation_spec(self) -> "collections.OrderedDict[str, ArraySpec]":
        return collections.OrderedDict(
            (k, ArraySpec((s.stop - s.start,), np.dtype(np.float64), k)) for k, s in self.obs_layout.items())

    def reward_spec(self) -> ArraySpec:
        return ArraySpec((), np.dtype(np.float64), "reward")

    def discount_spec(self) -> BoundedArraySpec:
        return BoundedArraySpec((), np.dtype(np.float64), np.array(0.0), np.array(1.0), "discount")

    # ------------------------------------------------------------------ dm_env API (batched)
    def _timestep(self) -> TimeStep:
        obs = collections.OrderedDict((k, self._obs[:, s]) for k, s in self.obs_layout.items())
        return TimeStep(self._task_i[:, 6], self._reward, self._discount, obs)

    def reset(self) -> TimeStep:
        backend._check(backend.lib().dex_env_reset(self.physics.ptr, self.physics._stream()))
        return self._timestep()

    def step(self, action: torch.Tensor) -> TimeStep:
        """`action`: [num_envs, nu] tensor on this env's device (float32 or the batch dtype)."""
        self._action[:, :action.shape[1]].copy_(action)
        return self.step_bound()

    def step_bound(self) -> TimeStep:
        """Steps with whatever is already stored in `self.action_buffer` (zero-copy path)."""
        backend._check(backend.lib().dex_env_step(self.physics.ptr, self.n_sub_steps, self.physics._stream()))
        return self._timestep()

    @property
    def action_buffer(self) -> torch.Tensor:
        return self._action[:, :self.compiled.nu]

    @property
    def observation_buffer(self) -> torch.Tensor:
        return self._obs[:, :self.obs_dim]

    @property
    def goal(self) -> torch.Tensor:
        return self._goal

    @property
    def goal_distance(self) -> torch.Tensor:
        return self._task_f[:, 2:7]

    @property
    def successes(self) -> torch.Tensor:
        return self._task_i[:, 0]

    # ------------------------------------------------------------------ hand observables outside HAND_OBSERVABLES
    # The reference defines them (models/hands/dexterous_hand.py:244-350) but no registered task enables them
    # (shared/observations.py:96-103); they are served from the device state on request instead of being packed into the
    # observation vector.  joint_torques (torque sensors, mj_rnePostConstraint) is not built.
    def _hand_qpos_adr(self) -> torch.Tensor:
        c = self.compiled
        ids = [i for i, n in enumerate(c.joint_names) if n.startswith(self.hand_name + "/")]
        return torch.as_tensor(np.asarray(c.JNT_QPOSADR)[ids].astype(np.int64), device=self.device)

    def joint_positions(self) -> torch.Tensor:
        """[num_envs, hand joints]: dexterous_hand.py:249-251 (MJCFFeature "qpos" of the hand's joints)."""
        return self.physics.qpos[:, self._hand_qpos_adr()]

    def _site_frames(self):
        """World positions [n, sites, 3] and rotation matrices [n, sites, 3, 3] of the fingertip sites and the pose of the hand's
        root body, by forward kinematics of the current qpos (torch; the hand is a tree of hinges on a fixed base)."""
        from dexterity_b200 import hands
        c = self.compiled
        hand = getattr(self, "hand", None) or hands.DexterousHand(c, prefix=scenes.HANDS[self.hand_name][1],
                                                                   fingertip_sites=scenes.HANDS[self.hand_name][2])
        root = c.body_names.index(f"{self.hand_name}/forearm")      # root_body (shadow_hand_e.py / adroit_hand.py: find("body", "forearm"))
        return _fk_sites(c, hand, self.physics, frames=True, extra_body=root)

    def fingertip_orientations(self) -> torch.Tensor:
        """[num_envs, 4 * sites] world-frame quaternions of the fingertip sites (dexterous_hand.py:286-295)."""
        _, R, _, _, _ = self._site_frames()
        return _mat_to_quat(R).reshape(self.num_envs, -1)

    def fingertip_angular_velocities(self) -> torch.Tensor:
        """[num_envs, 3 * sites] world-frame angular velocities of the fingertip sites (dexterous_hand.py:311-323:
        mj_objectVelocity of the site, rotated to the world frame): the angular velocity of the site's body."""
        return self._site_frames()[4].reshape(self.num_envs, -1)

    def fingertip_positions_ego(self) -> torch.Tensor:
        """[num_envs, 3 * sites] fingertip positions relative to the hand's root body, in that body's frame: the `framepos`
        sensors with reftype = body, refname = root_body of dexterous_hand.py:325-348."""
        p, _, rpos, rmat, _ = self._site_frames()
        return torch.einsum("nji,nsj->nsi", rmat, p - rpos[:, None, :]).reshape(self.num_envs, -1)

    def request_reset(self, mask: torch.Tensor) -> None:
        """Marks envs to be re-initialised at the next `step` (their action is ignored)."""
        self._reset_mask.copy_(mask.to(torch.uint8))

    def reduce_episode_stats(self) -> Dict[str, float]:
        """Episode statistics summed over all ranks (the ONLY collective on this path: one
        all-reduce of an 8-element float64 vector over NCCL / NVLink per logging interval)."""
        from dexterity_b200 import sharding
        return sharding.reduce_episode_stats(self.episode_stats)


class ReachEnvironment(BatchedGoalEnvironment):
    """Reach task (reference manipulation/tasks/reach.py:71-220) with the Shadow hand.

    Goal sampling and the collision-free initial hand configuration need masked physics
    sub-rollouts, so a step is split into its phases (dex_env_phase) with the host-side
    `FingertipCartesianPosition` generator acting between them; `GoalInitializationError` is retried
    as in reference dexterity/environment.py:14-34.
    """

    _INIT_JOINT_RANGE_FRACTION = 0.5     # reach.py:45

    def __init__(self, num_envs: int, seed=None, device: int = 0, dtype: torch.dtype = torch.float32,
                 time_limit=None, dense_reward: bool = True, hand: str = "adroit_hand") -> None:
        super().__init__("reach", num_envs, seed, device, dtype, time_limit, dense_reward=dense_reward, hand=hand)
        from dexterity_b200 import goals, hands
        self.hand = hands.DexterousHand(self.compiled, prefix=scenes.HANDS[hand][1], fingertip_sites=scenes.HANDS[hand][2])
        self.goal_generator = goals.FingertipCartesianPosition(self.hand, self.model)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(self.seed)
        self.goal_qpos = torch.zeros(num_envs, self.hand.num_joints, dtype=dtype, device=self.device)
        backend.lib().dex_env_phase.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        self._site_body = torch.as_tensor(self.compiled.SITE_BODY[self.hand.fingertip_site_ids], device=self.device)

    def fingertip_positions(self, physics=None) -> torch.Tensor:
        """Fingertip site world positions [n, 15] at the physics' current qpos, from the observation
        refresh kernel when `physics` is this env's own batch, else via a scratch forward pass."""
        if physics is None or physics is self.physics:
            return self._obs[:, self.obs_layout[f"{self.hand_name}/fingertip_positions"]]
        return _fk_sites(self.compiled, self.hand, physics)

    def _phase(self, phases: int, force_reset: int = 0) -> None:
        backend._check(backend.lib().dex_env_phase(self.physics.ptr, phases, self.n_sub_steps, force_reset,
                                                   self.physics._stream()))

    def _sample_collision_free(self, ids: torch.Tensor) -> None:
        """DexterousHand.sample_collision_free_joint_angles (dexterous_hand.py:144-168) for envs `ids`."""
        sc = self.goal_generator._scratch_batch(ids.numel(), self.dtype)
        pending = torch.arange(ids.numel(), device=self.device)
        for _ in range(200):
            m = pending.numel()
            if m == 0:
                return
            q = self.hand.sample_joint_angles(m, self._gen, self._INIT_JOINT_RANGE_FRACTION, self.device, self.dtype)
            sc.qpos[:m] = self.physics.qpos[ids[pending]]
            sc.qpos[:m, :self.hand.num_joints] = q
            sc.collide()
            ok = sc.ncon[:m] == 0
            self.physics.qpos[ids[pending[ok]], :self.hand.num_joints] = q[ok]
            pending = pending[~ok]
        raise RuntimeError("could not sample a collision-free hand configuration")

    def _host_hooks(self) -> None:
        """Runs after the pre-step kernel: episode initialisation and goal changes (reach.py:149-194)."""
        fresh = self._skip.bool()
        changed = self._task_i[:, 4].bool()
        ids = torch.nonzero(changed, as_tuple=False).flatten()
        if ids.numel():
            while True:     # GoalEnvironment retry loop (environment.py:16-23)
                try:
                    g, gq = self.goal_generator.next_goal(self, ids, self._gen)
                    break
                except Exception as e:  # noqa: BLE001
                    from dexterity_b200 import goals
                    if not isinstance(e, goals.GoalInitializationError):
                        raise
            self._goal[ids, :g.shape[1]] = g
            self.goal_qpos[ids] = gq
        fid = torch.nonzero(fresh, as_tuple=False).flatten()
        if fid.numel():
            self._sample_collision_free(fid)

    def reset(self) -> TimeStep:
        self._phase(1, force_reset=1)
        self._host_hooks()
        self._phase(4)
        return self._timestep()

    def step_bound(self) -> TimeStep:
        self._phase(1)
        self._host_hooks()
        self._phase(2 | 4)
        return self._timestep()


def _mat_to_quat(R: torch.Tensor) -> torch.Tensor:
    """[..., 3, 3] rotation matrices -> [..., 4] unit quaternions (w, x, y, z), w >= 0."""
    m00, m11, m22 = R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]
    w = (1 + m00 + m11 + m22).clamp_min(0).sqrt() / 2
    x = (1 + m00 - m11 - m22).clamp_min(0).sqrt() / 2
    y = (1 - m00 + m11 - m22).clamp_min(0).sqrt() / 2
    z = (1 - m00 - m11 + m22).clamp_min(0).sqrt() / 2
    x = torch.copysign(x, R[..., 2, 1] - R[..., 1, 2])
    y = torch.copysign(y, R[..., 0, 2] - R[..., 2, 0])
    z = torch.copysign(z, R[..., 1, 0] - R[..., 0, 1])
    q = torch.stack([w, x, y, z], -1)
    return q / q.norm(dim=-1, keepdim=True)


def _fk_sites(compiled, hand, physics, frames: bool = False, extra_body: Optional[int] = None):
    """Torch forward kinematics of the fingertip sites for a scratch batch (hinge chains only): positions [n, 3 * sites], or
    with `frames` (positions [n, sites, 3], rotations [n, sites, 3, 3], position and rotation of body `extra_body`)."""
    c = compiled
    n = physics.num_envs
    dev, dt = physics.device, physics.dtype
    q = physics.qpos
    pos = [None] * c.nbody
    rot = [None] * c.nbody
    eye = torch.eye(3, dtype=dt, device=dev).expand(n, 3, 3)
    pos[0], rot[0] = torch.zeros(n, 3, dtype=dt, device=dev), eye
    omega = [None] * c.nbody                       # world-frame angular velocity of every needed body (frames only)
    omega[0] = torch.zeros(n, 3, dtype=dt, device=dev)
    qd = physics.qvel
    needed = set()
    for s in hand.fingertip_site_ids:
        b = int(c.SITE_BODY[s])
        while b > 0:
            needed.add(b); b = int(c.BODY_PARENT[b])
    b = extra_body if extra_body is not None else 0
    while b > 0:
        needed.add(b); b = int(c.BODY_PARENT[b])
    from dexterity_b200.model import mjcf
    for b in range(1, c.nbody):
        if b not in needed:
            continue
        p = int(c.BODY_PARENT[b])
        R0 = torch.as_tensor(mjcf.quat_to_mat(c.BODY_QUAT[b]), dtype=dt, device=dev)
        x = pos[p] + (rot[p] @ torch.as_tensor(c.BODY_POS[b], dtype=dt, device=dev))
        R = rot[p] @ R0
        w = omega[p]
        for j in range(int(c.BODY_JNTADR[b]), int(c.BODY_JNTADR[b] + c.BODY_JNTNUM[b])):
            if c.JNT_TYPE[j] != 3:
                raise NotImplementedError
            ax = torch.as_tensor(c.JNT_AXIS[j], dtype=dt, device=dev)
            jp = torch.as_tensor(c.JNT_POS[j], dtype=dt, device=dev)
            ang = q[:, int(c.JNT_QPOSADR[j])] - float(c.QPOS0[int(c.JNT_QPOSADR[j])])
            K = torch.tensor([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]], dtype=dt, device=dev)
            Rj = torch.eye(3, dtype=dt, device=dev) + ang.sin()[:, None, None] * K + (1 - ang.cos())[:, None, None] * (K @ K)
            anchor = x + R @ jp
            if frames:
                w = w + (R @ ax) * qd[:, int(c.JNT_DOFADR[j])].unsqueeze(1)      # hinge: axis (world) x joint velocity
            R = R @ Rj
            x = anchor - R @ jp
        pos[b], rot[b], omega[b] = x, R, w
    out, mats = [], []
    for s in hand.fingertip_site_ids:
        b = int(c.SITE_BODY[s])
        out.append(pos[b] + rot[b] @ torch.as_tensor(c.SITE_POS[s], dtype=dt, device=dev))
        if frames:
            mats.append(rot[b] @ torch.as_tensor(mjcf.quat_to_mat(np.asarray(c.SITE_QUAT[s])), dtype=dt, device=dev))
    if frames:
        angvel = torch.stack([omega[int(c.SITE_BODY[s])] for s in hand.fingertip_site_ids], 1)
        return torch.stack(out, 1), torch.stack(mats, 1), pos[extra_body], rot[extra_body], angvel
    return torch.cat(out, dim=1)


def load(domain_name: str, task_name: str = "state_dense", seed: Optional[int] = None,
         strip_singleton_obs_buffer_dim: bool = True, time_limit: Optional[float] = None,
         num_envs: int = 1, device: int = 0, dtype: torch.dtype = torch.float32,
         hand: Optional[str] = None) -> BatchedGoalEnvironment:
    """Batched counterpart of reference `manipulation.load` (same leading arguments).  `hand` defaults to the hand of the
    reference's registered task: the Shadow hand for reorient (reorient.py:324-364), the Adroit hand for reach
    (reach.py:223-249); `hand="shadow_hand_e"` gives the Shadow-hand reach of BASELINE.json configs[0..1]."""
    del strip_singleton_obs_buffer_dim
    tasks = {"reorient": ("state_dense",), "reach": ("state_dense", "state_sparse")}
    if domain_name not in tasks:
        raise ValueError(f"Unknown domain: {domain_name}")
    if task_name not in tasks[domain_name]:
        raise ValueError(f"Unknown task: {task_name}")
    if domain_name == "reach":
        return ReachEnvironment(num_envs, seed, device, dtype, time_limit, dense_reward=(task_name == "state_dense"),
                                hand=hand or "adroit_hand")
    return BatchedGoalEnvironment(domain_name, num_envs, seed, device, dtype, time_limit, hand=hand or "shadow_hand_e")
