"""numpy restatement of the reference's per-step task logic (TEST INFRASTRUCTURE ONLY).

Every function cites the reference file:line it follows.  These are the MuJoCo-independent
parts of the hot path, so — unlike the physics — they ARE pinned by the reference's own exact
known-answer tests (reorient_test.py:44-50, smooth_action_test.py:25-44, hands_test.py:26-31),
which tests/test_task_kat.py replays against this module.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np


# --- dm_robotics.transformations (used by goals/prop_orientation.py:40-50) -------------------
def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def quat_diff_active(source_quat, target_quat):
    """target = diff * source (dm_robotics tr.quat_diff_active)."""
    return quat_mul(target_quat, quat_conj(source_quat))


def quat_to_axisangle(quat):
    angle = 2 * np.arccos(np.clip(quat[0], -1, 1))
    if angle < 1e-10:
        return np.zeros(3)
    qn = np.sin(angle / 2)
    angle = (angle + np.pi) % (2 * np.pi) - np.pi
    return quat[1:4] / qn * angle


def prop_orientation_goal_distance(goal_state, current_state):
    """reference goals/prop_orientation.py:45-50."""
    err_quat = quat_diff_active(source_quat=current_state, target_quat=goal_state)
    return np.linalg.norm(quat_to_axisangle(err_quat), keepdims=True)


def uniform_quaternion(u1, u2, u3):
    """dm_control rotations.UniformQuaternion with u1 in [0,1), u2,u3 in [0,2pi)."""
    return np.array([np.sqrt(1 - u1) * np.sin(u2), np.sqrt(1 - u1) * np.cos(u2),
                     np.sqrt(u1) * np.sin(u3), np.sqrt(u1) * np.cos(u3)])


# --- rewards -----------------------------------------------------------------------------------
def tolerance_hard(x, lower, upper):
    """dm_control rewards.tolerance(x, bounds, margin=0): 1 inside the closed interval, else 0."""
    return 1.0 if lower <= x <= upper else 0.0


def reorient_reward_terms(distance: float, ctrl: np.ndarray):
    """reference manipulation/tasks/reorient.py:238-284 (values, weights)."""
    return dict(
        orientation=(1.0 / (distance + 0.1), 1.0),
        success_bonus=(tolerance_hard(distance, 0.0, 0.1), 800.0),
        action_smoothing=(float(np.linalg.norm(ctrl) ** 2), -0.1),
    )


def weighted_average(terms) -> float:
    """reference manipulation/shared/rewards.py:13-15 (a weighted SUM)."""
    return sum(v * w for v, w in terms.values())


def tanh_squared(x, margin: float, loss_at_margin: float = 0.95):
    """reference manipulation/shared/rewards.py:18-28."""
    error = np.linalg.norm(x)
    w = np.arctanh(np.sqrt(loss_at_margin)) / margin
    s = np.tanh(w * error)
    return s * s


def reach_reward(goal_distance: np.ndarray, dense: bool) -> float:
    """reference manipulation/tasks/reach.py:196-210."""
    if dense:
        return float(np.mean(np.where(goal_distance <= 0.01, 0.0,
                                      [-tanh_squared(d, margin=0.1) for d in goal_distance])))
    return float(np.mean(np.where(goal_distance <= 0.01, 0.0, -1.0)))


def smooth_action(prev: Optional[np.ndarray], x: np.ndarray, alpha: float) -> np.ndarray:
    """reference effectors/wrappers/smooth_action.py:20-24,54-59 (first call passes x through)."""
    return x.copy() if prev is None else alpha * prev + (1 - alpha) * x


# --- GoalTask state machine ----------------------------------------------------------------------
@dataclasses.dataclass
class GoalTaskState:
    """reference dexterity/task.py:137-204 for one environment."""
    success_threshold: float
    successes_needed: int = 1
    steps_before_changing_goal: int = 0
    max_time_per_goal: Optional[float] = None
    successes: int = 0
    success_change_counter: int = 0
    solve_start_time: float = 0.0
    exceeded_single_goal_time: bool = False
    success_registered: bool = False
    goal_changed: bool = True

    def initialize_episode(self, time: float) -> None:            # task.py:137-152
        self.successes = 0
        self.success_change_counter = 0
        self.solve_start_time = time
        self.exceeded_single_goal_time = False
        self.success_registered = False
        self.goal_changed = True

    def before_step(self, time: float) -> bool:                   # task.py:154-165
        if self.success_change_counter > self.steps_before_changing_goal:
            self.success_change_counter = 0
            self.exceeded_single_goal_time = False
            self.solve_start_time = time
            self.goal_changed = True
            self.success_registered = False
        else:
            self.goal_changed = False
        return self.goal_changed

    def after_step(self, goal_distance: np.ndarray, time: float) -> None:   # task.py:167-185
        if np.all(goal_distance <= self.success_threshold):
            self.success_change_counter += 1
            if not self.success_registered:
                self.successes += 1
                self.success_registered = True
        elif self.max_time_per_goal is not None:
            if time - self.solve_start_time > self.max_time_per_goal:
                self.exceeded_single_goal_time = True

    def should_terminate_episode(self) -> bool:                    # task.py:187-193
        return self.successes >= self.successes_needed or (
            self.max_time_per_goal is not None and self.exceeded_single_goal_time)

    def get_discount(self) -> float:                               # task.py:195-204
        return 0.0 if self.successes >= self.successes_needed else 1.0


# --- observables (models/hands/dexterous_hand.py:254-310, utils/mujoco_utils.py:10-35) -----------
def hand_observations(compiled, data, nhand: int, site_ids) -> dict:
    """`data` is an oracle.OracleData on which forward()/step1() has been run."""
    qpos, qvel, _, _ = data.get_state()
    sxp = data.get("site_xpos").reshape(-1, 3)
    lin = []
    for s in site_ids:
        v6 = data.object_velocity(int(compiled.SITE_BODY[s]), sxp[s])   # [ang, lin], world frame
        lin.append(v6[3:])
    return dict(
        joint_positions_sin_cos=np.vstack([np.sin(qpos[:nhand]), np.cos(qpos[:nhand])]).T.ravel(),
        joint_velocities=qvel[:nhand].copy(),
        fingertip_positions=sxp[list(site_ids)].ravel(),
        fingertip_linear_velocities=np.concatenate(lin),
    )


def prop_observations(compiled, data, body: int, dofadr: int) -> dict:
    """primitive.Primitive frame sensors of the prop geom (framepos/quat/linvel/angvel, world frame)."""
    qpos, qvel, _, _ = data.get_state()
    xpos = data.get("xpos").reshape(-1, 3)[body]
    xquat = data.get("xquat").reshape(-1, 4)[body]
    v6 = data.object_velocity(body, xpos)
    return dict(position=xpos.copy(), orientation=xquat.copy(), linear_velocity=v6[3:], angular_velocity=v6[:3])
