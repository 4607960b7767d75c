"""numpy restatement of the reference's IK solver (TEST INFRASTRUCTURE ONLY).

Follows reference dexterity/inverse_kinematics/ik_solver.py:71-228 and
controllers/dls/dls.py:43-77 line by line, with MuJoCo's mj_kinematics / mj_jacSite replaced by
the numpy FK and point Jacobian of dexterity_b200.model.compile (pure kinematics: these are
definitions, not engine-specific numerics).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from dexterity_b200.model import compile as C

LINEAR_VELOCITY_GAIN = 0.95          # ik_solver.py:20-23
INTEGRATION_TIMESTEP_SEC = 1.0       # :25
REGULARIZATION_WEIGHT = 1e-5         # :27
PROGRESS_THRESHOLD = 20.0            # :31


def site_positions(arrays, qpos, site_ids):
    kin = C.kinematics(arrays, qpos)
    out = []
    for s in site_ids:
        b = arrays["SITE_BODY"][s]
        out.append(kin["xpos"][b] + kin["xmat"][b] @ arrays["SITE_POS"][s])
    return np.array(out), kin


def dls_joint_velocities(arrays, kin, site_ids, sites_xpos, twists):
    """DampedLeastSquaresMapper.compute_joint_velocities (dls.py:43-77)."""
    J = np.concatenate([C.point_jacobian(arrays, kin, int(arrays["SITE_BODY"][s]), sites_xpos[k])[0]
                        for k, s in enumerate(site_ids)], axis=0)
    twist = np.concatenate(twists, axis=0)
    hess = J.T @ J + np.eye(J.shape[1]) * REGULARIZATION_WEIGHT
    return np.linalg.solve(hess, J.T @ twist)


def solve_ik(arrays, site_ids, target_positions, qpos0, linear_tol, max_steps, early_stop) -> Tuple[np.ndarray, list]:
    """IKSolver._solve_ik (ik_solver.py:155-228) from the initial configuration qpos0."""
    rng_lo, rng_hi = arrays["JNT_RANGE"][:, 0], arrays["JNT_RANGE"][:, 1]
    qpos = qpos0.copy()
    cur, kin = site_positions(arrays, qpos, site_ids)
    prev = cur.copy()
    errs = [np.inf] * len(site_ids)
    for _ in range(max_steps):
        twists = [LINEAR_VELOCITY_GAIN * (t - c) / INTEGRATION_TIMESTEP_SEC for c, t in zip(cur, target_positions)]
        qdot = dls_joint_velocities(arrays, kin, site_ids, cur, twists)
        qpos = np.clip(qpos + qdot * INTEGRATION_TIMESTEP_SEC, rng_lo, rng_hi)   # mj_integratePos + clip (:189-194,240-246)
        cur, kin = site_positions(arrays, qpos, site_ids)
        errs, close_enough, not_enough_progress = [], True, False
        for i, t in enumerate(target_positions):
            e = float(np.linalg.norm(t - cur[i]))
            errs.append(e)
            if e > linear_tol:
                close_enough = False
            change = np.linalg.norm(cur[i] - prev[i])
            if e / (change + 1e-10) > PROGRESS_THRESHOLD:
                not_enough_progress = True
            prev[i] = cur[i]
        if (early_stop and close_enough) or not_enough_progress:
            break
    return qpos, errs


def solve(arrays, site_ids: Sequence[int], target_positions: np.ndarray, linear_tol: float = 1e-3,
          max_steps: int = 100, early_stop: bool = False, num_attempts: int = 30,
          stop_on_first_successful_attempt: bool = False,
          random_state: Optional[np.random.RandomState] = None) -> Optional[np.ndarray]:
    """IKSolver.solve (ik_solver.py:71-153)."""
    target_positions = np.asarray(target_positions, dtype=np.float64).reshape(-1, 3)
    if target_positions.shape[0] != len(site_ids):
        raise ValueError("The number of target positions must be equal to the number of end-effector sites.")
    random_state = random_state or np.random
    rng = arrays["JNT_RANGE"]
    nullspace_reference = rng.mean(axis=1)
    best, sol, success = np.inf, None, False
    for attempt in range(num_attempts):
        q0 = nullspace_reference.copy() if attempt == 0 else random_state.uniform(rng[:, 0], rng[:, 1])
        qpos, errs = solve_ik(arrays, site_ids, target_positions, q0, linear_tol, max_steps, early_stop)
        if all(e <= linear_tol for e in errs):
            success = True
            d = float(np.linalg.norm(qpos - nullspace_reference))
            if d < best:
                best, sol = d, qpos
        if success and stop_on_first_successful_attempt:
            break
    return sol
