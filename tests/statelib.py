"""Seeded state generation shared by the parity tests, smoke() and bench.py's cpu_baseline.

States are produced by rolling the FP64 oracle from the task's reset distribution with uniform
random controls (the recipe of reference manipulation_test.py:44-45), so they contain the
contact configurations the task actually visits.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from oracle import oracle as orc


def reorient_reset(compiled, rng: np.random.RandomState) -> np.ndarray:
    """qpos at episode start (reference reorient.py:72-78,143-151): hand at qpos0, cube placed
    uniformly in the 5 cm box above the palm with a uniform random orientation."""
    q = compiled.QPOS0.copy()
    nq = compiled.nq
    if nq >= 31:
        q[nq - 7:nq - 4] = [rng.uniform(-0.025, 0.025), rng.uniform(-0.155, -0.105), 0.16]
        quat = rng.randn(4)
        q[nq - 4:] = quat / np.linalg.norm(quat)
    else:
        # reach (reference reach.py:155-158): hand joints uniform in half of their range
        q[:24] = rng.uniform(0.5 * compiled.JNT_RANGE[:24, 0], 0.5 * compiled.JNT_RANGE[:24, 1])
    return q


def rollout_states(compiled, seed: int, nstates: int, every: int = 10, nsub: int = 1,
                   episodes: int = 1) -> Dict[str, np.ndarray]:
    """Rolls the oracle and snapshots (qpos, qvel, ctrl, qacc_warmstart) every `every` steps."""
    rng = np.random.RandomState(seed)
    om = orc.OracleModel(compiled)
    d = orc.OracleData(om)
    lo, hi = compiled.ACT_CTRLRANGE[:, 0], compiled.ACT_CTRLRANGE[:, 1]
    out = dict(qpos=[], qvel=[], ctrl=[], warm=[])
    per_ep = (nstates + episodes - 1) // episodes
    for _ in range(episodes):
        d.set_state(reorient_reset(compiled, rng), np.zeros(compiled.nv), np.zeros(compiled.nu), np.zeros(compiled.nv))
        ctrl = rng.uniform(lo, hi)
        for i in range(per_ep * every):
            if i % 5 == 0:
                ctrl = rng.uniform(lo, hi)
                q, v, w, _ = d.get_state()
                d.set_state(q, v, ctrl, w)
            d.step()
            if i % every == every - 1 and len(out["qpos"]) < nstates:
                q, v, w, _ = d.get_state()
                out["qpos"].append(q); out["qvel"].append(v); out["ctrl"].append(ctrl.copy()); out["warm"].append(w)
    return {k: np.array(v) for k, v in out.items()}


def rollout_states_batch(compiled, seed: int, nenv: int = 256, snapshots: int = 16, every: int = 5,
                         nthreads: int = 8) -> Dict[str, np.ndarray]:
    """nenv x snapshots states from `nenv` oracle rollouts run in parallel threads (orc_step_batch): every env takes
    `every` mj_steps with one uniform random control per snapshot; envs whose cube has fallen below the hand are
    reset from the task's reset distribution (reference reorient.py:206-209)."""
    rng = np.random.RandomState(seed)
    om = orc.OracleModel(compiled)
    lo, hi = compiled.ACT_CTRLRANGE[:, 0], compiled.ACT_CTRLRANGE[:, 1]
    qpos = np.stack([reorient_reset(compiled, rng) for _ in range(nenv)])
    qvel = np.zeros((nenv, compiled.nv)); warm = np.zeros((nenv, compiled.nv))
    out = dict(qpos=[], qvel=[], ctrl=[], warm=[])
    for _ in range(snapshots):
        ctrl = rng.uniform(lo, hi, size=(nenv, compiled.nu))
        om.step_batch(qpos, qvel, ctrl, warm, nsub=every, nthreads=nthreads)
        bad = ~np.isfinite(qpos).all(axis=1) | ~np.isfinite(qvel).all(axis=1)
        if compiled.nq >= 31:
            bad |= qpos[:, compiled.nq - 5] < 0.03          # cube centre near the ground: the episode would have ended
        for e in np.nonzero(bad)[0]:
            qpos[e] = reorient_reset(compiled, rng); qvel[e] = 0; warm[e] = 0
        out["qpos"].append(qpos.copy()); out["qvel"].append(qvel.copy()); out["ctrl"].append(ctrl.copy()); out["warm"].append(warm.copy())
    return {k: np.concatenate(v) for k, v in out.items()}


def oracle_step(compiled, states: Dict[str, np.ndarray], nsub: int = 1, forward_only: bool = False):
    """Oracle result for every state: next state, contacts (as sets), qacc."""
    om = orc.OracleModel(compiled)
    d = orc.OracleData(om)
    n = len(states["qpos"])
    res = dict(qpos=[], qvel=[], warm=[], ncon=[], pairs=[], dist=[], qacc=[], niter=[], normal=[], pos=[])
    for e in range(n):
        d.set_state(states["qpos"][e], states["qvel"][e], states["ctrl"][e], states["warm"][e])
        d.forward()
        c = d.contacts()
        res["ncon"].append(len(c["dist"]))
        res["pairs"].append([tuple(g) for g in c["geom"]])
        res["dist"].append(c["dist"]); res["normal"].append(c["frame"][:, 0] if len(c["dist"]) else np.zeros((0, 3)))
        res["pos"].append(c["pos"])
        res["qacc"].append(d.get("qacc")); res["niter"].append(d.solver_niter)
        if not forward_only:
            d.set_state(states["qpos"][e], states["qvel"][e], states["ctrl"][e], states["warm"][e])
            for _ in range(nsub):
                d.step()
        q, v, w, _ = d.get_state()
        res["qpos"].append(q); res["qvel"].append(v); res["warm"].append(w)
    for k in ("qpos", "qvel", "warm", "qacc"):
        res[k] = np.array(res[k])
    res["ncon"] = np.array(res["ncon"])
    return res


def upload(batch, states: Dict[str, np.ndarray]) -> None:
    import torch
    t = lambda a: torch.as_tensor(a, dtype=batch.dtype, device=batch.device)
    batch.qpos.copy_(t(states["qpos"])); batch.qvel.copy_(t(states["qvel"]))
    batch.ctrl.copy_(t(states["ctrl"])); batch.qacc_warmstart.copy_(t(states["warm"]))
    batch.time.zero_(); batch.status.zero_()


def mpr_stable_mask(compiled, states: Dict[str, np.ndarray], eps: float = 1e-7, trials: int = 3, seed: int = 0) -> np.ndarray:
    """Non-degeneracy probe.  libccd's MPR terminates on whichever portal first stops expanding, so
    for deep or face-face penetrations its (normal, depth) depend on sign decisions that flip under
    perturbations far below FP32 resolution — in the reference engine itself.  A state is called
    non-degenerate when the ORACLE's own contact set and normals survive `trials` random
    perturbations of qpos of size `eps` (FP32 rounding level).  Parity is asserted on those states;
    the degenerate share is reported alongside."""
    rng = np.random.RandomState(seed)
    om = orc.OracleModel(compiled)
    d = orc.OracleData(om)
    n = len(states["qpos"])
    stable = np.ones(n, bool)

    def contacts(q, e):
        d.set_state(q, states["qvel"][e], states["ctrl"][e], states["warm"][e])
        d.forward()
        c = d.contacts()
        return c["geom"].copy(), (c["frame"][:, 0].copy() if len(c["dist"]) else np.zeros((0, 3)))

    for e in range(n):
        g0, n0 = contacts(states["qpos"][e], e)
        if len(n0) == 0:
            continue
        for _ in range(trials):
            g1, n1 = contacts(states["qpos"][e] + eps * rng.randn(compiled.nq), e)
            if g0.shape != g1.shape or (g0 != g1).any() or np.abs(n0 - n1).max() > 1e-3:
                stable[e] = False
                break
    return stable


def many_contact_states(compiled, seed: int, nstates: int, lo_con: int = 9, hi_con: int = 28) -> Dict[str, np.ndarray]:
    """Static states with lo_con..hi_con contacts (a half-closed hand around a cube pushed into the palm): far
    more contacts than the random-action rollouts visit (one on average)."""
    rng = np.random.RandomState(seed)
    om = orc.OracleModel(compiled)
    d = orc.OracleData(om)
    lo, hi = compiled.JNT_RANGE[:24, 0], compiled.JNT_RANGE[:24, 1]
    out = dict(qpos=[], qvel=[], ctrl=[], warm=[])
    while len(out["qpos"]) < nstates:
        q = reorient_reset(compiled, rng)
        f = rng.uniform(0.25, 0.6)
        q[2:24] = lo[2:] + (hi[2:] - lo[2:]) * np.clip(f + 0.1 * rng.randn(22), 0, 1)
        q[24:27] = [rng.uniform(-0.02, 0.02), rng.uniform(-0.14, -0.10), rng.uniform(0.10, 0.14)]
        v = 0.1 * rng.randn(compiled.nv)
        d.set_state(q, v, np.zeros(compiled.nu), np.zeros(compiled.nv))
        d.forward()
        ct = d.contacts()
        if lo_con <= len(ct["dist"]) <= hi_con and (np.abs(ct["dist"]) > 1e-5).all():
            out["qpos"].append(q); out["qvel"].append(v)
            out["ctrl"].append(rng.uniform(compiled.ACT_CTRLRANGE[:, 0], compiled.ACT_CTRLRANGE[:, 1])); out["warm"].append(np.zeros(compiled.nv))
    return {k: np.array(v) for k, v in out.items()}
