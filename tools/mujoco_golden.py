"""Writes tests/golden/reorient_mujoco_steps.npz: TRUE MuJoCo results for the committed golden states.

RUN THIS ON A MACHINE THAT HAS THE REFERENCE'S DEPENDENCIES (`pip install mujoco dm_control
dm_robotics-transformations dm_robotics-geometry` and the reference checkout on PYTHONPATH); the build
container has none of them and no network, so this script has never been executed there.  It is the
missing half of the parity story: the file it writes has the same keys
as tests/golden/reorient_oracle_steps.npz, and once it is committed
`tests/test_oracle_cpu.py::test_true_mujoco_golden` compares the FP64 oracle (and through it the CUDA
path) with MuJoCo itself instead of skipping with "parity unpinned".

    python tools/mujoco_golden.py [--reference /path/to/dexterity] [--out tests/golden/reorient_mujoco_steps.npz]
    python tools/mujoco_golden.py --mjcf /tmp/reorient/reorient_shadow.xml --out tests/golden/reorient_mujoco_steps_own_mjcf.npz

The second form loads the flat MJCF written by tools/export_task.py (THIS build's scene assembly) into stock MuJoCo
instead of building the reference's environment: a difference between the two outputs is a difference in scene
assembly (PyMJCF attach, defaults), not in the engine restatement.

What it does: builds the reference's own reorient environment (dexterity.manipulation.load), maps
joints / actuators / geoms to this repo's compiled model BY NAME, and for every state of the oracle
fixture sets (qpos, qvel, ctrl, qacc_warmstart), runs mj_forward to read the contact list of that state,
then one mj_step, and records next qpos / qvel / qacc_warmstart, qacc and the contacts.  It also dumps the
compiled-model facts SURVEY.md §8c lists as unknown (mujoco version, counts, recentred mesh geom frames,
rbound, invweight0, solver options) under `model_*` keys so that the compile step can be checked too.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("DEXTERITY_REFERENCE", "/root/reference"))
    ap.add_argument("--states", default=os.path.join(ROOT, "tests", "golden", "reorient_oracle_steps.npz"))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "reorient_mujoco_steps.npz"))
    ap.add_argument("--mjcf", default=None, help="flat MJCF from tools/export_task.py instead of the reference's environment")
    args = ap.parse_args()
    sys.path.insert(0, args.reference)

    import mujoco                                   # noqa: E402  (absent in the build container)
    from dexterity_b200.model import scenes         # noqa: E402

    c = scenes.load_compiled("reorient_shadow")
    env = None
    if args.mjcf:
        m = mujoco.MjModel.from_xml_path(args.mjcf)
        d = mujoco.MjData(m)
        mujoco.mj_forward(m, d)
        # the hands' constant gravity compensation (reference utils/mujoco_utils.py:91-99) is not expressible in MJCF
        for b in range(m.nbody):
            if (mujoco.mj_id2name(m, mujoco.mjtObj.mjOBJ_BODY, b) or "").startswith("shadow_hand_e/"):
                d.xfrc_applied[b, :3] = -m.opt.gravity * m.body_mass[b]
        xfrc = d.xfrc_applied.copy()
    else:
        from dexterity import manipulation          # noqa: E402  (the reference itself)
        env = manipulation.load(domain_name="reorient", task_name="state_dense", seed=0)
        env.reset()
        physics = env.physics
        m, d = physics.model.ptr, physics.data.ptr
        xfrc = d.xfrc_applied.copy()

    def name(objtype, i):
        return mujoco.mj_id2name(m, objtype, i) or ""

    # ---- name maps (ours -> MuJoCo's addresses)
    jid = {name(mujoco.mjtObj.mjOBJ_JOINT, j): j for j in range(m.njnt)}
    aid = {name(mujoco.mjtObj.mjOBJ_ACTUATOR, a): a for a in range(m.nu)}
    gname = [name(mujoco.mjtObj.mjOBJ_GEOM, g) for g in range(m.ngeom)]
    ours_of_geom = {n: i for i, n in enumerate(c.geom_names)}
    qmap, vmap = [], []
    for jn in c.joint_names:
        key = jn if jn in jid else jn.rstrip("/") + "/"     # the prop's free joint is named after its body
        if key not in jid:
            raise SystemExit(f"joint {jn!r} of the compiled model not found in the MuJoCo model: {sorted(jid)[:8]} ...")
        j = jid[key]
        nq = {mujoco.mjtJoint.mjJNT_FREE: 7, mujoco.mjtJoint.mjJNT_BALL: 4}.get(m.jnt_type[j], 1)
        nv = {mujoco.mjtJoint.mjJNT_FREE: 6, mujoco.mjtJoint.mjJNT_BALL: 3}.get(m.jnt_type[j], 1)
        qmap += list(range(m.jnt_qposadr[j], m.jnt_qposadr[j] + nq))
        vmap += list(range(m.jnt_dofadr[j], m.jnt_dofadr[j] + nv))
    amap = [aid[a] for a in c.actuator_names]
    qmap, vmap, amap = np.array(qmap), np.array(vmap), np.array(amap)
    assert len(qmap) == c.nq and len(vmap) == c.nv and len(amap) == c.nu
    if m.nq != c.nq or m.nv != c.nv:
        print(f"note: MuJoCo model has nq={m.nq} nv={m.nv}, ours {c.nq}/{c.nv}; unmapped coordinates stay at their reset values")

    z = np.load(args.states)
    n = len(z["qpos"])
    out = {k: z[k] for k in ("qpos", "qvel", "ctrl", "warm")}
    nxt_q, nxt_v, nxt_w, qacc = np.zeros((n, c.nq)), np.zeros((n, c.nv)), np.zeros((n, c.nv)), np.zeros((n, c.nv))
    ncon = np.zeros(n, np.int32)
    niter = np.zeros(n, np.int32)
    cmax = 64
    con_geom = -np.ones((n, cmax, 2), np.int32)      # OUR geom ids (by name), -1 = geom unknown to the compiled model
    con_dist = np.zeros((n, cmax))
    con_pos = np.zeros((n, cmax, 3))
    con_normal = np.zeros((n, cmax, 3))
    qpos_reset, qvel_reset = d.qpos.copy(), d.qvel.copy()
    for i in range(n):
        mujoco.mj_resetData(m, d)
        d.xfrc_applied[:] = xfrc
        d.qpos[:] = qpos_reset; d.qvel[:] = qvel_reset
        d.qpos[qmap] = z["qpos"][i]; d.qvel[vmap] = z["qvel"][i]
        d.ctrl[amap] = z["ctrl"][i]; d.qacc_warmstart[vmap] = z["warm"][i]
        mujoco.mj_forward(m, d)
        d.qacc_warmstart[vmap] = z["warm"][i]         # mj_forward overwrote it with its own solution
        k = min(d.ncon, cmax)
        ncon[i] = d.ncon
        for j in range(k):
            ct = d.contact[j]
            con_geom[i, j] = [ours_of_geom.get(gname[ct.geom1], -1), ours_of_geom.get(gname[ct.geom2], -1)]
            con_dist[i, j] = ct.dist; con_pos[i, j] = ct.pos; con_normal[i, j] = ct.frame[:3]
        mujoco.mj_step(m, d)
        nxt_q[i], nxt_v[i], nxt_w[i] = d.qpos[qmap], d.qvel[vmap], d.qacc_warmstart[vmap]
        qacc[i] = d.qacc[vmap]
        niter[i] = int(np.max(d.solver_niter)) if hasattr(d, "solver_niter") else d.solver_iter
    out.update(next_qpos=nxt_q, next_qvel=nxt_v, next_warm=nxt_w, qacc=qacc, ncon=ncon, niter=niter,
               con_geom=con_geom, con_dist=con_dist, con_pos=con_pos, con_normal=con_normal)
    # ---- compiled-model facts (SURVEY.md §8c "unknowns")
    g_ours = np.array([ours_of_geom.get(nm, -1) for nm in gname])
    out.update(model_mujoco_version=np.array(mujoco.__version__), model_counts=np.array([m.nbody, m.njnt, m.nq, m.nv, m.nu, m.ngeom, m.nsensordata]),
               model_geom_names=np.array(gname), model_geom_ours=g_ours, model_geom_pos=m.geom_pos.copy(), model_geom_quat=m.geom_quat.copy(),
               model_geom_rbound=m.geom_rbound.copy(), model_geom_size=m.geom_size.copy(), model_body_invweight0=m.body_invweight0.copy(),
               model_dof_invweight0=m.dof_invweight0[vmap].copy(), model_body_mass=m.body_mass.copy(), model_body_inertia=m.body_inertia.copy(),
               model_opt=np.array([m.opt.timestep, m.opt.tolerance, m.opt.iterations, m.opt.solver, m.opt.cone, m.opt.integrator, m.opt.impratio]),
               model_stat_meaninertia=np.array(m.stat.meaninertia), model_observation_keys=np.array(list(env.observation_spec().keys()) if env is not None else []))
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, "mujoco", mujoco.__version__, "ncon histogram", np.bincount(ncon))


if __name__ == "__main__":
    main()
