"""CPU-only execution of the CUDA kernel BODIES under the test-only warp emulator
(tests/emu/warp_emu.cpp) against the oracle.  This guards kernel logic between GPU runs; the GPU
parity proper is tests/test_physics_gpu.py.  The emulator is test infrastructure, not a fallback."""
import ctypes

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import task_oracle as T
from tests import statelib
from tests.emu import emu as E


@pytest.fixture(scope="module")
def emu_model(reorient_model):
    return E.EmuModel(reorient_model)


@pytest.fixture(scope="module")
def states(reorient_model):
    return statelib.rollout_states(reorient_model, seed=77, nstates=12, every=11, episodes=3)


def test_kernel_bodies_fp64_match_oracle(reorient_model, emu_model, states):
    ref = statelib.oracle_step(reorient_model, states, nsub=1)
    out = emu_model.run(states["qpos"], states["qvel"], states["ctrl"], states["warm"], nsub=1, dtype=np.float64, debug=True)
    assert (out["ncon"] == ref["ncon"]).all() and (out["status"] == 0).all()
    assert np.abs(out["qacc"] - ref["qacc"]).max() < 1e-6 * max(1.0, np.abs(ref["qacc"]).max())
    assert np.abs(out["qpos"] - ref["qpos"]).max() < 1e-10 and np.abs(out["qvel"] - ref["qvel"]).max() < 1e-7


def test_tree_kernel_forms_agree(reorient_model, emu_model, states, monkeypatch):
    """K0 exists as a thread-per-env walk and as an 8-lanes-per-env schedule; both must give the same step."""
    assert emu_model.lib.emu_k0_rounds(ctypes.c_void_p(emu_model.ptr)) > 0   # the Shadow hand fits the 8-lane schedule
    a = emu_model.run(states["qpos"], states["qvel"], states["ctrl"], states["warm"], nsub=2, dtype=np.float64, debug=True)
    monkeypatch.setenv("DEX_EMU_TREE_SCALAR", "1")
    b = emu_model.run(states["qpos"], states["qvel"], states["ctrl"], states["warm"], nsub=2, dtype=np.float64, debug=True)
    for k in ("qpos", "qvel", "M", "fs"):
        assert np.abs(a[k] - b[k]).max() < 1e-10, k
    assert (a["ncon"] == b["ncon"]).all()


def test_hill_climbing_support_is_exact(emu_model):
    """The narrowphase climbs the hull graph instead of scanning every vertex (as MuJoCo's mesh support does).
    On a convex hull a local maximum is the global one: over all 204 hulls and cold / warm starts the support VALUE
    must equal the exhaustive maximum; only the vertex may differ, and only on exact ties."""
    out = (ctypes.c_double * 4)()
    emu_model.lib.emu_support_check.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.POINTER(ctypes.c_double)]
    emu_model.lib.emu_support_check(ctypes.c_void_p(emu_model.ptr), 400, 12345, out)
    n, other_vertex, smaller, shortfall = out[0], out[1], out[2], out[3]
    assert n > 50000
    assert smaller == 0, f"{smaller} of {n} supports below the exhaustive maximum (largest shortfall {shortfall:.3e})"
    assert other_vertex == 0


def test_kernel_bodies_fp32_match_oracle(reorient_model, emu_model, states):
    ref = statelib.oracle_step(reorient_model, states, nsub=1)
    out = emu_model.run(states["qpos"], states["qvel"], states["ctrl"], states["warm"], nsub=1, dtype=np.float32)
    assert (out["ncon"] == ref["ncon"]).all()
    rel = np.abs(out["qvel"] - ref["qvel"]).max(axis=1) / np.maximum(1, np.abs(ref["qvel"]).max(axis=1))
    assert np.median(rel) < 1e-5 and (rel < 1e-4).mean() >= 0.8


def _finger_of_body(c):
    """branch id of every body of the Shadow hand (0-4 = ff mf rf lf th), -1 for the rest"""
    out = []
    for name in c.body_names:
        tag = name.split("/")[-1][:2]
        out.append({"ff": 0, "mf": 1, "rf": 2, "lf": 3, "th": 4}.get(tag, -1) if name.startswith("shadow_hand_e/") else -1)
    return np.array(out)


def test_structured_solver_forms_match_dense_kernel_and_oracle(reorient_model, emu_model, monkeypatch):
    """K2 has three kernel bodies: the dense warp-per-env solver, and the structured eight-lanes-per-env solver
    (dex_arrow.cuh) in its full and its contact-free form.  On contact-rich states, including states where a contact
    joins two fingers (the low-rank correction), all minimise the same cost: in FP64 they agree to rounding with each
    other and with the oracle, Newton iteration for Newton iteration."""
    c = reorient_model
    st = statelib.rollout_states_batch(c, seed=21, nenv=48, snapshots=6, every=7, nthreads=8)
    ref = statelib.oracle_step(c, st, nsub=1)
    fb = _finger_of_body(c)
    cross = np.array([any(fb[c.GEOM_BODY[g1]] >= 0 and fb[c.GEOM_BODY[g2]] >= 0 and fb[c.GEOM_BODY[g1]] != fb[c.GEOM_BODY[g2]]
                          for g1, g2 in pairs) for pairs in ref["pairs"]])
    assert cross.sum() >= 10 and (ref["ncon"] == 0).sum() >= 20 and ((ref["ncon"] > 0) & ~cross).sum() >= 20
    cnt = (ctypes.c_long * 2)()
    emu_model.lib.emu_solver_counts.argtypes = [ctypes.POINTER(ctypes.c_long), ctypes.c_int]
    outs = {}
    for mode in (0, 1, 2, 3):
        monkeypatch.setenv("DEX_EMU_SOLVER", str(mode))
        emu_model.lib.emu_solver_counts(cnt, 1)
        outs[mode] = emu_model.run(st["qpos"], st["qvel"], st["ctrl"], st["warm"], nsub=1, dtype=np.float64)
        emu_model.lib.emu_solver_counts(cnt, 1)
        n = len(st["qpos"])
        if mode == 0: assert cnt[0] == 0
        if mode == 1: assert cnt[0] == (ref["ncon"] == 0).sum()
        if mode == 2: assert cnt[0] >= 0.9 * n and cnt[0] > (~cross).sum() - 8, "cross-contact envs must take the structured solver too"
        if mode == 3: assert cnt[0] == (~cross).sum(), "the sparse form takes exactly the envs without a finger-finger contact"
        assert (outs[mode]["status"] == 0).all() and (outs[mode]["ncon"] == ref["ncon"]).all()
    scale = np.maximum(1, np.abs(outs[0]["qvel"]).max(axis=1))
    for mode in (1, 2, 3):
        d = np.abs(outs[mode]["qvel"] - outs[0]["qvel"]).max(axis=1) / scale
        assert d.max() < 1e-10, (mode, d.max())
        assert (outs[mode]["niter"] == outs[0]["niter"]).mean() > 0.98
        assert np.abs(outs[mode]["warm"] - outs[0]["warm"]).max() < 1e-7 * max(1.0, np.abs(outs[0]["warm"]).max())
    # and against the oracle (MPR tie states excepted, as everywhere)
    rel = np.abs(outs[2]["qvel"] - ref["qvel"]).max(axis=1) / np.maximum(1, np.abs(ref["qvel"]).max(axis=1))
    assert np.median(rel) < 1e-12 and (rel < 1e-6).mean() > 0.97
    # FP32, the product precision: same statistics as the dense kernel
    monkeypatch.setenv("DEX_EMU_SOLVER", "2")
    f32 = emu_model.run(st["qpos"], st["qvel"], st["ctrl"], st["warm"], nsub=1, dtype=np.float32)
    rel32 = np.abs(f32["qvel"] - ref["qvel"]).max(axis=1) / np.maximum(1, np.abs(ref["qvel"]).max(axis=1))
    assert np.median(rel32) < 1e-6 and (rel32 <= 1e-4).mean() > 0.95 and (f32["status"] == 0).all()
    assert (rel32[cross] <= 1e-4).mean() > 0.85


def _params(c):
    from dexterity_b200 import envs
    p = envs.TaskParams()
    cfg = envs.REORIENT
    p.task, p.successes_needed, p.steps_before_changing_goal, p.fall_termination, p.has_max_time = 0, 1, 5, 1, 1
    p.prop_body = c.body_names.index("prop/"); j = c.joint_names.index("prop/")
    p.prop_qposadr, p.prop_dofadr = int(c.JNT_QPOSADR[j]), int(c.JNT_DOFADR[j])
    p.prop_geom, p.ground_geom = c.geom_names.index("prop/geom"), c.geom_names.index("ground")
    p.nhand, p.nsite = 24, 5
    for i, t in enumerate(("fftip", "mftip", "rftip", "lftip", "thtip")):
        p.site_id[i] = c.site_names.index(f"shadow_hand_e/{t}_site")
    p.success_threshold, p.max_time_per_goal, p.time_limit = 0.1, 7.5, float("inf")
    for k in range(3):
        p.spawn_lo[k], p.spawn_hi[k] = cfg["spawn_lo"][k], cfg["spawn_hi"][k]
    p.w_orientation, p.w_success_bonus, p.w_action_smoothing, p.orientation_eps, p.bonus_threshold = 1.0, 800.0, -0.1, 0.1, 0.1
    return p


def test_reset_request_is_consumed_and_time_is_double(reorient_model, emu_model):
    """the RESET_MASK byte is cleared by the pre-step that serves it (the env is MID on the next call), and
    data.time advances in double (time += timestep per mj_step, as MuJoCo does) although the batch is FP32."""
    env = E.EmuEnv(emu_model, _params(reorient_model), nenv=2, dtype=np.float32, seed=4)
    env.reset()
    a = env.a
    env.step(np.zeros((2, 20)), nsub=2)
    assert (a["TASK_I"][:, 6] == 1).all()
    a["RESET_MASK"][0] = 1
    env.step(np.zeros((2, 20)), nsub=2)
    assert a["RESET_MASK"][0] == 0 and a["TASK_I"][0, 6] == 0 and a["TASK_I"][1, 6] == 1
    assert a["TIME64"][0, 0] == 0.0 and a["TIME"][0] == 0.0
    env.step(np.zeros((2, 20)), nsub=2)
    assert a["TASK_I"][0, 6] == 1, "an env reset through the mask resumes MID steps without manual clearing"
    h = reorient_model.opt("TIMESTEP")
    t2 = (0.0 + h) + h
    t6 = t2
    for _ in range(4):
        t6 += h
    assert a["TIME64"][0, 0] == t2 and a["TIME64"][1, 0] == t6
    assert a["TIME"][1] == np.float32(t6)


def test_task_kernels_against_restatement(reorient_model, emu_model):
    c = reorient_model
    env = E.EmuEnv(emu_model, _params(c), nenv=6, dtype=np.float64, seed=9)
    env.reset()
    a = env.a
    assert (a["TASK_I"][:, 6] == 0).all() and (a["REWARD"] == 0).all() and (a["DISCOUNT"] == 1).all()
    q = a["QPOS"][:, :31]
    assert np.allclose(np.linalg.norm(q[:, 27:31], axis=1), 1) and np.allclose(np.linalg.norm(a["GOAL"][:, :4], axis=1), 1)
    assert (np.abs(q[:, 24]) <= 0.025).all() and (np.abs(q[:, 25] + 0.13) <= 0.025).all() and np.allclose(q[:, 26], 0.16)
    assert len({tuple(np.round(x, 6)) for x in q[:, 24:31]}) == 6, "per-env RNG streams must differ"
    rng = np.random.RandomState(0)
    lo, hi = c.ACT_CTRLRANGE[:, 0], c.ACT_CTRLRANGE[:, 1]
    om = orc.OracleModel(c); d = orc.OracleData(om)
    sites = [c.site_names.index(f"shadow_hand_e/{t}_site") for t in ("fftip", "mftip", "rftip", "lftip", "thtip")]
    machines = [T.GoalTaskState(0.1, 1, 5, 7.5) for _ in range(6)]
    for m_ in machines:
        m_.initialize_episode(0.0)
    for it in range(3):
        act = rng.uniform(lo, hi, size=(6, 20))
        env.step(act, nsub=5)
        for e in range(6):
            assert a["TASK_I"][e, 6] in (1, 2)
            d.set_state(a["QPOS"][e, :31], a["QVEL"][e, :30], a["CTRL"][e, :20], np.zeros(30)); d.step1()
            ho = T.hand_observations(c, d, 24, sites)
            np.testing.assert_allclose(a["OBS"][e, :48], ho["joint_positions_sin_cos"], atol=1e-12)
            np.testing.assert_allclose(a["OBS"][e, 48:72], ho["joint_velocities"], atol=1e-12)
            np.testing.assert_allclose(a["OBS"][e, 72:87], ho["fingertip_positions"], atol=1e-12)
            np.testing.assert_allclose(a["OBS"][e, 87:102], ho["fingertip_linear_velocities"], atol=1e-10)
            po = T.prop_observations(c, d, 32, 24)
            np.testing.assert_allclose(a["OBS"][e, 102:105], po["position"], atol=1e-12)
            np.testing.assert_allclose(a["OBS"][e, 105:109], po["orientation"], atol=1e-12)
            np.testing.assert_allclose(a["OBS"][e, 109:112], po["linear_velocity"], atol=1e-10)
            np.testing.assert_allclose(a["OBS"][e, 112:115], po["angular_velocity"], atol=1e-10)
            goal = a["GOAL"][e, :4]
            dist = T.prop_orientation_goal_distance(goal, po["orientation"])[0]
            np.testing.assert_allclose(a["TASK_F"][e, 2], dist, atol=1e-7)
            np.testing.assert_allclose(a["REWARD"][e], T.weighted_average(T.reorient_reward_terms(dist, a["CTRL"][e, :20])), rtol=1e-9)
            np.testing.assert_allclose(a["CTRL"][e, :20], act[e], atol=0)      # stored unclipped
            m_ = machines[e]
            m_.before_step(a["TIME"][e] - 0.025)
            m_.after_step(np.array([dist]), a["TIME"][e])
            assert a["TASK_I"][e, 0] == m_.successes and a["TASK_I"][e, 1] == m_.success_change_counter
            fallen = bool((d.contacts()["geom"] == [c.geom_names.index("ground"), c.geom_names.index("prop/geom")]).all(axis=1).any()) \
                if len(d.contacts()["dist"]) else False
            assert (a["TASK_I"][e, 6] == 2) == (m_.should_terminate_episode() or fallen)


def test_prop_placement_rejects_colliding_poses(reorient_model, emu_model):
    """dm_control PropPlacer semantics (reference reorient.py:143-151) on the device code path, run by the emulator:
    with a spawn box that reaches into the hand, every accepted pose is contact-free according to the ORACLE, poses are
    redrawn inside the box, and envs whose draws all collide are flagged and end their episode at the next step."""
    c = reorient_model
    p = _params(c)
    p.spawn_lo[2], p.spawn_hi[2] = 0.09, 0.16           # reaches down into the palm / fingers
    p.place_max_attempts = 20
    n = 24
    env = E.EmuEnv(emu_model, p, nenv=n, dtype=np.float64, seed=3)
    env.reset()
    a = env.a
    assert (a["SKIP"] == 1).all()                        # every env left the placement loop
    om = orc.OracleModel(c); d = orc.OracleData(om)
    prop = c.geom_names.index("prop/geom")
    flagged = (a["STATUS"] & 32) != 0
    redrawn = a["RNG"][:, 2] > 3                         # 3 counters are used by the reset itself, 2 more per redraw
    assert redrawn.any() and not flagged.all()
    for e in range(n):
        q = a["QPOS"][e, :31]
        assert (q[24:27] >= [p.spawn_lo[0], p.spawn_lo[1], p.spawn_lo[2]]).all() and (q[24:27] <= [p.spawn_hi[0], p.spawn_hi[1], p.spawn_hi[2]]).all()
        assert abs(np.linalg.norm(q[27:31]) - 1) < 1e-12
        d.set_state(q, np.zeros(30), np.zeros(20), np.zeros(30)); d.forward()
        ct = d.contacts()
        touching = bool(len(ct["dist"]) and ((ct["geom"] == prop).any(axis=1) & (ct["dist"] <= 0)).any())
        assert touching == bool(flagged[e]), (e, touching, flagged[e])
    # the static proof (dexterity_b200/placement.py) agrees: this box is not provably clear, the reference's is
    from dexterity_b200 import placement
    pb = c.body_names.index("prop/")
    assert placement.place_max_attempts(c, pb, list(p.spawn_lo), list(p.spawn_hi)) == 20
    from dexterity_b200 import envs
    assert placement.place_max_attempts(c, pb, envs.REORIENT["spawn_lo"], envs.REORIENT["spawn_hi"]) == 0
    assert placement.spawn_clearance(c, pb, envs.REORIENT["spawn_lo"], envs.REORIENT["spawn_hi"]) > 0.01


def test_reference_spawn_box_never_collides(reorient_model):
    """The clearance proof's claim, sampled: no reset pose of the reference's spawn box is in contact (oracle)."""
    c = reorient_model
    om = orc.OracleModel(c); d = orc.OracleData(om)
    rng = np.random.RandomState(5)
    from tests import statelib
    for _ in range(300):
        d.set_state(statelib.reorient_reset(c, rng), np.zeros(30), np.zeros(20), np.zeros(30)); d.forward()
        assert len(d.contacts()["dist"]) == 0


def test_many_contact_states(reorient_model, emu_model):
    """9..28 contacts per env (lane = contact in the solver kernel, per-contact records up to DEX_MAXCON = 32):
    same result as the oracle in FP64.  States whose MPR portal is tie-dependent (cube face flat on a hull
    face: DESIGN.md, degenerate states) are excluded by comparing the contact normals first."""
    st = statelib.many_contact_states(reorient_model, seed=11, nstates=12)
    ref = statelib.oracle_step(reorient_model, st, nsub=1)
    assert ref["ncon"].min() > 8
    out = emu_model.run(st["qpos"], st["qvel"], st["ctrl"], st["warm"], nsub=1, dtype=np.float64, debug=True)
    assert (out["ncon"] == ref["ncon"]).all()
    om = orc.OracleModel(reorient_model); d = orc.OracleData(om)
    checked = 0
    for e in range(12):
        d.set_state(st["qpos"][e], st["qvel"][e], st["ctrl"][e], st["warm"][e]); d.forward()
        ct = d.contacts()
        kc = out["contact"][e][:out["ncon"][e]]
        if np.abs(kc[:, 4:7] - ct["frame"][:, 0]).max() > 1e-9 or np.abs(kc[:, 0] - ct["dist"]).max() > 1e-9:
            continue
        checked += 1
        assert np.abs(out["qacc"][e] - ref["qacc"][e]).max() < 1e-6 * max(1.0, np.abs(ref["qacc"][e]).max())
        assert np.abs(out["qpos"][e] - ref["qpos"][e]).max() < 1e-10 and np.abs(out["qvel"][e] - ref["qvel"][e]).max() < 1e-7
    assert checked >= 3
