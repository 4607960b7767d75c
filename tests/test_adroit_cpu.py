"""The Adroit hand (reference models/hands/adroit_hand.py, vendor/adroit/.../adroit_hand.xml) through the same
compile -> oracle -> kernel-body chain as the Shadow hand: articulated dynamics with friction loss, joint limits with a
positive margin (adroit_hand.xml:21), general affine actuators (gainprm 10, biasprm 0 -10 0) and 24 actuated hinges.
The model's 19 explicit contact pairs (capsule-capsule, box-capsule, condim 1) are not compiled yet, so these are
contact-free dynamics; the fingertip golden vector of the reference is in tests/test_ik_cpu.py."""
import numpy as np
import pytest

from dexterity_b200.model import scenes
from tests import statelib
from tests.emu import emu as E


def adroit_states(c, seed, n):
    rng = np.random.RandomState(seed)
    lo, hi = c.JNT_RANGE[:, 0], c.JNT_RANGE[:, 1]
    return dict(qpos=lo + (hi - lo) * rng.uniform(-0.05, 1.05, (n, c.nq)),   # some joints beyond their limits
                qvel=rng.randn(n, c.nv) * 0.5,
                ctrl=rng.uniform(c.ACT_CTRLRANGE[:, 0], c.ACT_CTRLRANGE[:, 1], (n, c.nu)),
                warm=np.zeros((n, c.nv)))


@pytest.fixture(scope="module")
def adroit_model():
    return scenes.load_compiled("adroit_hand_only")


def test_adroit_limit_rows_are_exercised(adroit_model):
    from oracle import oracle as orc
    c = adroit_model
    st = adroit_states(c, 0, 16)
    d = orc.OracleData(orc.OracleModel(c))
    nlimit = 0
    for e in range(16):
        d.set_state(st["qpos"][e], st["qvel"][e], st["ctrl"][e], st["warm"][e]); d.forward()
        t, _ = d.efc_types()
        nlimit += int((t == 1).sum())
        assert (t == 0).sum() == 24                         # friction loss on every dof (adroit_hand.xml:21)
    assert nlimit >= 16


def test_adroit_kernel_bodies_match_oracle(adroit_model):
    c = adroit_model
    st = adroit_states(c, 0, 16)
    ref = statelib.oracle_step(c, st, nsub=1)
    m = E.EmuModel(c)
    out = m.run(st["qpos"], st["qvel"], st["ctrl"], st["warm"], nsub=1, dtype=np.float64, debug=True)
    assert (out["status"] == 0).all() and (out["ncon"] == 0).all()
    assert np.abs(out["qacc"] - ref["qacc"]).max() < 1e-9 * max(1.0, np.abs(ref["qacc"]).max())
    assert np.abs(out["qpos"][:, :c.nq] - ref["qpos"]).max() < 1e-12
    assert np.abs(out["qvel"][:, :c.nv] - ref["qvel"]).max() < 1e-10
    out32 = m.run(st["qpos"], st["qvel"], st["ctrl"], st["warm"], nsub=1, dtype=np.float32)
    rel = np.abs(out32["qvel"][:, :c.nv] - ref["qvel"]).max(axis=1) / np.maximum(1, np.abs(ref["qvel"]).max(axis=1))
    assert rel.max() <= 1e-4                                 # north-star single-step tolerance


# ----------------------------------------------------------------------------------------------------------------------
# The registered reach task's scene (reach_adroit): the hand's explicit <contact><pair condim="1"> elements
# (adroit_hand.xml:231-251): capsule-capsule in closed form (mjc_CapsuleCapsule), capsule-box through MPR, one frictionless
# constraint row per contact.
# ----------------------------------------------------------------------------------------------------------------------
def reach_adroit_states(c, seed, n, thumb_on_palm=False):
    rng = np.random.RandomState(seed)
    lo, hi = c.JNT_RANGE[:, 0], c.JNT_RANGE[:, 1]
    q = lo + (hi - lo) * rng.uniform(0.0, 1.0, (n, c.nq))
    if thumb_on_palm:   # thumb folded across the palm: the capsule-box pair C_thdistal - C_palm0
        for name, frac in (("THJ4", 0.95), ("THJ3", 0.9), ("THJ2", 0.5), ("THJ1", 0.2), ("THJ0", 0.1)):
            j = c.joint_names.index("adroit_hand/" + name)
            q[:, j] = lo[j] + (hi[j] - lo[j]) * np.clip(frac + 0.1 * rng.uniform(-1, 1, n), 0, 1)
    return dict(qpos=q, qvel=rng.randn(n, c.nv) * 0.5,
                ctrl=rng.uniform(c.ACT_CTRLRANGE[:, 0], c.ACT_CTRLRANGE[:, 1], (n, c.nu)), warm=np.zeros((n, c.nv)))


@pytest.fixture(scope="module")
def reach_adroit_model():
    return scenes.load_compiled("reach_adroit")


def test_reach_adroit_scene_facts(reach_adroit_model):
    c = reach_adroit_model
    assert (c.nq, c.nv, c.nu) == (24, 24, 24)                       # SURVEY.md section 8, reach scene
    assert c.npair == 18 and (np.asarray(c.PAIR_CONDIM) == 1).all()   # 19 <pair>s, one duplicated (adroit_hand.xml:247-248)
    types = sorted((int(c.GEOM_TYPE[a]), int(c.GEOM_TYPE[b])) for a, b in zip(c.PAIR_GEOM1, c.PAIR_GEOM2))
    assert types.count((3, 3)) == 17 and types.count((3, 6)) == 1
    np.testing.assert_allclose(np.asarray(c.PAIR_MARGIN), 0.0005)    # geom margin of class "adroit" (:22)
    # gravity compensation on every hand body (adroit_hand.py:31-37)
    hand_bodies = [i for i, n in enumerate(c.body_names) if n.startswith("adroit_hand/")]
    np.testing.assert_allclose(np.asarray(c.BODY_XFRC).reshape(-1, 6)[hand_bodies, 2], 9.81 * np.asarray(c.BODY_MASS)[hand_bodies])


def test_capsule_capsule_against_brute_force(reach_adroit_model):
    """The oracle's mjc_CapsuleCapsule restatement against a dense scan of the two axis segments: the contact distance is
    the minimum segment distance minus the radii, the normal points from geom 1 to geom 2 along the closest points."""
    from oracle import oracle as orc
    c = reach_adroit_model
    d = orc.OracleData(orc.OracleModel(c))
    st = reach_adroit_states(c, 11, 64)
    t = np.linspace(-1, 1, 401)
    checked = 0
    for e in range(64):
        d.set_state(st["qpos"][e], np.zeros(c.nv), np.zeros(c.nu), np.zeros(c.nv)); d.forward()
        con = d.contacts()
        gx, gm = d.get("geom_xpos").reshape(-1, 3), d.get("geom_xmat").reshape(-1, 3, 3)
        for k in range(len(con["dist"])):
            g1, g2 = con["geom"][k]
            if c.GEOM_TYPE[g1] != 3 or c.GEOM_TYPE[g2] != 3:
                continue
            p1 = gx[g1] + np.outer(t * c.GEOM_SIZE[g1, 1], gm[g1][:, 2])
            p2 = gx[g2] + np.outer(t * c.GEOM_SIZE[g2, 1], gm[g2][:, 2])
            dd = np.linalg.norm(p1[:, None, :] - p2[None, :, :], axis=2)
            i, j = np.unravel_index(dd.argmin(), dd.shape)
            assert abs(con["dist"][k] - (dd[i, j] - c.GEOM_SIZE[g1, 0] - c.GEOM_SIZE[g2, 0])) < 2e-5
            n = (p2[j] - p1[i]) / dd[i, j]
            assert np.dot(n, con["frame"][k][0]) > 1 - 1e-3
            checked += 1
    assert checked >= 30


def test_reach_adroit_kernel_bodies_match_oracle(reach_adroit_model):
    c = reach_adroit_model
    m = E.EmuModel(c)
    for thumb, seed in ((False, 3), (True, 4)):
        st = reach_adroit_states(c, seed, 128, thumb_on_palm=thumb)
        ref = statelib.oracle_step(c, st, nsub=1)
        kinds = {(int(c.GEOM_TYPE[a]), int(c.GEOM_TYPE[b])) for ps in ref["pairs"] for a, b in ps}
        assert (3, 3) in kinds and (not thumb or (3, 6) in kinds), kinds      # thumb states exercise the capsule-box pair
        assert (ref["ncon"] > 0).mean() > 0.5
        out = m.run(st["qpos"], st["qvel"], st["ctrl"], st["warm"], nsub=1, dtype=np.float64)
        assert (out["status"] == 0).all() and (out["ncon"] == ref["ncon"]).all()
        rel = np.abs(out["qvel"][:, :c.nv] - ref["qvel"]).max(axis=1) / np.maximum(1, np.abs(ref["qvel"]).max(axis=1))
        assert rel.max() < 1e-9, rel.max()
        out32 = m.run(st["qpos"], st["qvel"], st["ctrl"], st["warm"], nsub=1, dtype=np.float32)
        rel32 = np.abs(out32["qvel"][:, :c.nv] - ref["qvel"]).max(axis=1) / np.maximum(1, np.abs(ref["qvel"]).max(axis=1))
        assert (out32["ncon"] == ref["ncon"]).mean() >= 0.99 and (rel32 <= 1e-4).mean() >= 0.98     # north-star single-step tolerance


def _reach_params(c, dense):
    import math
    from dexterity_b200 import envs
    p = envs.TaskParams()
    cfg = envs.REACH
    p.task, p.successes_needed, p.steps_before_changing_goal, p.fall_termination, p.has_max_time = 1, cfg["successes_needed"], 5, 0, 1
    p.nhand, p.nsite = 24, 5
    for i, t in enumerate(scenes.ADROIT_FINGERTIP_SITES):
        p.site_id[i] = c.site_names.index("adroit_hand/" + t)
    p.dense_reward = int(dense)
    p.reach_tanh_weight = math.atanh(math.sqrt(cfg["reward_loss_at_margin"])) / cfg["reward_margin"]
    p.success_threshold, p.max_time_per_goal, p.time_limit = cfg["success_threshold"], cfg["max_time_per_goal"], float("inf")
    return p


@pytest.mark.parametrize("dense", [True, False])
def test_reach_task_kernels_on_the_adroit_scene(reach_adroit_model, dense):
    """The fused pre- / post-step kernels of the reach task (reference manipulation/tasks/reach.py:149-210, task.py:137-204) on the
    registered scene: 117 observation floats, per-finger goal distances, dense / sparse reward, success bookkeeping, checked
    against the numpy restatement (oracle/task_oracle.py) at the kernels' own state."""
    from oracle import oracle as orc
    from oracle import task_oracle as T
    c = reach_adroit_model
    m = E.EmuModel(c)
    n = 4
    env = E.EmuEnv(m, _reach_params(c, dense), nenv=n, dtype=np.float64, seed=3)
    env.reset()
    a = env.a
    assert (a["TASK_I"][:, 6] == 0).all() and (a["REWARD"] == 0).all() and (a["DISCOUNT"] == 1).all()
    rng = np.random.RandomState(1)
    lo, hi = c.JNT_RANGE[:, 0], c.JNT_RANGE[:, 1]
    d = orc.OracleData(orc.OracleModel(c))
    sites = [c.site_names.index("adroit_hand/" + t) for t in scenes.ADROIT_FINGERTIP_SITES]
    # goals = fingertip positions of a configuration near the mid-range (what FingertipCartesianPosition.next_goal produces);
    # env 0 gets the goal of its OWN start configuration so that it scores successes
    goal_q = 0.5 * (lo + hi) + 0.05 * (hi - lo) * rng.randn(n, 24)
    goal_q[0] = a["QPOS"][0, :24]
    for e in range(n):
        d.set_state(goal_q[e], np.zeros(24), np.zeros(24), np.zeros(24)); d.kinematics()
        a["GOAL"][e, :15] = T.hand_observations(c, d, 24, sites)["fingertip_positions"]
    machines = [T.GoalTaskState(0.01, 50, 5, 150 * 0.02) for _ in range(n)]
    for m_ in machines:
        m_.initialize_episode(0.0)
    for it in range(4):
        act = np.clip(goal_q + (0.3 * rng.randn(n, 24) if it % 2 else 0.0), c.ACT_CTRLRANGE[:, 0], c.ACT_CTRLRANGE[:, 1])
        act[0] = a["QPOS"][0, :24]     # holds its pose: stays on its goal
        env.step(act, nsub=1)
        for e in range(n):
            assert a["TASK_I"][e, 6] == 1
            d.set_state(a["QPOS"][e, :24], a["QVEL"][e, :24], a["CTRL"][e, :24], np.zeros(24)); d.step1()
            ho = T.hand_observations(c, d, 24, sites)
            np.testing.assert_allclose(a["OBS"][e, :48], ho["joint_positions_sin_cos"], atol=1e-12)
            np.testing.assert_allclose(a["OBS"][e, 48:72], ho["joint_velocities"], atol=1e-12)
            np.testing.assert_allclose(a["OBS"][e, 72:87], ho["fingertip_positions"], atol=1e-12)
            np.testing.assert_allclose(a["OBS"][e, 87:102], ho["fingertip_linear_velocities"], atol=1e-10)
            np.testing.assert_allclose(a["OBS"][e, 102:117], a["GOAL"][e, :15], atol=0)          # goal_state observable (task.py:206-216)
            dist = np.linalg.norm((a["GOAL"][e, :15] - ho["fingertip_positions"]).reshape(5, 3), axis=1)   # fingertip_position.py:126-130
            np.testing.assert_allclose(a["TASK_F"][e, 2:7], dist, atol=1e-10)
            np.testing.assert_allclose(a["REWARD"][e], T.reach_reward(dist, dense=dense), rtol=1e-9, atol=1e-12)
            m_ = machines[e]
            m_.before_step(a["TIME"][e] - 0.02)
            m_.after_step(dist, a["TIME"][e])
            assert a["TASK_I"][e, 0] == m_.successes and a["TASK_I"][e, 1] == m_.success_change_counter
            assert a["DISCOUNT"][e] == m_.get_discount()
    assert a["TASK_I"][0, 0] >= 1 and a["REWARD"][0] == 0.0       # the env that sits on its goal
    assert (a["TASK_I"][1:, 0] == 0).all()
