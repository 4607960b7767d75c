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
