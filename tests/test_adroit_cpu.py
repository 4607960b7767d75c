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
