"""GPU: the Adroit hand's golden fingertip vector of the reference (models/hands/hands_test.py:195-228) through
dex_ik_solve, and its contact-free articulated dynamics against the oracle (see tests/test_adroit_cpu.py)."""
import numpy as np
import pytest
import torch

from tests import statelib
from tests.test_adroit_cpu import adroit_states
from tests.test_ik_cpu import ADROIT_GOLDEN_FINGERTIPS

pytestmark = pytest.mark.gpu


def test_adroit_golden_fingertips_gpu():
    from dexterity_b200 import ik
    from oracle import ik_oracle
    solver = ik.IKSolver(hand_name="adroit_hand")
    assert solver.hand.num_joints == 24 and solver.hand.num_actuators == 24 and solver.num_sites == 5
    for dtype in (torch.float32, torch.float64):
        q = solver.solve(ADROIT_GOLDEN_FINGERTIPS, linear_tol=1e-3, early_stop=True,
                         stop_on_first_successful_attempt=True, dtype=dtype)
        assert q is not None                                  # hands_test.py:218 assertIsNotNone
        c = solver.compiled
        x, _ = ik_oracle.site_positions(c.arrays, q, solver._sites)
        np.testing.assert_allclose(x, ADROIT_GOLDEN_FINGERTIPS, atol=1e-3)   # hands_test.py:224
        assert (q >= c.JNT_RANGE[:, 0] - 1e-6).all() and (q <= c.JNT_RANGE[:, 1] + 1e-6).all()
    with pytest.raises(ValueError):
        ik.IKSolver(hand_name="mpl_hand")


def test_adroit_dynamics_match_oracle_gpu():
    from dexterity_b200 import backend
    from dexterity_b200.model import scenes
    c = scenes.load_compiled("adroit_hand_only")
    st = adroit_states(c, 3, 256)
    ref = statelib.oracle_step(c, st, nsub=1)
    model = backend.PhysicsModel(c, 0)
    for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 1e-4)):
        b = backend.PhysicsBatch(model, 256, dtype)
        statelib.upload(b, st)
        b.step(1)
        torch.cuda.synchronize()
        assert (b.status.cpu().numpy() == 0).all() and (b.ncon.cpu().numpy() == 0).all()
        for name, r in (("qpos", ref["qpos"]), ("qvel", ref["qvel"])):
            x = getattr(b, name).cpu().numpy().astype(np.float64)[:, :r.shape[1]]
            rel = np.abs(x - r).max(axis=1) / np.maximum(1.0, np.abs(r).max(axis=1))
            assert rel.max() <= tol, (dtype, name, rel.max())
