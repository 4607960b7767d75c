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


def test_reach_adroit_contacts_match_oracle_gpu():
    """The registered reach scene on the GPU: capsule-capsule (closed form) and capsule-box (MPR) contacts of the hand's explicit
    condim-1 pairs, one frictionless row per contact: contact counts equal, single-step state within the north-star tolerance."""
    from dexterity_b200 import backend
    from dexterity_b200.model import scenes
    from tests.test_adroit_cpu import reach_adroit_states
    c = scenes.load_compiled("reach_adroit")
    model = backend.PhysicsModel(c, 0)
    for thumb, seed in ((False, 5), (True, 6)):
        st = reach_adroit_states(c, seed, 1024, thumb_on_palm=thumb)
        ref = statelib.oracle_step(c, st, nsub=1)
        assert (ref["ncon"] > 0).mean() > 0.5
        for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 1e-4)):
            b = backend.PhysicsBatch(model, 1024, dtype)
            statelib.upload(b, st)
            b.step(1)
            torch.cuda.synchronize()
            assert (b.status.cpu().numpy() == 0).all()
            same = b.ncon.cpu().numpy() == ref["ncon"]
            assert same.mean() >= (1.0 if dtype == torch.float64 else 0.995), same.mean()
            v = b.qvel.cpu().numpy().astype(np.float64)
            rel = np.abs(v - ref["qvel"]).max(axis=1) / np.maximum(1.0, np.abs(ref["qvel"]).max(axis=1))
            print(f"reach_adroit thumb={thumb} {dtype}: median {np.median(rel):.2e} max {rel.max():.2e} share <= tol {(rel <= tol).mean():.4f}")
            assert (rel[same] <= tol).mean() >= (1.0 if dtype == torch.float64 else 0.99), (dtype, rel.max())


def test_registered_reach_is_the_adroit_hand():
    """`load("reach", ...)` without a hand builds the reference's registered task (reach.py:223-269: AdroitHand): 24 actuators,
    117 observation floats, and the behaviour reach_test.py:12-35 checks: sparse reward -1 away from the goal, 0 once the
    position-controlled hand (ctrl = goal qpos, adroit_hand.py:104-111) has reached it."""
    from dexterity_b200 import envs
    env = envs.load("reach", "state_sparse", seed=12345, num_envs=64)
    assert env.hand_name == "adroit_hand" and env.action_spec().shape == (24,)
    assert not env.hand.underactuated
    ts = env.reset()
    torch.cuda.synchronize()
    assert (ts.step_type == 0).all()
    assert sum(s.shape[0] for s in env.observation_spec().values()) == 117
    g = env.goal[:, :15]
    assert torch.isfinite(g).all() and (g.reshape(-1, 5, 3).norm(dim=-1) < 0.6).all()
    ts = env.step(torch.zeros(64, 24, device=env.device))
    far = env.goal_distance.min(dim=1).values > 0.01          # every fingertip further than the success threshold
    assert far.any() and (ts.reward[far] == -1.0).all()
    reached = torch.zeros(64, dtype=torch.bool, device=env.device)
    for _ in range(150):
        ts = env.step(env.hand.joint_positions_to_control(env.goal_qpos))
        reached |= (env.successes > 0) & (ts.reward == 0.0)
    assert (env.physics.status == 0).all()
    assert reached.float().mean() > 0.7, reached.float().mean()
