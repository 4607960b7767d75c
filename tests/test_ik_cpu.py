"""IK: numpy restatement of the reference solver vs the k_ik kernel body (emulated), plus the
reference's own properties (ik_solver_test.py:18-65) on the restatement."""
import numpy as np
import pytest

from dexterity_b200 import hands, ik
from dexterity_b200.model import scenes
from oracle import ik_oracle
from tests.emu import emu as E


@pytest.fixture(scope="module")
def setup():
    c = scenes.load_compiled("shadow_hand_only")
    h = hands.DexterousHand(c)
    sites = [c.site_names.index(f"shadow_hand_e/{t}_site") for t in scenes.SHADOW_FINGERTIPS]
    return c, h, sites


def _params(sites, **kw):
    p = ik.IkParams()
    p.nsite = 5
    for i, s in enumerate(sites):
        p.site_id[i] = s
    p.max_steps, p.num_attempts, p.early_stop, p.stop_on_first = 100, 30, 1, 1
    p.linear_tol, p.gain, p.regularization, p.progress_threshold, p.seed = 1e-3, 0.95, 1e-5, 20.0, 7
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _reachable_targets(c, h, sites, n, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    q = h.sample_joint_angles(n, g, dtype=torch.float64).numpy()
    return q, np.array([ik_oracle.site_positions(c.arrays, qq, sites)[0].ravel() for qq in q])


def test_restatement_properties(setup):
    """reference ik_solver_test.py: wrong shape raises, unreachable -> None, reachable targets are
    solved within tolerance and joint limits."""
    c, h, sites = setup
    with pytest.raises(ValueError):
        ik_oracle.solve(c.arrays, sites, np.zeros((4, 3)))
    assert ik_oracle.solve(c.arrays, sites, np.full((5, 3), 10.0), num_attempts=2, early_stop=True) is None
    _, targets = _reachable_targets(c, h, sites, 4, seed=12345)
    rs = np.random.RandomState(12345)
    for t in targets:
        q = ik_oracle.solve(c.arrays, sites, t.reshape(5, 3), early_stop=True, stop_on_first_successful_attempt=True,
                            random_state=rs)
        assert q is not None
        x, _ = ik_oracle.site_positions(c.arrays, q, sites)
        assert np.linalg.norm(x - t.reshape(5, 3), axis=1).max() <= 1e-3
        assert (q >= c.JNT_RANGE[:, 0] - 1e-12).all() and (q <= c.JNT_RANGE[:, 1] + 1e-12).all()


def test_kernel_body_matches_restatement_first_attempt(setup):
    """With a single attempt the start (mid-range) is deterministic, so the kernel body and the
    restatement must follow the same iterates."""
    c, h, sites = setup
    em = E.EmuModel(c)
    _, targets = _reachable_targets(c, h, sites, 6, seed=3)
    p = _params(sites, num_attempts=1)
    qk, okk, _ = E.emu_ik(em, p, targets, dtype=np.float64)
    for i, t in enumerate(targets):
        mid = c.JNT_RANGE.mean(axis=1)
        qr, errs = ik_oracle.solve_ik(c.arrays, sites, t.reshape(5, 3), mid, 1e-3, 100, True)
        ok_r = all(e <= 1e-3 for e in errs)
        assert okk[i] == ok_r
        if ok_r:
            np.testing.assert_allclose(qk[i, :24], qr, atol=1e-8)


def test_kernel_body_restarts_and_unreachable(setup):
    c, h, sites = setup
    em = E.EmuModel(c)
    _, targets = _reachable_targets(c, h, sites, 6, seed=11)
    targets = np.vstack([targets, np.full((1, 15), 10.0)])
    p = _params(sites, num_attempts=8)
    q, ok, attempts = E.emu_ik(em, p, targets, dtype=np.float32)
    assert not ok[-1] and attempts[-1] == 8
    assert ok[:-1].sum() >= 5
    for i in np.where(ok)[0]:
        x, _ = ik_oracle.site_positions(c.arrays, q[i, :24].astype(np.float64), sites)
        assert np.linalg.norm(x - targets[i].reshape(5, 3), axis=1).max() <= 1e-3 + 1e-5
        assert (q[i, :24] >= c.JNT_RANGE[:, 0] - 1e-6).all() and (q[i, :24] <= c.JNT_RANGE[:, 1] + 1e-6).all()


# reference models/hands/hands_test.py:195-228: golden fingertip positions of the Adroit hand (its own root model),
# which IKSolver must reach within 1e-3 and FK must reproduce at the solution
ADROIT_GOLDEN_FINGERTIPS = np.asarray([
    [-0.003572, -0.020904, 0.371999],
    [-0.028277, -0.036063, 0.391271],
    [-0.052305, -0.006066, 0.393481],
    [-0.089808, -0.042816, 0.423813],
    [0.026246, -0.017261, 0.416314],
])


@pytest.fixture(scope="module")
def adroit():
    asset, prefix, tips = scenes.HANDS["adroit_hand"]
    c = scenes.load_compiled(asset)
    h = hands.DexterousHand(c, prefix, tips)
    return c, h, h.fingertip_site_ids


def test_adroit_model_facts(adroit):
    """reference models/hands/adroit_hand_constants.py: 24 joints, 24 actuators (fully actuated), 5 fingertip sites."""
    c, h, sites = adroit
    assert (c.nq, c.nv, c.nu, len(sites)) == (24, 24, 24, 5)
    assert h.joint_names[:3] == ("WRJ1", "WRJ0", "FFJ3") and h.joint_names[-1] == "THJ0"
    assert h.actuator_names[0] == "A_WRJ1" and not h.underactuated
    assert np.array_equal(h.POSITION_TO_CONTROL, np.eye(24))          # ctrl = qpos (adroit_hand.py:97-113)


def test_adroit_golden_fingertips_restatement(adroit):
    """The reference's golden vector through the numpy restatement of its solver."""
    c, h, sites = adroit
    q = ik_oracle.solve(c.arrays, sites, ADROIT_GOLDEN_FINGERTIPS, linear_tol=1e-3, early_stop=True,
                        stop_on_first_successful_attempt=True, random_state=np.random.RandomState(0))
    assert q is not None
    x, _ = ik_oracle.site_positions(c.arrays, q, sites)
    np.testing.assert_allclose(x, ADROIT_GOLDEN_FINGERTIPS, atol=1e-3)
    assert (q >= c.JNT_RANGE[:, 0] - 1e-12).all() and (q <= c.JNT_RANGE[:, 1] + 1e-12).all()


def test_adroit_golden_fingertips_kernel_body(adroit):
    """The same golden vector through the k_ik kernel body (emulated), FP64 and FP32."""
    c, h, sites = adroit
    em = E.EmuModel(c)
    for dtype in (np.float64, np.float32):
        q, ok, attempts = E.emu_ik(em, _params(sites), ADROIT_GOLDEN_FINGERTIPS.reshape(1, 15), dtype=dtype)
        assert ok[0]
        x, _ = ik_oracle.site_positions(c.arrays, q[0, :24].astype(np.float64), sites)
        np.testing.assert_allclose(x, ADROIT_GOLDEN_FINGERTIPS, atol=1e-3)
