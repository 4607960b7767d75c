"""The reference's MuJoCo-independent known-answer tests, replayed on this build's host-side
mirrors and on the numpy restatement used as the task-logic checker."""
import numpy as np
import pytest

from dexterity_b200 import hands
from oracle import task_oracle as T


def test_projection_matrices_identity(reorient_model):
    """reference models/hands/hands_test.py:26-31 (exact)."""
    h = hands.DexterousHand(reorient_model)
    assert h.num_joints == 24 and h.num_actuators == 20 and h.underactuated
    np.testing.assert_array_equal(h.POSITION_TO_CONTROL @ h.CONTROL_TO_POSITION, np.eye(20))
    # reference shadow_hand_e_constants.py:127-154: the four *FJ1 actuators drive (*FJ1, *FJ0)
    coupled = [tuple(h.joint_names[j] for j in js) for js in h.actuator_joint_mapping if len(js) > 1]
    assert coupled == [("FFJ1", "FFJ0"), ("MFJ1", "MFJ0"), ("RFJ1", "RFJ0"), ("LFJ1", "LFJ0")]


def test_wrong_shapes_raise(reorient_model):
    """reference hands_test.py:71-85."""
    h = hands.DexterousHand(reorient_model)
    with pytest.raises(ValueError):
        h.control_to_joint_positions(np.zeros(19))
    with pytest.raises(ValueError):
        h.joint_positions_to_control(np.zeros(25))


def test_reorient_reward_kat():
    """reference manipulation/tasks/reorient_test.py:44-50 (assert_equal, exact): at the goal the
    orientation reward is 1/0.1, the success bonus 1.0 and the smoothing term ||ctrl||^2."""
    goal = np.array([0.5, 0.5, 0.5, 0.5])
    d = T.prop_orientation_goal_distance(goal, goal)[0]
    ctrl = np.linspace(-1, 1, 20)
    terms = T.reorient_reward_terms(d, ctrl)
    assert terms["orientation"][0] == 1.0 / 0.1
    assert terms["success_bonus"][0] == 1.0
    assert terms["action_smoothing"][0] == np.linalg.norm(ctrl) ** 2
    assert T.weighted_average(terms) == 10.0 + 800.0 - 0.1 * np.linalg.norm(ctrl) ** 2


def test_goal_distance_is_rotation_angle():
    rng = np.random.RandomState(0)
    for _ in range(50):
        ang = rng.uniform(0, np.pi)
        ax = rng.randn(3); ax /= np.linalg.norm(ax)
        q0 = rng.randn(4); q0 /= np.linalg.norm(q0)
        dq = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax])
        q1 = T.quat_mul(dq, q0)
        for sign in (1, -1):      # double cover: -q is the same rotation
            assert abs(T.prop_orientation_goal_distance(sign * q1, q0)[0] - ang) < 1e-7


def test_smooth_action_kat():
    """reference effectors/wrappers/smooth_action_test.py:25-44,82-94."""
    x0, x1 = np.full(3, 1.0), np.full(3, 3.0)
    for alpha, expect in ((0.0, 3.0), (0.9, 0.9 * 1 + 0.1 * 3), (1.0, 1.0)):
        v = T.smooth_action(None, x0, alpha)
        np.testing.assert_array_equal(v, x0)
        np.testing.assert_allclose(T.smooth_action(v, x1, alpha), np.full(3, expect), rtol=0, atol=1e-15)


def test_goal_task_state_machine_quirks():
    """cumulative counter, strict '>', timeout only outside the
    threshold, first success terminates reorient with discount 0."""
    s = T.GoalTaskState(success_threshold=0.1, successes_needed=1, steps_before_changing_goal=5, max_time_per_goal=7.5)
    s.initialize_episode(0.0)
    s.after_step(np.array([0.5]), 0.025)
    assert not s.should_terminate_episode() and s.get_discount() == 1.0
    s.after_step(np.array([0.5]), 7.6)
    assert s.exceeded_single_goal_time and s.should_terminate_episode() and s.get_discount() == 1.0
    s.initialize_episode(0.0)
    s.after_step(np.array([0.1]), 0.025)          # inclusive threshold
    assert s.successes == 1 and s.should_terminate_episode() and s.get_discount() == 0.0
    r = T.GoalTaskState(success_threshold=0.01, successes_needed=50, steps_before_changing_goal=5, max_time_per_goal=3.0)
    r.initialize_episode(0.0)
    for k in range(5):
        r.after_step(np.zeros(5), 0.02 * k)
        assert not r.before_step(0.02 * k)
    r.after_step(np.ones(5), 0.1)                  # leaving the threshold does not reset the counter
    r.after_step(np.zeros(5), 0.12)
    assert r.success_change_counter == 6 and r.before_step(0.12) and r.successes == 1


def test_reach_reward_restated():
    """reference manipulation/tasks/reach.py:196-210 and reach_test.py:12-24 (sparse = -1 away)."""
    assert T.reach_reward(np.full(5, 0.2), dense=False) == -1.0
    assert T.reach_reward(np.zeros(5), dense=True) == 0.0
    assert abs(T.tanh_squared(0.1, margin=0.1) - 0.95) < 1e-12
