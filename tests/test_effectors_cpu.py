"""Effector wrappers: reference effectors/wrappers/{smooth_action,previous_action}_test.py,
effectors/hand_effector_test.py:12-20 and task_test.py:114-139 (action-spec content)."""
import numpy as np
import pytest
import torch

from dexterity_b200 import effectors, hands


class FakePhysics:
    def __init__(self, n, nu):
        self.num_envs, self.dtype, self.device = n, torch.float64, torch.device("cpu")
        self.ctrl = torch.zeros(n, nu, dtype=torch.float64)


def test_hand_effector_writes_ctrl_and_spec(reorient_model):
    h = hands.DexterousHand(reorient_model)
    eff = effectors.HandEffector(h, "shadow_hand_e")
    spec = eff.action_spec()
    assert spec.shape == (20,) and spec.dtype == np.float32 and eff.prefix == "shadow_hand_e_joint"
    assert spec.name.split("\t")[0] == "shadow_hand_e_joint0"
    np.testing.assert_allclose(spec.minimum[0], -0.488692, atol=1e-6)      # shadow_hand_position_actuators.xml:26
    np.testing.assert_allclose(spec.maximum[4], np.pi, atol=1e-6)          # tendon actuator A_FFJ1
    phys = FakePhysics(3, 20)
    cmd = torch.rand(3, 20, dtype=torch.float64) * 10
    eff.set_control(phys, cmd)
    torch.testing.assert_close(phys.ctrl, cmd)                             # unclipped


@pytest.mark.parametrize("alpha,expect", [(0.0, 3.0), (0.9, 0.9 * 1 + 0.1 * 3), (1.0, 1.0)])
def test_smooth_action_kat(reorient_model, alpha, expect):
    h = hands.DexterousHand(reorient_model)
    eff = effectors.SmoothAction(effectors.HandEffector(h, "h"), alpha=alpha)
    phys = FakePhysics(2, 20)
    eff.initialize_episode(phys)
    eff.set_control(phys, torch.full((2, 20), 1.0, dtype=torch.float64))
    assert (phys.ctrl == 1.0).all()
    eff.set_control(phys, torch.full((2, 20), 3.0, dtype=torch.float64))
    torch.testing.assert_close(phys.ctrl, torch.full((2, 20), expect, dtype=torch.float64), rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        effectors.SmoothAction(effectors.HandEffector(h, "h"), alpha=1.5)


def test_previous_action_is_a_copy(reorient_model):
    h = hands.DexterousHand(reorient_model)
    eff = effectors.PreviousAction(effectors.HandEffector(h, "h"))
    phys = FakePhysics(2, 20)
    eff.initialize_episode(phys)
    assert (eff.previous_action == 0).all()
    cmd = torch.ones(2, 20, dtype=torch.float64)
    eff.set_control(phys, cmd)
    cmd += 5
    assert (eff.previous_action == 1).all()
