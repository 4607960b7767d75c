"""Hand description derived from the compiled model (batched mirror of the reference's
`ShadowHandSeriesE` accessors; reference models/hands/shadow_hand_e.py:97-129,
shadow_hand_e_constants.py:127-187)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch


class DexterousHand:
    def __init__(self, compiled, prefix: str = "shadow_hand_e/", fingertip_sites=None) -> None:
        c = compiled
        self.compiled, self.prefix = c, prefix
        self.joint_ids = [i for i, n in enumerate(c.joint_names) if n.startswith(prefix)]
        self.joint_names = tuple(c.joint_names[i][len(prefix):] for i in self.joint_ids)
        self.actuator_ids = [i for i, n in enumerate(c.actuator_names) if n.startswith(prefix)]
        self.actuator_names = tuple(c.actuator_names[i][len(prefix):] for i in self.actuator_ids)
        if fingertip_sites is None:
            self.fingertip_site_ids = [i for i, n in enumerate(c.site_names) if n.startswith(prefix) and n.endswith("tip_site")]
        else:   # explicit names, in the reference's order (e.g. adroit_hand_constants.py:86-92)
            self.fingertip_site_ids = [c.site_names.index(prefix + n) for n in fingertip_sites]
        nj, na = len(self.joint_ids), len(self.actuator_ids)
        jpos = {j: k for k, j in enumerate(self.joint_ids)}
        # actuator -> joints it drives (a tendon actuator drives every joint of its fixed tendon)
        self.actuator_joint_mapping: List[Tuple[int, ...]] = []
        for a in self.actuator_ids:
            if c.ACT_TRNTYPE[a] == 0:
                self.actuator_joint_mapping.append((jpos[int(c.ACT_TRNID[a])],))
            else:
                t = int(c.ACT_TRNID[a])
                w = range(int(c.TEN_ADR[t]), int(c.TEN_ADR[t] + c.TEN_NUM[t]))
                self.actuator_joint_mapping.append(tuple(sorted(jpos[int(c.WRAP_JNT[k])] for k in w)))
        p2c = np.zeros((na, nj)); c2p = np.zeros((nj, na))
        self.coupled_joint_ids = []
        for a, joints in enumerate(self.actuator_joint_mapping):
            p2c[a, list(joints)] = 1.0
            c2p[list(joints), a] = 1.0 / len(joints)
            if len(joints) > 1:
                self.coupled_joint_ids.append(list(joints))
        self.POSITION_TO_CONTROL, self.CONTROL_TO_POSITION = p2c, c2p
        self.joint_range = c.JNT_RANGE[self.joint_ids]

    @property
    def num_joints(self) -> int: return len(self.joint_ids)
    @property
    def num_actuators(self) -> int: return len(self.actuator_ids)
    dofs = num_joints
    @property
    def underactuated(self) -> bool: return self.num_actuators < self.num_joints

    def control_to_joint_positions(self, control):
        """[.., nu] -> [.., nq_hand]: coupled joints share their actuator's command evenly."""
        if control.shape[-1] != self.num_actuators:
            raise ValueError(f"Expected control of shape ({self.num_actuators}), got {tuple(control.shape)}")
        M = self.CONTROL_TO_POSITION
        return control @ (torch.as_tensor(M.T, dtype=control.dtype, device=control.device) if torch.is_tensor(control) else M.T)

    def joint_positions_to_control(self, qpos):
        if qpos.shape[-1] != self.num_joints:
            raise ValueError(f"Expected qpos of shape ({self.num_joints}), got {tuple(qpos.shape)}")
        M = self.POSITION_TO_CONTROL
        return qpos @ (torch.as_tensor(M.T, dtype=qpos.dtype, device=qpos.device) if torch.is_tensor(qpos) else M.T)

    def sample_joint_angles(self, num: int, generator: torch.Generator, range_fraction: float = 1.0,
                            device="cpu", dtype=torch.float32) -> torch.Tensor:
        """Batched reference dexterous_hand.py:120-142 (+ coupled-joint fix shadow_hand_e.py:124-129)."""
        if not 0 <= range_fraction <= 1:
            raise ValueError("range_fraction must be between 0 and 1.")
        lo = torch.as_tensor(range_fraction * self.joint_range[:, 0], dtype=dtype, device=device)
        hi = torch.as_tensor(range_fraction * self.joint_range[:, 1], dtype=dtype, device=device)
        q = lo + (hi - lo) * torch.rand(num, self.num_joints, generator=generator, device=device, dtype=dtype)
        for ids in self.coupled_joint_ids:
            q[:, ids] = q[:, ids[-1]].clone().unsqueeze(1)
        return q
