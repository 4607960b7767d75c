"""IK throughput (BASELINE.json configs[3]): 65536 target sets, reference test parameters."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dexterity_b200 import ik
from oracle import ik_oracle
s = ik.IKSolver(seed=12345)
g = torch.Generator().manual_seed(12345)
q = s.hand.sample_joint_angles(2048, g, dtype=torch.float64).numpy()
t = np.array([ik_oracle.site_positions(s.compiled.arrays, qq, s._sites)[0] for qq in q])
T = torch.as_tensor(t, dtype=torch.float32).repeat(32, 1, 1).cuda()
for _ in range(2):
    s.solve(T, early_stop=True, stop_on_first_successful_attempt=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); K = 5
for _ in range(K):
    qq, ok = s.solve(T, early_stop=True, stop_on_first_successful_attempt=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("IK 65536 sets: %.2f ms -> %.3e solves/s, success %.4f, mean attempts %.2f" % (ms, 65536 / ms * 1e3, ok.float().mean().item(), s.last_attempts.float().mean().item()))
t0 = time.time()
rs = np.random.RandomState(0)
n = 0
while time.time() - t0 < 8:
    ik_oracle.solve(s.compiled.arrays, s._sites, t[n], early_stop=True, stop_on_first_successful_attempt=True, random_state=rs); n += 1
print("numpy restatement (1 core): %.2f solves/s" % (n / (time.time() - t0)))
