"""Histogram of the narrowphase work items (GPU box): hull sizes / geom types of the convex-convex candidates."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dexterity_b200 import envs
N = 16384
env = envs.load("reorient", "state_dense", seed=12345, num_envs=N)
lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
env.reset()
c = env.physics.model.compiled
vn = np.asarray(c.GEOM_VERTNUM); gt = np.asarray(c.GEOM_TYPE); g1 = np.asarray(c.PAIR_GEOM1); g2 = np.asarray(c.PAIR_GEOM2)
for it in range(30):
    env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
    if it % 10 == 9:
        w = env.physics.tensor("WORK").cpu().numpy()
        n = int(w[0]); items = w[16:16 + 2 * n].reshape(n, 2)
        pair = items[:, 1] & 0xffffff
        box = gt[g1[pair]] == 6
        na = np.where(box, 0, vn[g1[pair]]); nb = vn[g2[pair]]
        mx = np.maximum(na, nb)
        ncon = env.physics.ncon.float().mean().item()
        print("step", it, "items", n, "per env %.2f" % (n / N), "box-mesh share %.2f" % box.mean(), "ncon/env %.2f" % ncon)
        edges = [0, 16, 32, 64, 128, 256, 512]
        for nm, sel in (("box-mesh", box), ("mesh-mesh", ~box)):
            h, _ = np.histogram(mx[sel], bins=edges)
            print("   ", nm, "max-hull-size hist", dict(zip(["<16", "<32", "<64", "<128", "<256", "<512"], h.tolist())),
                  "sum of (nA+nB): %.1f per item" % ((na + nb)[sel].mean() if sel.any() else 0))
