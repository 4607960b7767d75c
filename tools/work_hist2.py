"""Which pairs dominate the narrowphase work list, and how many would a bounding-capsule test cull? (GPU box)"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dexterity_b200 import envs
from dexterity_b200.model import mjcf
N = 16384
env = envs.load("reorient", "state_dense", seed=12345, num_envs=N)
lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
env.reset()
c = env.physics.model.compiled
vn = np.asarray(c.GEOM_VERTNUM); va = np.asarray(c.GEOM_VERTADR); gt = np.asarray(c.GEOM_TYPE); g1 = np.asarray(c.PAIR_GEOM1); g2 = np.asarray(c.PAIR_GEOM2)
gb = np.asarray(c.GEOM_BODY)
V = np.asarray(c.VERT).reshape(-1, 3)
# body-frame vertices and bounding capsules (PCA axis)
caps = {}
for g in range(c.ngeom):
    if gt[g] != 7: continue
    R = mjcf.quat_to_mat(np.asarray(c.GEOM_QUAT)[g]); p = np.asarray(c.GEOM_POS)[g]
    v = p + V[va[g]:va[g] + vn[g]] @ R.T
    ctr = v.mean(0); u, s, vt = np.linalg.svd(v - ctr); ax = vt[0]
    t = (v - ctr) @ ax; perp = (v - ctr) - np.outer(t, ax); r = np.linalg.norm(perp, axis=1).max()
    # shrink the segment so that end caps still cover: conservative simple choice: full extent, radius r
    a = ctr + ax * (t.min() + min(r, (t.max()-t.min())/2)); b = ctr + ax * (t.max() - min(r, (t.max()-t.min())/2))
    # make sure all verts inside capsule(a,b,r'): compute needed radius
    ab = b - a; L2 = max(ab @ ab, 1e-12); tt = np.clip(((v - a) @ ab) / L2, 0, 1); d = np.linalg.norm(v - (a + np.outer(tt, ab)), axis=1)
    caps[g] = (a, b, d.max())
def segdist(p1, q1, p2, q2):
    d1 = q1 - p1; d2 = q2 - p2; r = p1 - p2
    a = (d1 * d1).sum(-1); e = (d2 * d2).sum(-1); f = (d2 * r).sum(-1); cc = (d1 * r).sum(-1); b = (d1 * d2).sum(-1)
    den = a * e - b * b
    s = np.where(den > 1e-12, np.clip((b * f - cc * e) / np.maximum(den, 1e-12), 0, 1), 0.0)
    t = (b * s + f) / np.maximum(e, 1e-12)
    t2 = np.clip(t, 0, 1)
    s = np.where(t != t2, np.clip((b * t2 - cc) / np.maximum(a, 1e-12), 0, 1), s)
    c1 = p1 + d1 * s[:, None]; c2 = p2 + d2 * t2[:, None]
    return np.linalg.norm(c1 - c2, axis=1)
for it in range(20):
    env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
w = env.physics.tensor("WORK").cpu().numpy(); n = int(w[0]); items = w[16:16 + 2 * n].reshape(n, 2)
envi = items[:, 0]; pair = items[:, 1] & 0xffffff; slot = items[:, 1] >> 24
stage = env.physics.tensor("STAGE").cpu().numpy()
hit = stage[envi, slot, 7] >= 0
xk = env.physics.tensor("XKIN").cpu().numpy().reshape(N, -1, 16)
print("items", n, "hits", hit.sum())
mm = gt[g1[pair]] == 7
culled = np.zeros(n, bool)
idx = np.nonzero(mm)[0]
P1 = np.zeros((len(idx), 3)); Q1 = P1.copy(); P2 = P1.copy(); Q2 = P1.copy(); RR = np.zeros(len(idx))
for k, i in enumerate(idx):
    ga, gb_ = g1[pair[i]], g2[pair[i]]
    a1, b1, r1 = caps[ga]; a2, b2, r2 = caps[gb_]
    x1 = xk[envi[i], gb[ga]]; x2 = xk[envi[i], gb[gb_]]
    R1 = x1[7:16].reshape(3, 3); R2 = x2[7:16].reshape(3, 3)
    P1[k] = x1[:3] + R1 @ a1; Q1[k] = x1[:3] + R1 @ b1; P2[k] = x2[:3] + R2 @ a2; Q2[k] = x2[:3] + R2 @ b2; RR[k] = r1 + r2
d = segdist(P1, Q1, P2, Q2)
culled[idx] = d > RR
print("mesh-mesh items", len(idx), "capsule-culled", culled.sum(), "of which were hits (must be 0):", (culled & hit).sum())
names = c.geom_names
import collections
cnt = collections.Counter(pair.tolist())
print("distinct pairs", len(cnt))
for p, k in cnt.most_common(25):
    sel = pair == p
    print("%6d  hit %.2f capscull %.2f  %s(%d) x %s(%d)" % (k, hit[sel].mean(), culled[sel].mean(), names[g1[p]], vn[g1[p]], names[g2[p]], vn[g2[p]]))
