"""How concentrated is the narrowphase work list on a few geom pairs? (GPU box)"""
import os, sys, collections
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from dexterity_b200 import envs
N = 16384
torch.manual_seed(0)
env = envs.load("reorient", "state_dense", seed=12345, num_envs=N)
lo = torch.as_tensor(env.action_spec().minimum, device=env.device); hi = torch.as_tensor(env.action_spec().maximum, device=env.device)
env.reset()
for _ in range(25): env.step(lo + (hi - lo) * torch.rand(N, 20, device=env.device))
torch.cuda.synchronize()
c = env.compiled
w = env.physics.tensor("WORK").cpu().numpy(); n = int(w[0]); items = w[16:16 + 2 * n].reshape(n, 2)
envi = items[:, 0]; pair = items[:, 1] & 0xffffff; slot = items[:, 1] >> 24
stage = env.physics.tensor("STAGE").cpu().numpy()
hit = stage[envi, slot, 7] >= 0
cnt = collections.Counter(pair.tolist())
vals = np.array(sorted(cnt.values(), reverse=True))
print("items", n, "hits", int(hit.sum()), "distinct pairs", len(cnt))
cum = np.cumsum(vals) / n
for k in (10, 20, 50, 100, 200, 500, 1000):
    if k <= len(vals): print("top %4d pairs cover %.3f of the items" % (k, cum[k - 1]))
g1 = np.asarray(c.PAIR_GEOM1); g2 = np.asarray(c.PAIR_GEOM2); vn = np.asarray(c.GEOM_VERTNUM); gb = np.asarray(c.GEOM_BODY)
print("items with the cube as geom 1: %.3f" % (np.asarray(c.GEOM_TYPE)[g1[pair]] == 6).mean())
bp = collections.Counter(zip(gb[g1[pair]].tolist(), gb[g2[pair]].tolist()))
bv = np.array(sorted(bp.values(), reverse=True)); print("distinct body pairs", len(bp), "top 10/20/40 cover", (np.cumsum(bv) / n)[[9, 19, 39]])
# coherence of 32 consecutive items in list order vs sorted order: distinct pairs per warp
def distinct_per_warp(p): return np.mean([len(set(p[i:i + 32].tolist())) for i in range(0, len(p) - 31, 32)])
print("distinct pairs per 32 consecutive items: list order %.1f, sorted by pair %.1f" % (distinct_per_warp(pair), distinct_per_warp(np.sort(pair))))
print("distinct geom-B per 32 items: list order %.1f, sorted %.1f" % (distinct_per_warp(g2[pair]), distinct_per_warp(np.sort(g2[pair]))))
for p, k in cnt.most_common(12):
    sel = pair == p
    print("%6d  hit %.2f  %s(%d) x %s(%d)" % (k, hit[sel].mean(), c.geom_names[g1[p]], vn[g1[p]], c.geom_names[g2[p]], vn[g2[p]]))
