"""Compiles the suite's scenes from the reference MJCF/STL model data (read at build time from
DEXTERITY_REFERENCE_ROOT, default /root/reference) into the committed blobs under
dexterity_b200/assets/.  Run here (where the reference checkout is mounted), never on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dexterity_b200.model import scenes

for name in ("reorient_shadow", "reach_shadow", "reach_adroit", "shadow_hand_only", "adroit_hand_only"):
    m = scenes.load_compiled(name, rebuild=True)
    print(name, dict(nbody=m.nbody, nq=m.nq, nv=m.nv, nu=m.nu, ngeom=m.ngeom, npair=m.npair, nsite=m.nsite),
          os.path.getsize(os.path.join(scenes.ASSET_DIR, name + ".npz")), "bytes")
