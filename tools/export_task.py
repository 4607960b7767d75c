"""Exports an assembled scene as MJCF + assets (zip), the artefact of reference scripts/export_task.py:31-46.
Needs the reference checkout for the vendored model files (DEXTERITY_REFERENCE_ROOT, default /root/reference).

    python tools/export_task.py reorient_shadow /tmp/reorient        # -> /tmp/reorient/reorient_shadow.zip
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dexterity_b200.model import export, scenes

name = sys.argv[1] if len(sys.argv) > 1 else "reorient_shadow"
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join("/tmp", name + "_mjcf")
model, kw = scenes.ASSEMBLIES[name]()
model.name = name
xml = export.export_with_assets(model, out, timestep=kw["timestep"], zip_path=os.path.join(out, name + ".zip"))
print("wrote", xml, "and", os.path.join(out, name + ".zip"), os.path.getsize(os.path.join(out, name + ".zip")), "bytes")
