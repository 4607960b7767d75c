"""MJCF export (dexterity_b200/model/export.py; reference scripts/export_task.py:31-46): the exported flat MJCF + assets,
parsed back and compiled, must reproduce the committed device model (integers exactly, reals to 1e-11: the palm-upright quaternion of the reference, 0.707106781186, is not normalised and the parser normalises it).  Needs the reference checkout for the
vendored mesh / XML files, like every compile step (skipped where it is absent, e.g. on the GPU box)."""
import os
import zipfile

import numpy as np
import pytest

from dexterity_b200.model import compile as compile_lib
from dexterity_b200.model import export, mjcf, scenes

pytestmark = pytest.mark.skipif(not os.path.isdir(scenes.reference_root()), reason="reference model files not present")


@pytest.mark.parametrize("name", ["reorient_shadow", "reach_shadow"])
def test_export_round_trip_reproduces_the_compiled_model(name, tmp_path):
    model, kw = scenes.ASSEMBLIES[name]()
    model.name = name
    zip_path = str(tmp_path / (name + ".zip"))
    xml = export.export_with_assets(model, str(tmp_path / "out"), timestep=kw["timestep"], zip_path=zip_path)
    with zipfile.ZipFile(zip_path) as z:                       # what a MuJoCo user would unpack
        names = z.namelist()
        assert name + ".xml" in names and sum(n.startswith("assets/") for n in names) > 100
        z.extractall(str(tmp_path / "unzipped"))
    back = mjcf.load(str(tmp_path / "unzipped" / (name + ".xml")), name=name)
    again = compile_lib.compile_model(back, **kw)
    committed = scenes.load_compiled(name)
    assert again.body_names == committed.body_names and again.geom_names == committed.geom_names
    assert again.joint_names == committed.joint_names and again.actuator_names == committed.actuator_names
    for k, v in committed.arrays.items():
        if np.issubdtype(v.dtype, np.floating):      # quaternions are re-normalised on load: last-bit differences only
            assert again.arrays[k].shape == v.shape and np.allclose(again.arrays[k], v, rtol=1e-11, atol=1e-11), k
        else:
            assert np.array_equal(again.arrays[k], v), k
    text = open(xml).read()
    assert "<default" not in text and 'type="free"' in text or name == "reach_shadow"
