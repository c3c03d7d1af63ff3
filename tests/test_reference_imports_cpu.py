"""CPU, only where the reference checkout is present (this container; skipped on the GPU box): the
reference's own modules on the path -- gaussian_renderer/__init__.py and scene/gaussian_model.py --
import against OUR packages without modification, i.e. every name they take from
`diff_gaussian_rasterization` and `gaussian_hierarchy._C` exists here.
Third-party packages that are neither ours nor on this path (simple_knn, plyfile) are stubbed."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, types, inspect
def stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m
c = stub("simple_knn._C", distCUDA2=lambda *a, **k: None)
stub("simple_knn", _C=c)
stub("plyfile", PlyData=object, PlyElement=object)
import gaussian_renderer                      # reference module, unmodified
import scene.gaussian_model as gm             # reference module, unmodified
pkg = sys.argv[1]
for obj in (gaussian_renderer.GaussianRasterizationSettings, gaussian_renderer.GaussianRasterizer, gaussian_renderer._C,
            gm.load_hierarchy, gm.write_hierarchy):
    src = inspect.getsourcefile(obj)
    assert src.startswith(pkg), (obj, src)
for name in ("render", "render_post", "render_coarse"):
    assert callable(getattr(gaussian_renderer, name))
# the scripts' own imports of the LOD ops (train_post.py:26, render_hierarchy.py:27)
from gaussian_hierarchy._C import expand_to_size, get_interpolation_weights
assert inspect.getsourcefile(expand_to_size).startswith(pkg)
# the settings tuple the three call sites build (all 17 by keyword)
import re
text = open(inspect.getsourcefile(gaussian_renderer)).read()
calls = re.findall(r"GaussianRasterizationSettings\((.*?)\n    \)", text, re.S)
assert len(calls) == 3, len(calls)
for call in calls:
    kws = re.findall(r"^\s*(\w+)\s*=", call, re.M)
    assert set(kws) == set(gaussian_renderer.GaussianRasterizationSettings._fields), kws
print("ok")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "gaussian_renderer")), reason="reference checkout not present")
def test_reference_modules_import_against_our_packages():
    pkg = os.path.join(ROOT, "hierarchical-3d-gaussians_b200")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([pkg, REF]))
    r = subprocess.run([sys.executable, "-c", SCRIPT, pkg], capture_output=True, text=True, env=env, cwd="/tmp")
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
