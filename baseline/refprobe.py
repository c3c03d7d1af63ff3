"""MEASUREMENT / TEST INFRASTRUCTURE ONLY: looks for a build of the REFERENCE's own CUDA packages of the path --
`diff_gaussian_rasterization` (graphdeco-inria/hierarchy-rasterizer @ 63fa2476, /root/reference/.gitmodules:5-7,
requirements.txt:11) and `gaussian_hierarchy` (gaussian-hierarchy @ 677c8553, .gitmodules:11-13, requirements.txt:13) --
on this machine, WITHOUT importing this repo's drop-in packages of the same name by mistake.

Search order (SURVEY.md 8c / 8d "reference CUDA baseline"):
  1. <repo>/baseline/_ref/                (where a driver- or user-provided `pip install --target` of the pinned commits lands)
  2. every other sys.path / site-packages entry that is not inside this repository
Both submodule directories are empty in /root/reference and there is no network, so on the boxes this project has seen
the probe comes back empty; bench.py prints its result into the bench line and `--impl reference-cuda` /
tests/test_gpu_vs_reference_build.py engage the moment a build shows up.

load() imports the found packages under PRIVATE module names (ref_diff_gaussian_rasterization, ref_gaussian_hierarchy)
so that ours and theirs can live in one process and be fed identical inputs."""
import importlib.machinery
import importlib.util
import os
import site
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
NAMES = ("diff_gaussian_rasterization", "gaussian_hierarchy")


def _candidates():
    seen, out = set(), []
    paths = [REF_DIR] + list(sys.path)
    try:
        paths += site.getsitepackages() + [site.getusersitepackages()]
    except Exception:
        pass
    for p in paths:
        p = os.path.abspath(p or ".")
        if p in seen or not os.path.isdir(p):
            continue
        seen.add(p)
        if p.startswith(ROOT + os.sep) and not p.startswith(REF_DIR):
            continue                                  # this repository's own drop-in packages are not "the reference"
        out.append(p)
    return out


def probe():
    """-> dict(available, packages{name: path or None}, searched[...], note)."""
    found = {}
    searched = _candidates()
    for name in NAMES:
        found[name] = None
        for p in searched:
            spec = importlib.machinery.PathFinder.find_spec(name, [p])
            if spec is None or not spec.origin:
                continue
            pkg_dir = os.path.dirname(spec.origin)
            # the reference packages carry a compiled extension `_C*.so`; ours is a ctypes shim `_C.py`
            has_ext = any(f.startswith("_C") and f.endswith((".so", ".pyd")) for f in os.listdir(pkg_dir))
            if has_ext:
                found[name] = spec.origin
                break
    ok = found["diff_gaussian_rasterization"] is not None
    note = ("reference CUDA build found" if ok else
            "no build of hierarchy-rasterizer / gaussian-hierarchy on this machine (submodules empty in /root/reference, "
            "no network): reference-CUDA baseline UNMEASURED")
    return dict(available=ok, packages=found, searched=len(searched), ref_dir_exists=os.path.isdir(REF_DIR), note=note)


def load(name):
    """Import the found reference package `name` as module `ref_<name>` (None when absent)."""
    origin = probe()["packages"].get(name)
    if origin is None:
        return None
    alias = "ref_" + name
    if alias in sys.modules:
        return sys.modules[alias]
    spec = importlib.util.spec_from_file_location(alias, origin, submodule_search_locations=[os.path.dirname(origin)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    import json
    print(json.dumps(probe(), indent=1))
