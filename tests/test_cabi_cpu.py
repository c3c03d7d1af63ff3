"""CPU: the C-ABI library loads without a GPU and exports every symbol include/h3dgs.h declares;
the ctypes mirror of h3dgs_raster_args matches the header field for field; the Python packages
expose the reference's names.  No compute calls (there is no GPU here)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from h3dgs import _lib
    l = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "h3dgs.h")).read()
    declared = set(re.findall(r"\b(h3dgs_[a-z0-9_]+)\s*\(", hdr)) - {"h3dgs_alloc_fn"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(l, name), name
    assert set(_lib.EXPORTS) <= declared
    assert l.h3dgs_version() == 1
    assert l.h3dgs_backward_scratch_bytes(1000) >= 1000 * 10 * 4
    assert l.h3dgs_expand_scratch_bytes(1000) > 0
    assert l.h3dgs_stage_name(6) == b"render_forward" and l.h3dgs_stage_name(3) == b"sort"


def test_ctypes_struct_mirrors_header():
    from h3dgs import _lib
    hdr = open(os.path.join(ROOT, "include", "h3dgs.h")).read()
    body = hdr[hdr.index("typedef struct h3dgs_raster_args {"):hdr.index("} h3dgs_raster_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    body = body[body.index("{") + 1:]
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            m = re.search(r"(\w+)\s*(?:\[\w+\])?\s*$", part.strip())
            if m:
                names.append(m.group(1))
    assert names == [f[0] for f in _lib.RasterArgs._fields_]


def test_drop_in_package_surface():
    import inspect
    import diff_gaussian_rasterization as d
    import gaussian_hierarchy._C as g
    assert d.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug", "render_indices", "parent_indices", "interpolation_weights",
        "num_node_kids", "do_depth")
    for name in ("GaussianRasterizer", "rasterize_gaussians", "_RasterizeGaussians", "_C"):
        assert hasattr(d, name)
    sig = inspect.signature(d.GaussianRasterizer.forward)
    assert list(sig.parameters)[1:] == ["means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                        "rotations", "cov3D_precomp"]
    assert list(inspect.signature(g.expand_to_size).parameters) == [
        "nodes", "boxes", "size", "viewpoint", "viewdir", "render_indices", "parent_indices", "nodes_for_render_indices"]
    assert list(inspect.signature(g.get_interpolation_weights).parameters) == [
        "node_indices", "size", "nodes", "boxes", "viewpoint", "viewdir", "interpolation_weights", "num_siblings"]


def test_no_product_import_of_the_oracle():
    """the product path must never route through oracle/"""
    pkg = os.path.join(ROOT, "hierarchical-3d-gaussians_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, os.path.join(dp, f)
