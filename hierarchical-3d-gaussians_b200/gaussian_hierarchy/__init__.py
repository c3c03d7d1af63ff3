"""Drop-in replacement for the `gaussian_hierarchy` package of
graphdeco-inria/gaussian-hierarchy (absent from /root/reference): `_C.expand_to_size`,
`_C.get_interpolation_weights` (train_post.py:26,91-113; render_hierarchy.py:27,58-80) and the `.hier`
reader/writer `_C.load_hierarchy` / `_C.write_hierarchy` (scene/gaussian_model.py:24; layout unpinned, hier_io.py)."""
from . import _C  # noqa: F401
