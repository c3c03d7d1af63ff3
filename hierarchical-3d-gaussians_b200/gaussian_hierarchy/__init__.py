"""Drop-in replacement for the `gaussian_hierarchy` package of
graphdeco-inria/gaussian-hierarchy (absent from /root/reference): `_C.expand_to_size`,
`_C.get_interpolation_weights` (train_post.py:26,91-113; render_hierarchy.py:27,58-80)."""
from . import _C  # noqa: F401
