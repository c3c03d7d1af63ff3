"""h3dgs: host-side plumbing of the B200-native hierarchical-Gaussian rasterizer.

`_lib`  - ctypes binding of lib/libh3dgs.so (the C-ABI of include/h3dgs.h)
`synth` - synthetic scenes for tests/bench
`dist`  - screen-tile-sharded multi-GPU mode (torch.distributed / NCCL)
"""
