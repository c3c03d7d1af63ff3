"""CPU: the library's tile_sort_gather_kernel (emulation build) driven directly over tile lists of every size class of the
in-register sort and its boundaries (1 .. 8192 entries, many equal depths), against std::sort of the same keys."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))


def test_tile_sort_kernel_orders_every_size_class(tmp_path):
    from build_emu import build
    out = str(tmp_path / "emu")
    build(out)
    src = os.path.join(out, "hierarchical-3d-gaussians_b200", "csrc")
    exe = str(tmp_path / "tile_sort_test")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-w", "-D__CUDACC__", "-DH3_PAIR_HOST_EMU", "-DH3_HOST_EMU", "-I", src,
                    "-o", exe, os.path.join(HERE, "emul", "tile_sort_test.cpp"), "-L", out, "-lh3dgs_emu", f"-Wl,-rpath,{out}", "-lm"],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "all tiles in order" in r.stdout, r.stdout + r.stderr
