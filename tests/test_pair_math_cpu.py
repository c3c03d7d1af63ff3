"""CPU: the blend kernels' per-entry arithmetic (csrc/pair_math.cuh, compiled for the host over a
two-float struct instead of the sm_100 packed-FP32 instructions) against a double-precision restatement
of the per-pixel recurrences -- exponent, capped alpha, hierarchy weight and its derivative, termination,
and the back-to-front gradient in its immediate form vs the classic deferred one."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pair_math_matches_double_precision_recurrences(tmp_path):
    exe = str(tmp_path / "pair_math_test")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "emul", "pair_math_test.cpp"),
                    "-lm"], check=True)
    r = subprocess.run([exe, "500"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
