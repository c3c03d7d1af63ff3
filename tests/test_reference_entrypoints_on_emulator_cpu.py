"""CPU, where the reference checkout is present: tests/test_gpu_reference_entrypoints.py -- the reference's unmodified
render() / render_post() on top of the drop-in packages -- and tests/test_gpu_reference_train_post.py -- the loop body of
the reference's train_post.py (config #4 in miniature) -- executed against the emulation build of the kernels
(H3DGS_EMULATE=1, tests/conftest.py).  On a GPU box that has no /root/reference this is the run that counts."""
import os
import re
import subprocess
import sys

import pytest

import refharness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not refharness.have_reference(), reason="reference checkout not present")
def test_reference_render_and_render_post_run_unmodified_on_the_emulator():
    env = dict(os.environ, H3DGS_EMULATE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_reference_entrypoints.py", "tests/test_gpu_reference_train_post.py", "-q", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = r.stdout[-1500:]
    assert r.returncode == 0, tail + r.stderr[-1500:]
    m = re.search(r"(\d+) passed", tail)
    assert m and int(m.group(1)) == 3 and "skipped" not in tail, tail
