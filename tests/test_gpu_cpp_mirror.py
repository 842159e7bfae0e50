"""Runs the C++ host-mirror test program (tests/cpp/test_mirror.cpp → tests/cpp/test_mirror),
i.e. the reference's unit tests restated against include/ronk_b200.hpp, on the GPU."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_unit_tests():
    exe = os.path.join(HERE, "cpp", "test_mirror")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(HERE, "..", "ronkathon_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "cpp mirror ok" in out.stdout, out.stdout + out.stderr
