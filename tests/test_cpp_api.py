"""The C++17 host mirror (include/cppoptlib_b200/): compiles without a GPU
(g++ translation unit + an nvcc user-functor translation unit); runs on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "build")


def test_cpp_programs_compile():
    from cppnumericalsolvers_b200 import build
    exes = build.build_cpp_tests()
    assert all(os.path.exists(e) for e in exes)
    # the host program is a plain g++ translation unit that needs no CUDA compiler
    syms = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(BUILD, "verify_host")],
                          capture_output=True, text=True).stdout
    assert "cno_minimize" in syms and "cno_default_stop" in syms


@pytest.mark.gpu
@pytest.mark.parametrize("exe", ["verify_host", "user_functor", "sharded_nccl"])
def test_cpp_programs_pass_on_gpu(exe):
    path = os.path.join(BUILD, exe)
    if not os.path.exists(path):
        from cppnumericalsolvers_b200 import build
        build.build_cpp_tests()
    r = subprocess.run([path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
