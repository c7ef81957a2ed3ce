"""src/gprc_shim.c -- the `.Call` shim of the drop-in R package -- EXECUTED against libgprc on the GPU.  R is not installed
in this image, so a miniature R runtime (tests/stubs/mini_r.c: SEXPs, PROTECT, Rf_error as longjmp, external pointers with
finalizers, R_ToplevelExec, the registered routine table) stands in for it and tests/shim_exec.c plays the R host code:
the four known answers of the reference's tests/testthat/test-gpr.R, the first case of test-gpc.R, shapes of every
return value, the error and interrupt paths, and the finalizers."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gaussian-process-regression_b200")


def build(tmp_path):
    exe = str(tmp_path / "shim_exec")
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-D_GNU_SOURCE", "-I" + os.path.join(ROOT, "tests", "stubs"),
           "-I" + os.path.join(ROOT, "include"), os.path.join(PKG, "src", "gprc_shim.c"),
           os.path.join(ROOT, "tests", "stubs", "mini_r.c"), os.path.join(ROOT, "tests", "shim_exec.c"), "-L" + PKG, "-lgprc",
           "-lm", "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_shim_and_driver_compile_and_link(tmp_path):
    """CPU tier: the shim, the miniature runtime and the driver build warning-free and link against libgprc.so"""
    build(tmp_path)


@pytest.mark.gpu
def test_shim_entry_points_execute_against_the_library(tmp_path):
    exe = build(tmp_path)
    env = dict(os.environ, LD_LIBRARY_PATH=PKG + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "ALL OK" in out.stdout and "FAIL" not in out.stdout, out.stdout + out.stderr
    assert out.stdout.count(" ok\n") >= 16, out.stdout
