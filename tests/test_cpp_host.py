"""The C++ host side (include/pop.hpp) compiles against the C ABI and its host-only parts behave (defaults, wire
format, no CPU fallback) — built with g++ and run without a GPU.  examples/cpp_host.cpp runs the GPU part when
POP_CPP_GPU=1."""
import os
import shutil
import subprocess

import pytest

from tests.conftest import ROOT


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_cpp_host_compiles_and_runs(pop, tmp_path):
    pop.load()   # libpop.so must exist
    exe = str(tmp_path / "cpp_host")
    pkg = os.path.join(ROOT, "pred_occ_planner_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-pthread", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "cpp_host.cpp"), "-o", exe, os.path.join(pkg, "libpop.so"), f"-Wl,-rpath,{pkg}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    env = dict(os.environ)
    env.pop("POP_CPP_GPU", None)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host-only checks ok" in r.stdout


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_cpp_host_on_gpu(pop, tmp_path):
    """The same program with its GPU part: BASELINE config 1 through pop::Engine / RiskVoxelMap / HybridAstar."""
    exe = str(tmp_path / "cpp_host")
    pkg = os.path.join(ROOT, "pred_occ_planner_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cpp_host.cpp"),
           "-o", exe, os.path.join(pkg, "libpop.so"), f"-Wl,-rpath,{pkg}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=dict(os.environ, POP_CPP_GPU="1"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GPU checks ok" in r.stdout
    # two GPUs: the same program flies a four-agent swarm as two ranks (a host thread and an engine per GPU) and
    # exchanges the trajectories through pop_comm_init / pop_exchange_trajs
    import torch
    if torch.cuda.device_count() >= 2:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, POP_CPP_GPU="2"))
        assert r.returncode == 0, r.stdout + r.stderr
        assert "two-rank exchange ok" in r.stdout
