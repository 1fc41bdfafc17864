"""The compiled C++ programs on top of libgrbda_cuda.so: tests/cpp/test_batched_model.cpp (batched methods of
the host ClusterTreeModel against the oracle) and benchmarks/batchedBenchmark.cpp. CPU: they compile and link
(every symbol resolves); GPU: they run."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = "/usr/local/cuda"


def compile_cpp(grbda, oracle, source, exe, with_oracle):
    libdir = os.path.dirname(grbda.library_path())
    cmd = ["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "generalized_rbda_b200", "csrc"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"), "-o", exe, source,
           "-L", libdir, "-lgrbda_cuda", "-L", os.path.join(CUDA, "lib64"), "-lcudart",
           "-Wl,-rpath," + libdir]
    if with_oracle:
        odir = os.path.join(ROOT, "oracle")
        oracle.lib()
        cmd += ["-L", odir, "-loracle", "-Wl,-rpath," + odir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_programs_compile_and_link(grbda, oracle, tmp_path):
    compile_cpp(grbda, oracle, os.path.join(ROOT, "tests", "cpp", "test_batched_model.cpp"), str(tmp_path / "t"), True)
    compile_cpp(grbda, oracle, os.path.join(ROOT, "benchmarks", "batchedBenchmark.cpp"), str(tmp_path / "b"), False)


def test_casadi_phi_bridge_with_a_stand_in_function(grbda, oracle, tmp_path):
    """include/grbda_cuda_casadi.hpp (the reference-side casadi::Function -> grbda_phi_op walker of
    INTEGRATION.md) driven by a stand-in with casadi::Function's instruction interface."""
    exe = compile_cpp(grbda, oracle, os.path.join(ROOT, "tests", "cpp", "test_phi_bridge.cpp"), str(tmp_path / "p"), False)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_batched_model_against_oracle(grbda, oracle, tmp_path):
    exe = compile_cpp(grbda, oracle, os.path.join(ROOT, "tests", "cpp", "test_batched_model.cpp"), str(tmp_path / "t"), True)
    r = subprocess.run([exe, grbda.URDF_DIR], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_batched_benchmark_runs(grbda, oracle, tmp_path):
    exe = compile_cpp(grbda, oracle, os.path.join(ROOT, "benchmarks", "batchedBenchmark.cpp"), str(tmp_path / "b"), False)
    urdf = os.path.join(ROOT, "tests", "urdf_corpus", "revolute_rotor_branch_2_3.urdf")
    r = subprocess.run([exe, "--batch", "65536", "--steps", "3", "--urdf-dir", grbda.URDF_DIR, "mini_cheetah", urdf],
                       capture_output=True, text=True, timeout=600)
    lines = r.stdout.strip().splitlines()
    assert r.returncode == 0 and len(lines) == 3, r.stdout + r.stderr
    assert lines[0].startswith("model,bodies,clusters")
    for ln in lines[1:]:
        f = ln.split(",")
        assert float(f[7]) > 0 and float(f[8]) > 0 and float(f[11]) > 1e6
