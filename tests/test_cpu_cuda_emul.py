"""The device code of the counts-in-spheres kernel (corrfunc_b200/csrc/cuda/spheres_kernel.cuh), compiled for the CPU by
tests/cuda_emul/emul_spheres.cpp -- one std::thread per CUDA thread, real barriers for __syncthreads / __syncwarp -- and
compared with a brute force: float and double, periodic and open, shell and neighbour-count modes, lattices down to
one cell.  Covers the kernel's indexing and arithmetic; what the CUDA runtime does is covered by the -m gpu tests."""
import os
import subprocess

import harness as H


def test_k_spheres_text_runs_correctly_on_the_cpu(tmp_path):
    exe = str(tmp_path / "emul_spheres")
    subprocess.check_call(["/usr/bin/g++", "-std=c++20", "-O1", "-pthread", "-ffp-contract=off",
                           "-I", os.path.join(H.ROOT, "corrfunc_b200", "csrc", "cuda"),
                           os.path.join(H.ROOT, "tests", "cuda_emul", "emul_spheres.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "all cases agree" in out.stdout, out.stdout + out.stderr
    assert out.stdout.count(": ok") == 11
