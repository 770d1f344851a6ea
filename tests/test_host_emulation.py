"""Builds and runs tests/host_emul/fft_emul.cu: the cooperative FFT passes and the fused z-kernel data flow of
csrc/fft_core.cuh / fft_kernels.cuh executed on the CPU (nvcc host compile, every 'thread' in sequence) against
a long double DFT.  Guards butterflies, twiddle indices, shared-memory index maps and the packing."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_fft_passes_emulated_on_cpu():
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "fft_emul")
        subprocess.run(["nvcc", "-std=c++17", "-O1", "--extended-lambda", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-diag-suppress", "20013,20015", "-o", exe, os.path.join(ROOT, "tests", "host_emul", "fft_emul.cu")], check=True)
        p = subprocess.run([exe], capture_output=True, text=True)
        assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout
