"""The GEMM-shaped steps have two implementations: tcgen05 (3xTF32, default) and the FP32-pipe kernels
(GNNFP_TC=0 / GNNFP_TC_DW=0, also the path for shapes the tensor-core kernels do not take).  The rest of the GPU suite
runs the default; this test re-runs the forward / backward / golden parity files with the tensor-core kernels
switched off (the switches are read once per process, hence the subprocess)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_fp32_pipe_kernels_pass_the_same_parity_tests():
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, GNNFP_TC="0", GNNFP_TC_DW="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        os.path.join(here, "test_gpu_forward.py"), os.path.join(here, "test_gpu_backward.py"),
                        os.path.join(here, "test_gpu_golden.py")],
                       env=env, cwd=os.path.dirname(here), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
