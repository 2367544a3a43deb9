"""N >= 2 GPUs of one box (skipped on a single GPU): sharded Schur-complement assembly, block-cyclic
Cholesky with NCCL panel broadcasts and the driver on every rank, through scripts/multi_gpu_check.py
under torchrun.  Checks (inside the script): sharded H == single-rank H to 1e-12, distributed factor
== single-rank factor to 1e-12, solve vs SciPy 1e-8, all ranks bitwise identical, identical driver
decisions on all ranks."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("args,block,bignj", [(["600", "300", "5"], 128, None), (["700", "600", "4"], 256, None),
                                              (["rand", "300", "700"], 128, None),
                                              # dense top set forced on every supernode with >= 12 rows: completion, chol(Y_aa) and
                                              # the local phase of the inverse Hessian are shared out over the ranks (NCCL broadcasts)
                                              (["rand", "300", "700"], 128, "12")])
def test_sharded_assembly_and_block_cyclic_factor(args, block, bignj):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    nproc = 4 if n >= 4 else 2
    env = dict(os.environ, SMCP_BLOCK=str(block), MASTER_ADDR="127.0.0.1")
    if bignj:
        env["SMCP_B200_BIG_NJ"] = bignj
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "scripts", "multi_gpu_check.py")] + args,
                         capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_CHECK OK" in out.stdout
