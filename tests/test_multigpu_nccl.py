"""N > 1 on real GPUs: the cross-GPU merges of the C ABI (NCCL) and the multi-device operator layer, checked
against the CPU oracle over the union of the partitions (tests/mgpu_worker.py).  Needs >= 2 GPUs in the box
(`gpurun --gpus 2`); skipped otherwise."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("peer_merge", ["1", "0"], ids=["peer_memory", "nccl_only"])
def test_nccl_merges_match_oracle(world, peer_merge):
    """peer_memory: the fixed-size aggregation states are merged by k_merge_peer_compact through the CUDA-IPC mailbox
    (comm.cu); nccl_only (QSGPU_PEER_MERGE=0): the same merge as ncclAllGather + fold.  Both must give the oracle's answers."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + (os.getpid() % 300) + world + (10 if peer_merge == "0" else 0)
    env = dict(os.environ, QSGPU_PEER_MERGE=peer_merge)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")],
                       capture_output=True, text=True, timeout=850, env=env)
    if r.returncode != 0 or "MGPU OK" not in r.stdout:
        # the interesting part of eight interleaved tracebacks does not survive pytest's repr: keep all of it
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        log = os.path.join(ROOT, "gpurun_out", f"mgpu_worker_w{world}_peer{peer_merge}.log")
        with open(log, "w") as f:
            f.write(r.stdout + "\n---- stderr ----\n" + r.stderr)
        errs = [ln for ln in r.stderr.splitlines() if "Error" in ln or "assert" in ln.lower()]
        pytest.fail(f"worker failed (full output in {log}): " + " | ".join(errs[:6])[:1500])
    assert ("peer mailbox: on" in r.stdout) == (peer_merge == "1"), r.stdout[-2000:]
