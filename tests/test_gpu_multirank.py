"""Multi-GPU parity under NCCL (needs >= 2 visible GPUs; skipped otherwise): the CUDA fold and the
NCCL all-reduce of profiles and counts TOGETHER against the single-process oracle, the cfg4
pipeline in small, and channel-sharded dedispersion with the global ref_freq / crop.  The CPU
suite covers the same host logic with gloo (tests/test_sharding.py)."""

import glob
import json
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_cuda_fold_and_nccl_allreduce_match_single_rank_oracle(world, tmp_path):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    out = str(tmp_path / "mr")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "multirank_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(HERE))
    reports = [json.load(open(p)) for p in sorted(glob.glob(out + ".rank*.json"))]
    assert r.returncode == 0, (r.stderr[-3000:], reports)
    assert len(reports) == world and all(rep["ok"] and rep["backend"] == "nccl" for rep in reports)
    for rep in reports:
        for name, chk in rep["checks"].items():
            assert chk.get("counts_equal", True) and chk.get("bins_equal", True), (name, chk)
    log = os.environ.get("PBK_MULTIRANK_LOG")
    if log:                      # keep the evidence (profiles/ in a gpurun call)
        with open(log, "a") as f:
            f.write(json.dumps({"world": world, "rank0": reports[0]["checks"]}) + "\n")
