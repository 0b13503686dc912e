"""Two ranks on two GPUs (skipped on a single-GPU box): the sharded sweep -- clip-affine partition, the per-clip rank
groups that split felzenszwalb / SIFT work and all-reduce it over NCCL, the final gather -- gives the scores of the
single-process sweep."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import PKG

pytestmark = pytest.mark.gpu


def test_two_rank_sweep_equals_single_rank(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import pandas as pd
    outs = {}
    for world in (1, 2):
        out = tmp_path / ("results_%d.csv" % world)
        base = [os.path.join(PKG, "benchmark.py"), "--videos", "soldier,bmx", "--limit", "48", "--max-frames", "14", "--out", str(out)]
        cmd = [sys.executable] + base if world == 1 else \
            [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
             "--master-port", "29547"] + base
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=PKG)
        assert res.returncode == 0, res.stderr[-3000:]
        outs[world] = pd.read_csv(out)
    a, b = outs[1], outs[2]
    assert list(a.columns) == list(b.columns) and len(a) == len(b)
    for col in a.columns:
        if col.endswith("_benchmark"):
            assert np.array_equal(a[col].to_numpy(), b[col].to_numpy(), equal_nan=True), col
