#!/usr/bin/env python
"""IoU scoring and the hyper-parameter sweep of the reference's benchmark.py.

`computeBenchmark(mask, truth)` keeps the reference signature (:8-14) and runs on the GPU
(`pcm_iou`).  Run as a script it executes the VIDEOS x HYPERPARAMS grid (:41-51), one process
per GPU:

    python benchmark.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
           --master-port 29511 benchmark.py               # 8 GPUs, one final score gather

and writes benchmark_results.csv (columns as in the reference, :86-89).
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

_handle = None


def computeBenchmark(mask, truth):
    """Intersection / union of the non-zero pixels of `mask` and `truth` (float64; NaN when the
    union is empty, like the reference's numpy quotient)."""
    global _handle
    from pcm import capi
    if _handle is None:
        _handle = capi.Handle(int(os.environ.get("LOCAL_RANK", "0")))
    inter, union = _handle.iou_counts(np.ascontiguousarray(mask) if mask.strides[-1] != 1 else mask, truth)
    return np.float64(inter) / np.float64(union) if union else np.float64("nan")


def main():
    import yaml
    from pcm import sweep
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", default=",".join(sweep.VIDEOS))
    ap.add_argument("--limit", type=int, default=0, help="only the first N sequences of the grid")
    ap.add_argument("--max-frames", type=int, default=0, help="truncate every clip (smoke runs)")
    ap.add_argument("--out", default="benchmark_results.csv")
    ap.add_argument("--train-jobs", type=int, default=0)
    args = ap.parse_args()
    with open(os.path.join(HERE, "config_benchmark.yaml")) as f:
        base = yaml.full_load(f)
    with open(os.path.join(HERE, "polygons.yaml")) as f:
        polygons = yaml.full_load(f)
    rank = int(os.environ.get("RANK", "0"))
    real_stdout, sys.stdout = sys.stdout, sys.stderr
    try:
        summary, table = sweep.run(base, polygons, videos=args.videos.split(","), limit=args.limit or None,
                                   max_frames=args.max_frames or None, out_csv=args.out if rank == 0 else None,
                                   train_jobs=args.train_jobs or None,
                                   log=lambda *a: print(*a, file=sys.stderr, flush=True))
    finally:
        sys.stdout = real_stdout
    if rank == 0:
        print(json.dumps({"metric": "grid-sweep sequences/sec", "value": summary["sequences_per_s"], "unit": "sequences/s",
                          **summary}), flush=True)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
