#!/usr/bin/env python
"""A/B of the gradient all-reduce schedule inside ONE torchrun job (same boxes, same clocks): full-model training step with
  (a) bucketed all-reduces issued during backward (parallel.BUCKET_BYTES = 64 MB, round 2)
  (b) one all-reduce of the whole flat buffer after backward (BUCKET_BYTES = inf, round 1 behaviour)
  (c) no all-reduce at all (the compute-only step)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/train_allreduce_ab.py
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from slotdiffusion_b200 import parallel  # noqa: E402


def main():
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    args = types.SimpleNamespace(train_batch=64, train_steps=6, warmup=3, no_train_graph=False, video_clips=8)
    out = {}
    for name, bucket, enable in (('bucketed_64MB', 64 << 20, True), ('single_allreduce_after_backward', 1 << 60, True),
                                 ('no_allreduce', 64 << 20, False), ('bucketed_64MB_again', 64 << 20, True)):
        parallel.BUCKET_BYTES = bucket
        w = world if enable else 1          # world = 1 makes train_bench skip enable_grad_allreduce()
        r = bench.train_bench(args, dev, w, rank, full=True)
        t = torch.tensor([r['ms_per_step']], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = round(t.item(), 3)
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({'world': world, 'ms_per_step': out}))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
