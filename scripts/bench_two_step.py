#!/usr/bin/env python
"""Config 4 (BASELINE.json): bids_two_steps_registration.py cascaded two-step warp composition over a
synthetic 64-subject BIDS set, subject-sharded across ranks (one process per GPU, no collective).

Per subject (bids_two_steps_registration.py:318-355,504-516, whole-volume branch): the two models'
second outputs w1, w2 (half-resolution fields) are composed with vxm.utils.compose([w1, w2]); the
composed field is applied to the moving image through Transform(rescale=2) -- linear for the image,
nearest for a segmentation -- and rescaled x2 once more for the saved SCT warp.

  python scripts/bench_two_step.py [--subjects 64] [--steps 5]
  python -m torch.distributed.run --nproc-per-node 8 scripts/bench_two_step.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench
import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops, sharding

FULL, HALF = bench.FULL, bench.HALF
N_F, N_H = bench.N_F, bench.N_H
# algorithmic bytes per subject: compose (36 B/voxel half-res) + rescale x2 + linear warp + nearest warp
# + rescale x2 for the saved warp
BYTES = 36 * N_H + (12 * N_H + 12 * N_F) + 20 * N_F + 20 * N_F + (12 * N_H + 12 * N_F)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--subjects', type=int, default=64)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    args = ap.parse_args()
    rank, world = sharding.env_rank_world()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    mine = sharding.shard_items(args.subjects, rank, world)
    B = len(mine)
    g = torch.Generator(device='cpu').manual_seed(7 + rank)

    def field(std):
        c = torch.randn(B, 3, 10, 10, 12, generator=g) * std
        return torch.nn.functional.interpolate(c, size=HALF, mode='trilinear', align_corners=True).permute(0, 2, 3, 4, 1).contiguous().to(dev)

    w1, w2 = field(3.0), field(1.0)
    moving = torch.rand(B, *FULL, 1, generator=g).to(dev)
    seg = torch.randint(0, 26, (B, *FULL, 1), generator=g).float().to(dev)

    def step():
        with torch.no_grad():
            warp = ops.compose([w1, w2])                              # :324
            full = ops.rescale_dense_transform(warp, 2)               # Transform(rescale=2) / :515
            moved = ops.warp(moving, full)                            # linear image warp
            moved_seg = ops.warp(seg, full, 'nearest')                # nearest variant (:338-355)
            saved = ops.rescale_dense_transform(warp, 2, out_layout='cl')
        return moved, moved_seg, saved

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, _ = bench.measured_peak_gbs()
        ms = float(ms.item())
        gbs = args.subjects * BYTES / (ms * 1e-3) / 1e9
        print(json.dumps({'workload': 'bids_two_steps_registration.py: compose + x2 rescale + linear and nearest warp per subject',
                          'n_gpus': world, 'subjects': args.subjects, 'subjects_per_gpu': B, 'ms_per_pass': ms,
                          'subjects_per_s': args.subjects / (ms * 1e-3), 'algorithmic_GB_per_subject': BYTES / 1e9,
                          'aggregate_GBps': gbs, 'frac_of_peak_per_gpu': gbs / world / peak, 'scaling': 'strong'}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
