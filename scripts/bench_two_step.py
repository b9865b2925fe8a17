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
INT_STEPS = 7
# algorithmic bytes per subject: two VxmDense tails (7 SS steps + x2 rescale + linear warp each),
# compose at half resolution, Transform(nearest, rescale=2) of a segmentation with the composed warp, x2 rescale
# of the saved SCT warp
TAIL = INT_STEPS * 24 * N_H + (12 * N_H + 12 * N_F) + 20 * N_F
BYTES = 2 * TAIL + 36 * N_H + ((12 * N_H + 12 * N_F) + 20 * N_F) + (12 * N_H + 12 * N_F)


def measure(rank, world, dev, subjects=64, steps=5, warmup=3):
    """One process per GPU; the process group (if world > 1) is already initialised.  Subjects are sharded
    across ranks (strong scaling: the 64-subject set is fixed); times are max over ranks.  Everything runs through
    pipelines.two_steps_tail -- the device-resident restatement of bids_two_steps_registration.py:316-355,504-516."""
    from multimodal_registration_b200 import pipelines
    from multimodal_registration_b200.voxelmorph import networks
    mine = sharding.shard_items(subjects, rank, world)
    B = len(mine)
    g = torch.Generator(device='cpu').manual_seed(7 + rank)

    def field(std):
        c = torch.randn(B, 3, 10, 10, 12, generator=g) * std
        return torch.nn.functional.interpolate(c, size=HALF, mode='trilinear', align_corners=True).permute(0, 2, 3, 4, 1).contiguous().to(dev)

    flow1, flow2 = field(3.0), field(1.0)                             # what the two U-Nets' flow convolutions emit
    moving = torch.rand(B, *FULL, 1, generator=g).to(dev)
    fixed = torch.rand(B, *FULL, 1, generator=g).to(dev)
    seg = torch.randint(0, 26, (B, *FULL, 1), generator=g).float().to(dev)
    model1 = networks.VxmDense(FULL, int_steps=INT_STEPS, svf_resolution=2, int_resolution=2)
    model2 = networks.VxmDense(FULL, int_steps=INT_STEPS, svf_resolution=2, int_resolution=2)

    def step():
        with torch.no_grad():
            res = pipelines.two_steps_tail(moving, fixed, model1, model2, flow1, flow2, 'linear')    # :316-324
            saved = ops.rescale_dense_transform(res['warp'], res['scale'], out_layout='cl')          # :515
            # nearest variant of the final transform (:338-355, Transform(nearest, rescale=scale)): the rescaled warp is
            # needed for the export anyway, so the label warp reads it instead of re-deriving it inside a fused kernel
            moved_seg = ops.warp(seg, saved, 'nearest')
        return res['moved'], moved_seg, saved

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    peak, _ = bench.measured_peak_gbs()
    ms = float(ms.item())
    gbs = subjects * BYTES / (ms * 1e-3) / 1e9
    return {'workload': 'bids_two_steps_registration.py via pipelines.two_steps_tail: two VxmDense tails (7 SS steps, x2, linear warp), compose, nearest seg transform, x2 rescale of the saved warp, per subject',
            'n_gpus': world, 'subjects': subjects, 'subjects_per_gpu': B, 'ms_per_pass': ms,
            'subjects_per_s': subjects / (ms * 1e-3), 'algorithmic_GB_per_subject': BYTES / 1e9,
            'aggregate_GBps': gbs, 'frac_of_peak_per_gpu': gbs / world / peak, 'scaling': 'strong'}


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
    res = measure(rank, world, dev, args.subjects, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
