#!/usr/bin/env python
"""Config 5 (BASELINE.json): eval_reg_with_jacobian.py:62-80 on 256^3 displacement fields -- Jacobian-determinant
map (252^3 interior), folding-voxel count, mean / std -- `fields` volumes per GPU per step (weak scaling; fields
are independent, no collective).  Two inputs: a smooth field (folds rare) and raw std-8 noise (many folds).

  python scripts/bench_jacobian.py [--fields 8] [--steps 5]
  python -m torch.distributed.run --nproc-per-node N scripts/bench_jacobian.py
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

S = 256
BYTES = 12 * S ** 3 + 4 * (S - 4) ** 3          # fp32 planar field in, fp32 determinant map out


def measure(rank, world, dev, fields=8, steps=5, warmup=3):
    g = torch.Generator(device='cpu').manual_seed(55 + rank)
    coarse = (torch.randn(fields, 3, 16, 16, 16, generator=g) * 4).to(dev)
    smooth = torch.nn.functional.interpolate(coarse, size=(S, S, S), mode='trilinear').contiguous().permute(0, 2, 3, 4, 1)   # planar storage
    res = {}
    for name, make in (('smooth', lambda: smooth), ('noise_std8', lambda: (torch.randn(fields, 3, S, S, S, device=dev) * 8).permute(0, 2, 3, 4, 1))):
        field = make()
        for _ in range(warmup):
            det, stats = ops.jacobian_determinant(field, out_dtype=torch.float32)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            det, stats = ops.jacobian_determinant(field, out_dtype=torch.float32)
        t1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([t0.elapsed_time(t1) / steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        peak, _ = bench.measured_peak_gbs()
        gbs = world * fields * BYTES / (ms * 1e-3) / 1e9
        res[name] = {'ms_per_step': ms, 'voxels_per_s': world * fields * (S - 4) ** 3 / (ms * 1e-3),
                     'aggregate_GBps': gbs, 'frac_of_peak_per_gpu': gbs / world / peak,
                     'folding_fraction_rank0': float(stats[:, 0].sum().item()) / (fields * (S - 4) ** 3)}
        del field, det
    return {'workload': 'eval_reg_with_jacobian.py:62-80 on %d x 256^3 fp32 fields per GPU: determinant map + fold count + moments' % fields,
            'n_gpus': world, 'fields_per_gpu': fields, 'algorithmic_GB_per_field': BYTES / 1e9, 'scaling': 'weak', **res}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--fields', type=int, default=8)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    args = ap.parse_args()
    rank, world = sharding.env_rank_world()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    out = measure(rank, world, dev, args.fields, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
