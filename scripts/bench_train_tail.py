#!/usr/bin/env python
"""Config 3 (BASELINE.json): the deformation hot path of one train_synthmorph.py step with
config/config.json shapes, forward + backward, batch-sharded across ranks.

Per item: two labels_to_image generators (VecInt(5) -> x2 rescale -> nearest
warp of a label map, fill_value 0), the VxmDense tail (x0.5 rescale of the full-resolution flow ->
VecInt(5) -> x2 rescale -> linear warp of the source image, C=1), and
`pred = SpatialTransformer('linear')([one_hot(26), flow])`; backward: d pred / d flow (C=26), the x2
rescale adjoint, 5 SS-step adjoints, the x0.5 rescale adjoint.  The U-Net, the losses and the optimizer
are outside the hot path; the upstream gradient of `pred` is a random tensor, the gradient that reaches
the flow convolution is all-reduced across ranks together with a 1 446 979-float stand-in for the
U-Net gradient bucket (the reference's only collective, train_synthmorph.py:284-285).

  python scripts/bench_train_tail.py [--items 2] [--steps 5]
  python -m torch.distributed.run --nproc-per-node N scripts/bench_train_tail.py
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

CFG = json.load(open(os.path.join(ROOT, 'config', 'config.json')))
FULL = tuple(CFG['in_shape'])
HALF = tuple(d // CFG['int_res'] for d in FULL)
C = CFG['num_labels']
STEPS = CFG['int_steps']
N_F = FULL[0] * FULL[1] * FULL[2]
N_H = HALF[0] * HALF[1] * HALF[2]
# algorithmic bytes per item, SURVEY.md section 8(d)
FWD = 2 * (STEPS * 24 * N_H + 12 * N_H + 12 * N_F + 20 * N_F) + (12 * N_F + 12 * N_H) + STEPS * 24 * N_H + \
    (12 * N_H + 12 * N_F) + 20 * N_F + (8 * C + 12) * N_F
BWD = (8 * C + 24) * N_F + (12 * N_F + 12 * N_H) + STEPS * 36 * N_H + (12 * N_F + 12 * N_H)


def measure(rank, world, dev, items=2, steps=5, warmup=3):
    """One process per GPU; the process group (if world > 1) is already initialised.  Returns the result dict
    (identical on every rank: times are max over ranks)."""
    B = items
    g = torch.Generator(device='cpu').manual_seed(100 + rank)

    def smooth(shape_c, std):
        c = torch.randn(B, 3, *shape_c, generator=g) * std
        return torch.nn.functional.interpolate(c, size=HALF, mode='trilinear', align_corners=True).permute(0, 2, 3, 4, 1).contiguous().to(dev)

    vel = [smooth((10, 10, 12), CFG['vel_std'] / 2), smooth((10, 10, 12), CFG['vel_std'] / 2)]   # U(0, vel_std) mean
    labels = [torch.randint(0, C, (B, *FULL, 1), generator=g).float().to(dev) for _ in range(2)]
    flow_full = torch.nn.functional.interpolate(torch.randn(B, 3, 20, 20, 24, generator=g) * 1.5, size=FULL, mode='trilinear',
                                                align_corners=True).permute(0, 2, 3, 4, 1).contiguous().to(dev)
    # the flow is what a torch flow convolution emits: NCDHW storage, i.e. this package's "planar" layout of the logical
    # channels-last tensor (the gradient that comes back in the same layout is taken as is, no layout copy)
    if not os.environ.get('DFM_TRAIN_FLOW_CL'):
        flow_full = ops.to_layout(flow_full, 'planar')
    image = torch.rand(B, *FULL, 1, generator=g).to(dev)
    vel_cat, labels_cat = torch.cat(vel, 0), torch.cat(labels, 0)      # the two generators' inputs, drawn into one batch
    unet_grad = torch.zeros(1446979, device=dev)

    # harness-only tensors, made once: the one-hot encoding belongs to labels_to_image's intensity model
    # (outside the hot path) and the upstream gradient of `pred` comes from the Dice loss
    onehot_cl = torch.nn.functional.one_hot(labels[0][..., 0].long(), C).float()   # channels-last, like the reference
    gpred = torch.rand(B, *FULL, C, generator=g).to(dev)                             # channels-last too
    labels_u8 = None if os.environ.get('DFM_TRAIN_GENERIC_PRED') else labels[0][..., 0].to(torch.uint8)   # what the generator attaches to its map

    def step():
        # generators (no gradient)
        with torch.no_grad():
            # both generators in one batch of 2 B items: VecInt -> RescaleTransform(2) -> nearest warp; the full-resolution
            # field is an intermediate (fused kernel)
            m = ops.rescale_warp(labels_cat, ops.vecint(vel_cat, STEPS), 2, 0, 'nearest')
            maps = [m[:B], m[B:]]
        flow = flow_full.detach().requires_grad_(True)            # what the flow convolution emits (full res)
        svf = ops.rescale_dense_transform(flow, 0.5)
        pos = ops.rescale_dense_transform(ops.vecint(svf, STEPS), 2)
        with torch.no_grad():
            y_source = ops.warp(image, pos.detach())
        # pred = SpatialTransformer('linear')([map_1, flow]): map_1 is the generator's one-hot map, warped from its label map
        # (ops.warp_onehot: the same bits as the generic channels-last warp of the one-hot tensor, 8 bytes gathered per voxel)
        pred = ops.warp_onehot(labels_u8, pos, C) if labels_u8 is not None else ops.warp(onehot_cl, pos)
        # the gradient that reaches the flow convolution (autograd.grad: handed on as produced, as it would be to the conv's
        # backward -- .backward() on a leaf adds AccumulateGrad's layout copy, which is not part of the path)
        gflow, = torch.autograd.grad(pred, flow, gpred)
        sharding.allreduce_mean_(unet_grad)                       # the step's only collective
        return gflow, y_source

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
    peak, src = bench.measured_peak_gbs()
    ms = float(ms.item())
    gbs = world * B * (FWD + BWD) / (ms * 1e-3) / 1e9
    return {'workload': 'train_synthmorph.py step, deformation hot path fwd+bwd (config/config.json shapes)',
            'n_gpus': world, 'items_per_gpu': B, 'ms_per_step': ms, 'items_per_s': world * B / (ms * 1e-3),
            'algorithmic_GB_per_item': (FWD + BWD) / 1e9, 'aggregate_GBps': gbs, 'frac_of_peak_per_gpu': gbs / world / peak,
            'includes': '26-channel pred and its gradient in the reference channels-last layout, computed from the one-hot map\'s label map '
                        '(dfm_warp_onehot_fwd / _bwd; DFM_TRAIN_GENERIC_PRED=1: generic channels-last kernels), 5.79 MB NCCL all-reduce',
            'scaling': 'weak'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--items', type=int, default=2, help='items per GPU per step')
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    args = ap.parse_args()
    rank, world = sharding.env_rank_world()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    res = measure(rank, world, dev, args.items, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
