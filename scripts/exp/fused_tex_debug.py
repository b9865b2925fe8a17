"""Stress: the fused texture-gather kernel against the two stand-alone kernels, many launches, both builds."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops, _lib
B = 32
svf, img = bench.synth_inputs(B, 'cpu', 0)
svf, img = svf.cuda(), img.cuda()
for exact in (False, True):
    _lib.use(exact)
    half = ops.vecint(svf, 7)
    ref = ops.warp(img, ops.rescale_dense_transform(half, 2))
    bad = 0
    for it in range(40):
        out = ops.rescale_warp(img, half, 2)
        n = int((out != ref).sum())
        bad += n
        if n:
            print('exact', exact, 'it', it, 'n diff', n)
    print('exact=%s: %d differing voxels in 40 launches' % (exact, bad))
