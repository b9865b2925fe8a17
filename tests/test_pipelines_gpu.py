"""Config 4: the device-resident tail of bids_two_steps_registration.py:register against the chained oracle
(each step restated with oracle/interp_oracle.py in the order the reference script runs them)."""
import numpy as np
import pytest
import torch

import multimodal_registration_b200 as mrb
from multimodal_registration_b200 import ops, pipelines
from oracle import interp_oracle as io
from oracle import sct_oracle, stitch_oracle

pytestmark = pytest.mark.gpu
vxm = mrb.voxelmorph
RTOL, ATOL = 1e-5, 1e-4


@pytest.fixture(autouse=True, params=['fast', 'exact'])
def arithmetic_mode(request):
    mrb._lib.use(request.param == 'exact')
    yield request.param
    mrb._lib.use(False)


def smooth(rng, shape, std):
    c = rng.standard_normal(shape).astype(np.float32)
    for ax in (1, 2, 3):
        c = (c + np.roll(c, 1, ax) + np.roll(c, -1, ax)) / 3
    return (c / max(c.std(), 1e-6) * std).astype(np.float32)


def check(got, want):
    got = ops.to_layout(got, 'cl').cpu().numpy() if isinstance(got, torch.Tensor) else got
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
    if mrb._lib.exact_order():
        np.testing.assert_array_equal(got, want)


def oracle_tail(source, flow, int_steps):
    """VxmDense tail with svf_res = int_res = 2 on a half-resolution flow."""
    pos = io.rescale_dense_transform(io.vec_int(flow, int_steps), 2)
    return io.spatial_transformer(source, pos), flow          # [y_source, preint_flow]


@pytest.mark.parametrize('interp', ['linear', 'nearest'])
def test_two_steps_whole_volume(interp):
    rng = np.random.default_rng(3)
    full, half = (16, 24, 40), (8, 12, 20)
    moving = rng.random((1,) + full + (1,)).astype(np.float32)
    if interp == 'nearest':
        moving = np.floor(moving * 5).astype(np.float32)
    fixed = rng.random((1,) + full + (1,)).astype(np.float32)
    flow1 = smooth(rng, (1,) + half + (3,), 1.5)
    f2 = smooth(rng, (1,) + half + (3,), 0.7)
    model1 = vxm.networks.VxmDense(full, int_steps=5, svf_resolution=2, int_resolution=2)
    model2 = vxm.networks.VxmDense(full, int_steps=5, svf_resolution=2, int_resolution=2)
    seen = {}

    def flow2(moved_first, fx):                              # the second U-Net sees the first step's result
        seen['moved_first'] = ops.to_layout(moved_first, 'cl').cpu().numpy()
        return torch.from_numpy(f2).cuda()

    res = pipelines.two_steps_tail(moving, fixed, model1, model2, flow1, flow2, warp_interp=interp)
    # --- the reference's order of operations (bids_two_steps_registration.py:316-355) on the oracle
    if interp == 'linear':
        moved_first, w1 = oracle_tail(moving, flow1, 5)
        moved, w2 = oracle_tail(moved_first, f2, 5)
        warp = io.compose([w1[0], w2[0]])[None]
    else:
        _, w1 = oracle_tail(moving, flow1, 5)
        moved_first = io.transform_model(moving, w1, interp, rescale=2)
        _, w2 = oracle_tail(moved_first, f2, 5)
        warp = io.compose([w1[0], w2[0]])[None]
        moved = io.transform_model(moving, warp, interp, rescale=2)
    assert res['scale'] == 2
    check(seen['moved_first'], moved_first)
    check(res['warp'], warp)
    check(res['moved'], moved)
    # export (:504-546): x2 rescale, time axis, RAI components, for an oblique-free LPS-like affine
    affine = np.diag([-1.0, -1.0, 1.0, 1.0])
    got = pipelines.export_sct_warp(res['warp'], res['scale'], affine)
    # aff2axcodes(-affine) of this affine is ('R', 'A', 'I'): perm (0, 1, 2), no inversion (3d_reg.py:399-417)
    want = sct_oracle.apply(io.rescale_dense_transform(warp, 2)[0], ('R', 'A', 'I'))
    assert got.shape == full + (1, 3)
    check(got, want)


def test_two_steps_subvolumes_linear():
    rng = np.random.default_rng(5)
    tile, im = (16, 16, 32), (24, 16, 48)
    coords = [(0, 16, 0, 16, 0, 32), (8, 24, 0, 16, 0, 32), (0, 16, 0, 16, 16, 48), (8, 24, 0, 16, 16, 48)]
    moving = rng.random((1,) + im + (1,)).astype(np.float32)
    model1 = vxm.networks.VxmDense(tile, int_steps=5, svf_resolution=2, int_resolution=2)
    model2 = vxm.networks.VxmDense(tile, int_steps=5, svf_resolution=2, int_resolution=2)
    half = tuple(d // 2 for d in tile)
    sub_mov = [moving[:, c[0]:c[1], c[2]:c[3], c[4]:c[5]] for c in coords]
    sub_fx = [rng.random((1,) + tile + (1,)).astype(np.float32) for _ in coords]
    flows1 = [smooth(rng, (1,) + half + (3,), 1.0) for _ in coords]
    flows2 = [smooth(rng, (1,) + half + (3,), 0.5) for _ in coords]
    res = pipelines.two_steps_tail_subvol(moving, sub_mov, sub_fx, coords, model1, model2, flows1, flows2)
    fields = []
    for k in range(len(coords)):
        moved_first, w1 = oracle_tail(sub_mov[k], flows1[k], 5)
        _, w2 = oracle_tail(moved_first, flows2[k], 5)
        fields.append(io.compose([w1[0], w2[0]]))
    hc = [tuple(v // 2 for v in c) for c in coords]
    stitched = stitch_oracle.get_def_field_from_subvol(np.array(half), np.array(im) // 2, hc, fields)       # float64, :226-271
    moved = io.transform_model(moving, stitched[None].astype(np.float32), 'linear', rescale=2)
    assert res['scale'] == 2
    check(res['warp'], stitched[None].astype(np.float32))
    check(res['moved'], moved)


@pytest.mark.parametrize('interp', ['linear', 'nearest'])
def test_bids_two_steps_registration_script(tmp_path, interp, arithmetic_mode):
    """scripts/bids_two_steps_registration.py end to end (NIfTI + flow files in, moved image + SCT warp out, the
    reference's file names) against the chained oracle."""
    import json
    import os
    import runpy
    from multimodal_registration_b200 import _nifti
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rng = np.random.default_rng(11)
    full, half = (16, 32, 48), (8, 16, 24)
    moving = rng.random(full).astype(np.float32)
    fixed = rng.random(full).astype(np.float32)
    if interp == 'nearest':
        moving = np.floor(moving * 4).astype(np.float32)
    flow1 = smooth(rng, (1,) + half + (3,), 1.2)[0]
    flow2 = smooth(rng, (1,) + half + (3,), 0.6)[0]
    affine = np.array([[0, 0, 2.0, -10], [-1.5, 0, 0, 4], [0, 1.0, 0, 7], [0, 0, 0, 1]])      # a permuted, flipped orientation
    mp, fp = str(tmp_path / 'sub-03_T2w_proc.nii.gz'), str(tmp_path / 'sub-03_T1w_proc.nii.gz')
    _nifti.save_nifti(moving, mp, affine)
    _nifti.save_nifti(fixed, fp, affine)
    np.save(str(tmp_path / 'flow1.npy'), flow1)
    _nifti.save_nifti(flow2[:, :, :, None, :], str(tmp_path / 'flow2.nii.gz'), affine, intent_code=1007)
    cfg = json.load(open(os.path.join(root, 'config', 'config_inference.json')))
    cfg['warp_interpolation'] = interp
    json.dump(cfg, open(str(tmp_path / 'cfg.json'), 'w'))
    mod = runpy.run_path(os.path.join(root, 'scripts', 'bids_two_steps_registration.py'))
    assert mod['main'](['--model1-path', str(tmp_path / 'flow1.npy'), '--model2-path', str(tmp_path / 'flow2.nii.gz'),
                        '--config-path', str(tmp_path / 'cfg.json'), '--fx-img-path', fp, '--mov-img-path', mp,
                        '--fx-img-contrast', 'T1w']) == 0
    moved, aff2 = _nifti.load_nifti(str(tmp_path / 'sub-03_T2w_proc_reg_to_T1w.nii.gz'))
    warp, _, hdr = _nifti.load_nifti(str(tmp_path / 'sub-03_T2w_proc_field_to_T1w.nii.gz'), return_header=True)
    np.testing.assert_allclose(aff2, affine)
    mv5 = moving[None, ..., None]
    steps = cfg['int_steps']
    if interp == 'linear':
        moved_first, w1 = oracle_tail(mv5, flow1[None], steps)
        want_moved, w2 = oracle_tail(moved_first, flow2[None], steps)
        want_warp = io.compose([w1[0], w2[0]])
    else:
        _, w1 = oracle_tail(mv5, flow1[None], steps)
        moved_first = io.transform_model(mv5, w1, interp, rescale=2)
        _, w2 = oracle_tail(moved_first, flow2[None], steps)
        want_warp = io.compose([w1[0], w2[0]])
        want_moved = io.transform_model(mv5, want_warp[None], interp, rescale=2)
    check(moved, want_moved[0, ..., 0])
    axcodes = _nifti.aff2axcodes(-affine)
    want_sct = sct_oracle.apply(io.rescale_dense_transform(want_warp[None], 2)[0], axcodes)
    assert warp.shape == full + (1, 3) and int(hdr['intent_code']) == 1007
    check(warp, want_sct)
