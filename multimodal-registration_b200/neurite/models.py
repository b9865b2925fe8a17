"""ne.models.labels_to_image mirror (train_synthmorph.py:258-268,288-289).

``labels_to_image(**gen_args, id=k)`` builds the SynthMorph generator: a label map is deformed by an integrated random
SVF (draw_perlin -> VecInt(5) -> RescaleTransform(2) -> nearest SpatialTransformer with fill_value 0 -- the hot path's
kernels) and turned into a grayscale image (per-label Gaussian intensities, Gaussian blur, multiplicative bias field,
clipping, min-max normalisation, gamma) plus a one-hot label map.  The voxel passes of the intensity model are the
kernels of dfm_synth.cu behind the C ABI.  The semantics follow the published SynthMorph generator as recalled
(`[UR]`: neurite is not vendored); random numbers come from a counter-based Philox stream / torch's generator and
cannot match TensorFlow's, so parity is distributional (tests/test_synth_gpu.py).

The returned object is callable on a label map ``[B, X, Y, Z, 1]`` (numpy or tensor) and returns
``[image [B, X, Y, Z, 1], one_hot [B, X, Y, Z, C]]`` as device tensors, like ``gen_model.outputs`` (:289-290);
``.predict`` returns numpy arrays."""
import numpy as np
import torch

from .. import _host, _lib, ops
from ..ops import _ptr, _stream
from .utils import augment


def gaussian_taps(sigma, max_sigma):
    """1-D Gaussian kernel, width 2 * ceil(2.5 * max_sigma) + 1 (fixed by the model's blur_std), normalised."""
    half = int(np.ceil(2.5 * max(max_sigma, 1e-6)))
    x = np.arange(-half, half + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / max(sigma, 1e-6)) ** 2) if sigma > 1e-6 else (x == 0).astype(np.float64)
    return (k / k.sum()).astype(np.float32)


class LabelsToImage:
    def __init__(self, in_shape, in_label_list, out_label_list=None, warp_std=0.5, warp_res=(8, 16, 32), blur_std=1.0,
                 bias_std=0.3, bias_res=40, gamma_std=0.25, mean_min=25.0, mean_max=225.0, std_min=5.0, std_max=25.0,
                 vel_int_steps=5, id=0, seeds=None, return_def=False):
        self.in_shape = tuple(int(d) for d in in_shape)
        if len(self.in_shape) != 3:
            raise NotImplementedError('labels_to_image: 3-D label maps only (the reference uses 160 x 160 x 192)')
        self.in_labels = np.unique(np.asarray(in_label_list).astype(np.int64))
        out = self.in_labels if out_label_list is None else out_label_list
        # out_label_list: list (labels kept, in order) or dict {input label: output label}
        mapping = dict(out) if isinstance(out, dict) else {int(l): int(l) for l in np.asarray(out).reshape(-1)}
        self.out_labels = np.unique(np.asarray(list(mapping.values()), dtype=np.int64))
        lut = -np.ones(int(self.in_labels.max()) + 1, np.int32)
        for src, dst in mapping.items():
            if 0 <= int(src) < lut.size:
                lut[int(src)] = int(np.searchsorted(self.out_labels, dst))     # channel index of the output label
        self._lut_host = lut
        self.warp_std, self.blur_std, self.bias_std, self.gamma_std = float(warp_std), float(blur_std), float(bias_std), float(gamma_std)
        self.warp_res = [warp_res] if np.isscalar(warp_res) else list(warp_res)
        self.bias_res = [bias_res] if np.isscalar(bias_res) else list(bias_res)
        self.mean_range, self.std_range = (float(mean_min), float(mean_max)), (float(std_min), float(std_max))
        self.vel_int_steps = int(vel_int_steps)
        self.id, self.return_def = int(id), bool(return_def)
        self.gen = None
        self.seed = None if seeds is None else int(dict(seeds).get('all', 0))
        self._calls = 0

    # ------------------------------------------------------------------------------------------------
    def _generator(self, dev):
        if self.gen is None or self.gen.device != dev:
            self.gen = torch.Generator(device=dev)
            self.gen.manual_seed(self.seed if self.seed is not None else (torch.seed() + 7919 * self.id) % (2 ** 62))
        return self.gen

    def deform(self, labels):
        """draw_perlin SVF at half resolution -> VecInt -> RescaleTransform(2) -> nearest warp with fill_value 0."""
        B = labels.shape[0]
        half = tuple(d // 2 for d in self.in_shape)
        gen = self._generator(labels.device)
        vel = torch.stack([augment.draw_perlin(half + (3,), scales=[r / 2 for r in self.warp_res], max_std=self.warp_std / 2,
                                               seeds={'noise': int(torch.randint(0, 2 ** 31 - 1, (1,), generator=gen, device=labels.device))})
                           for _ in range(B)], 0)
        flow = ops.rescale_dense_transform(ops.vecint(vel, self.vel_int_steps), 2)
        return ops.warp(labels, flow, 'nearest', fill_value=0), flow

    @torch.no_grad()
    def __call__(self, labels):
        labels = _host.to_device(labels, torch.float32, tag='labels')
        if labels.dim() == 4:
            labels = labels[..., None]
        if tuple(labels.shape[1:4]) != self.in_shape or labels.shape[-1] != 1:
            raise ValueError('labels must be [B, %d, %d, %d, 1], got %s' % (self.in_shape + (tuple(labels.shape),)))
        dev = labels.device
        B, (X, Y, Z) = labels.shape[0], self.in_shape
        n = X * Y * Z
        gen = self._generator(dev)
        warped, flow = self.deform(labels) if self.warp_std > 0 else (labels, None)
        warped = ops.to_layout(warped, 'cl').contiguous()
        nlab = int(self.in_labels.max()) + 1
        # per-item, per-label intensity statistics
        means = self.mean_range[0] + (self.mean_range[1] - self.mean_range[0]) * torch.rand((B, nlab), generator=gen, device=dev)
        stds = self.std_range[0] + (self.std_range[1] - self.std_range[0]) * torch.rand((B, nlab), generator=gen, device=dev)
        img = torch.empty((B, X, Y, Z), device=dev, dtype=torch.float32)
        tmp = torch.empty_like(img)
        self._calls += 1
        for b in range(B):
            seed = int(torch.randint(0, 2 ** 62, (1,), generator=gen, device=dev))
            _lib.call('dfm_synth_intensity', _ptr(warped[b]), _ptr(means[b]), _ptr(stds[b]), nlab, seed, _ptr(img[b]), n, _stream())
        # separable Gaussian blur (sigma ~ U(0, blur_std) per item and axis), ping-pong between two buffers per item;
        # then the multiplicative bias field exp(perlin) and the clip to [0, 255] land the result back in `img`
        for b in range(B):
            cur, oth = img[b], tmp[b]
            if self.blur_std > 0:
                for axis in range(3):
                    sigma = self.blur_std * float(torch.rand((), generator=gen, device=dev))
                    taps = torch.from_numpy(gaussian_taps(sigma, self.blur_std)).to(dev)
                    _lib.call('dfm_conv1d_axis', _ptr(cur), _ptr(oth), 1, X, Y, Z, axis, _ptr(taps), int(taps.numel()), _stream())
                    cur, oth = oth, cur
            logbias = None
            if self.bias_std > 0:
                logbias = augment.draw_perlin(self.in_shape + (1,), scales=self.bias_res, max_std=self.bias_std,
                                              seeds={'noise': int(torch.randint(0, 2 ** 31 - 1, (1,), generator=gen, device=dev))}).contiguous()
            _lib.call('dfm_scale_exp_clip', _ptr(cur), _ptr(logbias), _ptr(img[b]), n, 0.0, 255.0, _stream())
        # min-max normalisation and gamma augmentation, per item
        mm = torch.empty((B, 2), device=dev, dtype=torch.float64)
        work = torch.empty(max(_lib.load().dfm_metrics_workspace_bytes() // 8, 1), device=dev, dtype=torch.float64)
        for b in range(B):
            _lib.call('dfm_minmax', _ptr(img[b]), n, 0, _ptr(mm[b]), _ptr(work), _stream())
        gamma = torch.exp(self.gamma_std * torch.randn((B,), generator=gen, device=dev)).float().contiguous()
        _lib.call('dfm_norm_gamma', _ptr(img), _ptr(mm), _ptr(gamma), _ptr(img), B, n, _stream())
        # one-hot output labels (channels-last, like the reference)
        C = int(self.out_labels.size)
        lut = torch.from_numpy(self._lut_host).to(dev)
        onehot = torch.empty((B, X, Y, Z, C), device=dev, dtype=torch.float32)
        _lib.call('dfm_onehot', _ptr(warped), _ptr(lut), int(lut.numel()), C, _ptr(onehot), B * n, _stream())
        if C <= 255:
            # the map is one-hot by construction: it carries its channel indices (255 = no channel) so that the
            # SpatialTransformer that produces `pred` (train_synthmorph.py:298) can warp it from 8 bytes per voxel instead of
            # 8 x C floats (ops.warp_onehot, same bits).  Any tensor operation on the map yields a plain tensor without them.
            lab = warped[..., 0].long()
            idx = lut.long()[lab.clamp(0, lut.numel() - 1)]
            ok = (lab >= 0) & (lab < lut.numel()) & (idx >= 0) & (idx < C)
            onehot.dfm_labels = (torch.where(ok, idx, torch.full_like(idx, 255)).to(torch.uint8), C)
        outs = [img[..., None], onehot]
        if self.return_def:
            outs.append(flow)
        return outs

    def predict(self, labels, **kwargs):
        return [_host.to_host(t, tag='gen%d' % i) for i, t in enumerate(self(labels))]


def labels_to_image(in_shape, in_label_list, out_label_list=None, **kwargs):
    """Same call as ``ne.models.labels_to_image(in_shape=..., in_label_list=..., out_label_list=..., warp_std=...,
    warp_res=..., blur_std=..., bias_std=..., bias_res=..., gamma_std=..., id=...)`` at train_synthmorph.py:258-289."""
    return LabelsToImage(in_shape, in_label_list, out_label_list, **kwargs)
