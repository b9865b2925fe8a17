"""vxm.networks mirror: Transform and the deformation tail of VxmDense.

``predict`` follows Keras: numpy (host) inputs in, numpy out; ``__call__`` keeps tensors on
the device.  Host inputs go through pinned staging buffers (see ``_host``).
"""
import types

import numpy as np
import torch

from .. import _host, ops
from . import layers


class Transform(torch.nn.Module):
    """[scan, trf] -> RescaleTransform(rescale)? -> SpatialTransformer
    (gen_apply_def_field.py:74-76, 3d_reg.py:331-334,377-380, bids_*.py)."""

    def __init__(self, inshape, affine=False, interp_method='linear', rescale=None,
                 fill_value=None, nb_feats=1):
        super().__init__()
        if affine:
            raise NotImplementedError('Transform(affine=True) is not on the reference hot path')
        self.inshape = tuple(int(d) for d in inshape)
        if len(self.inshape) != 3:
            raise NotImplementedError('Transform: only 3-D volumes are supported')
        self.nb_feats = nb_feats
        self.rescale = rescale
        self.trf_shape = self.inshape if rescale is None else tuple(int(d / rescale) for d in self.inshape)
        self.rescaler = layers.RescaleTransform(rescale) if rescale is not None else None
        self.transformer = layers.SpatialTransformer(interp_method=interp_method, fill_value=fill_value)

    def _check(self, scan, trf):
        if tuple(scan.shape[1:]) != self.inshape + (self.nb_feats,):
            raise ValueError('Transform: scan input has shape %s, expected [B, %s]'
                             % (tuple(scan.shape), ', '.join(map(str, self.inshape + (self.nb_feats,)))))
        if tuple(trf.shape[1:]) != self.trf_shape + (3,):
            raise ValueError('Transform: transform input has shape %s, expected [B, %s]'
                             % (tuple(trf.shape), ', '.join(map(str, self.trf_shape + (3,)))))

    def forward(self, inputs):
        scan = _host.to_device(inputs[0], torch.float32, tag='scan')
        trf = _host.to_device(inputs[1], torch.float32, tag='trf')
        self._check(scan, trf)
        if self.rescaler is not None:
            tr = self.transformer
            fusable = (tr.interp_method in ('linear', 'nearest') and self.rescale >= 1 and self.nb_feats == 1 and
                       tr.indexing == 'ij' and not tr.single_transform and
                       not (torch.is_grad_enabled() and (scan.requires_grad or trf.requires_grad)))
            if fusable:        # the rescaled field is an intermediate: one kernel (dfm_rescale_warp*_fwd), same result
                return ops.rescale_warp(scan, trf, self.rescale, tr.fill_value, tr.interp_method)
            trf = self.rescaler(trf)
        return self.transformer([scan, trf])

    @torch.no_grad()
    def predict(self, inputs, copy=True, **kwargs):
        return _host.to_host(self.forward(inputs), copy=copy)


class VxmDense(torch.nn.Module):
    """Deformation tail of vxm.networks.VxmDense: given the flow the
    U-Net's flow convolution emits, ``RescaleTransform`` to the SVF / integration resolution,
    ``VecInt(int_steps)``, ``RescaleTransform`` back to full resolution and the linear
    ``SpatialTransformer`` on the source (3d_reg.py:305,310, bids_*.py:311-322,
    train_synthmorph.py:296-297).

    The U-Net itself (dense convolutions) is outside the hot path: pass ``flow_model``, a
    callable ``(source, target) -> flow``, to get the full ``[source, target]`` call; without
    it, call ``deform([source, flow])``.  Outputs ``[y_source, preint_flow]`` like the
    reference default (``reg_field='preintegrated'``); ``references.pos_flow`` holds the
    full-resolution integrated warp of the last call (train_synthmorph.py:297).
    """

    def __init__(self, inshape, nb_unet_features=None, int_steps=7, svf_resolution=1,
                 int_resolution=2, fill_value=None, reg_field='preintegrated', flow_model=None,
                 fuse_rescale_warp=True, **kwargs):
        super().__init__()
        self.inshape = tuple(int(d) for d in inshape)
        if len(self.inshape) != 3:
            raise NotImplementedError('VxmDense: only 3-D volumes are supported')
        if reg_field not in ('svf', 'preintegrated', 'postintegrated', 'warp'):
            raise ValueError('Unknown option "%s" for reg_field.' % reg_field)
        self.int_steps = int_steps
        self.svf_resolution = svf_resolution
        self.int_resolution = int_resolution
        self.reg_field = reg_field
        self.flow_model = flow_model
        # one fused kernel for the last RescaleTransform + warp wherever the full-resolution field is only an
        # intermediate (deform(keep_pos_flow=False), i.e. predict_deform unless reg_field asks for the warp): the
        # up-sampler's march on the LSU pipe, the image corners by texture gathers (dfm_warp_tex.cu) -- 0.73 ms
        # instead of 0.39 + 0.81 ms at B=32, bit-identical to the two kernels
        self.fuse_rescale_warp = fuse_rescale_warp
        self.svf_size = tuple(int(np.round(d / svf_resolution)) for d in self.inshape)
        self.int_size = tuple(int(np.round(d / int_resolution)) for d in self.inshape)
        self.integrator = layers.VecInt(method='ss', int_steps=int_steps) if int_steps > 0 else None
        self.transformer = layers.SpatialTransformer(interp_method='linear', indexing='ij', fill_value=fill_value)
        self.references = types.SimpleNamespace(pos_flow=None, svf=None, preint_flow=None)

    def deform(self, inputs, keep_pos_flow=True):
        """keep_pos_flow=False (inference): the final RescaleTransform and the warp of a
        one-channel source run as ONE fused kernel and ``references.pos_flow`` is not produced."""
        keep_pos_flow = keep_pos_flow or self.reg_field in ('postintegrated', 'warp')      # the second output needs it
        source = _host.to_device(inputs[0], torch.float32, tag='source')
        flow = _host.to_device(inputs[1], torch.float32, tag='flow')
        pre_svf_size = tuple(flow.shape[1:-1])
        svf_size = self.svf_size
        if pre_svf_size != svf_size:
            flow = ops.rescale_dense_transform(flow, svf_size[0] / pre_svf_size[0])
        svf = flow
        if self.int_steps > 0 and self.int_resolution > 1 and svf_size != self.int_size:
            flow = ops.rescale_dense_transform(flow, self.int_size[0] / svf_size[0])
        preint_flow = flow
        pos_flow = flow
        y_source = None
        if self.int_steps > 0:
            pos_flow = self.integrator(pos_flow)
            if self.int_resolution > 1:
                factor = self.inshape[0] / self.int_size[0]
                fusable = (self.fuse_rescale_warp and not keep_pos_flow and source.shape[-1] == 1 and factor >= 1 and
                           not (torch.is_grad_enabled() and (pos_flow.requires_grad or source.requires_grad)))
                if fusable:
                    y_source = ops.rescale_warp(source, pos_flow, factor, self.transformer.fill_value)
                    pos_flow = None
                else:
                    pos_flow = ops.rescale_dense_transform(pos_flow, factor)
        if y_source is None:
            y_source = self.transformer([source, pos_flow])
        self.references.pos_flow, self.references.svf, self.references.preint_flow = pos_flow, svf, preint_flow
        second = {'svf': svf, 'preintegrated': preint_flow}.get(self.reg_field, pos_flow)
        return [y_source, second]

    def deform_graphed(self, inputs):
        """``deform`` through a cached CUDA graph (one per input shape): single-volume inference is
        launch-bound, a replay submits the whole tail in one call.  Device tensors in, static device
        tensors out (valid until the next call with the same shapes)."""
        source = _host.to_device(inputs[0], torch.float32, tag='source')
        flow = _host.to_device(inputs[1], torch.float32, tag='flow')
        key = (tuple(source.shape), tuple(flow.shape), ops.layout_of(source), ops.layout_of(flow))
        if not hasattr(self, '_graphs'):
            self._graphs = {}
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = ops.Graphed(lambda s, f: tuple(self.deform([s, f])), source, flow)
        return list(g(source, flow))

    def forward(self, inputs):
        if self.flow_model is None:
            raise NotImplementedError(
                'VxmDense: the U-Net is outside the B200 hot path; construct with flow_model=<callable '
                '(source, target) -> flow> or call deform([source, flow]) with the flow it would emit')
        source = _host.to_device(inputs[0], torch.float32, tag='source')
        target = _host.to_device(inputs[1], torch.float32, tag='target')
        return self.deform([source, self.flow_model(source, target)])

    @torch.no_grad()
    def predict(self, inputs, copy=True, **kwargs):
        return [_host.to_host(t, tag='out%d' % i, copy=copy) for i, t in enumerate(self.forward(inputs))]

    @torch.no_grad()
    def predict_deform(self, inputs, copy=True, batch_size=2):
        """Keras-style (numpy in / numpy out) call of the deformation tail.

        Host inputs are processed in chunks of ``batch_size`` items on three CUDA streams (H2D copy,
        kernels, D2H copy), so the PCIe transfers of neighbouring chunks overlap the kernels and each
        other.  A second output that is the untouched input flow (``reg_field='preintegrated'`` with
        the flow already at integration resolution) is returned from the host copy, not re-downloaded.
        copy=False returns views of cached pinned buffers (valid until the next call)."""
        keep = self.reg_field in ('postintegrated', 'warp')
        src_h, flow_h = inputs[0], inputs[1]
        on_host = not (isinstance(src_h, torch.Tensor) and src_h.is_cuda) and \
            not (isinstance(flow_h, torch.Tensor) and flow_h.is_cuda)
        B = int(src_h.shape[0])
        if not on_host or B <= batch_size:
            outs = self.deform(inputs, keep_pos_flow=keep)
            return [_host.to_host(t, tag='out%d' % i, copy=copy) for i, t in enumerate(outs)]

        src_p = _host.pinned_view(src_h, torch.float32, 'source')
        flow_p = _host.pinned_view(flow_h, torch.float32, 'flow')
        dev = _host.device()
        cur = torch.cuda.current_stream()
        s_in, s_out = _host.side_streams()
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        y_host = second_host = None
        second_is_input = False
        for lo in range(0, B, batch_size):
            hi = min(lo + batch_size, B)
            with torch.cuda.stream(s_in):
                src_d = src_p[lo:hi].to(dev, non_blocking=True)
                flow_d = flow_p[lo:hi].to(dev, non_blocking=True)
            cur.wait_stream(s_in)
            # chunk tensors are released as soon as the streams that use them have passed (record_stream), so
            # device memory is O(batch_size), not O(B)
            src_d.record_stream(cur)
            flow_d.record_stream(cur)
            y, second = self.deform([src_d, flow_d], keep_pos_flow=keep)
            second_is_input = second is flow_d
            if y_host is None:
                y_host = _host.pinned_out((B,) + tuple(y.shape[1:]), y.dtype, 'out0')
                if not second_is_input:
                    second_host = _host.pinned_out((B,) + tuple(second.shape[1:]), second.dtype, 'out1')
            y_c = ops.to_layout(y, 'cl')
            sec_c = None if second_is_input else ops.to_layout(second, 'cl')
            s_out.wait_stream(cur)
            with torch.cuda.stream(s_out):
                y_host[lo:hi].copy_(y_c, non_blocking=True)
                if sec_c is not None:
                    second_host[lo:hi].copy_(sec_c, non_blocking=True)
            y_c.record_stream(s_out)
            if sec_c is not None:
                sec_c.record_stream(s_out)
        s_out.synchronize()
        cur.wait_stream(s_out)
        torch.cuda.current_stream().synchronize()
        y_np = _host.host_copy(y_host) if copy else y_host.numpy()
        if second_is_input:
            sec_np = flow_h if isinstance(flow_h, np.ndarray) else flow_p.numpy()
            sec_np = np.asarray(sec_np, dtype=np.float32)
        else:
            sec_np = _host.host_copy(second_host) if copy else second_host.numpy()
        return [y_np, sec_np]
