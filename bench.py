#!/usr/bin/env python
"""bench.py -- headline benchmark of the deformation hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1], "3d_reg.py inference"): per volume pair, a SynthMorph-shaped
half-resolution SVF [80,80,96,3] -> 7-step scaling-and-squaring VecInt -> x2 RescaleTransform ->
trilinear SpatialTransformer of a 160x160x192 image (the last two as one fused kernel: the full-resolution
field is an intermediate of the inference call).  One "step" = that pipeline over a batch of B volume pairs
per GPU (B=32: SVFs 236 MB, SS ping-pong fields 472 MB, images in + out 1.26 GB -- far larger than the 126 MB
L2, so no L2 flush is needed between iterations).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Own arm:   `value` = warped voxels/s with inputs resident in HBM (CUDA events, max over ranks);
           `e2e`   = the same through VxmDense.predict_deform with pinned HOST buffers, copies
                     inside the timed region; `roofline` for the dominant kernel; `cpu_baseline`
                     = the restated reference (oracle/torch_oracle.py) on the host cores.
Reference arm (--impl reference): the reference's CPU path.  TensorFlow/voxelmorph/neurite are
not installable here (no network, no wheels), so this times the oracle's restatement of the
reference algorithm (gather formulation, fp32, torch-CPU, all host threads) on a bounded sample
(one volume pair per step) of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FULL = (160, 160, 192)
HALF = (80, 80, 96)
INT_STEPS = 7
N_F = FULL[0] * FULL[1] * FULL[2]
N_H = HALF[0] * HALF[1] * HALF[2]
# algorithmic bytes per volume pair (unique external input + output bytes)
BYTES_SS_STEP = 24 * N_H                       # read v (12 B/voxel), write v' (12 B/voxel)
BYTES_RESCALE = 12 * N_H + 12 * N_F            # read half-res field, write full-res field
BYTES_WARP = (8 * 1 + 12) * N_F                # read image + field, write image (C = 1)
BYTES_FUSED = 12 * N_H + 8 * N_F               # fused rescale+warp: read coarse field + image, write image
METRIC = 'warped voxels/sec (160x160x192, 7-step VecInt+warp)'
UNIT = 'voxels/s'


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def synth_inputs(batch, device, seed):
    """Seeded synthetic inputs: smooth SVF (std 3 voxels, config.json vel_std) and U[0,1) images."""
    import torch
    g = torch.Generator(device='cpu')
    g.manual_seed(1234 + seed)
    coarse = torch.randn(batch, 3, 10, 10, 12, generator=g) * 3.0
    svf = torch.nn.functional.interpolate(coarse, size=HALF, mode='trilinear', align_corners=True)
    svf = svf.permute(0, 2, 3, 4, 1).contiguous()                 # channels-last like the reference
    img = torch.rand(batch, *FULL, 1, generator=g)
    return svf, img


# ---------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the restated reference on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_pipeline_seconds(n_items, seed=0):
    import torch
    from oracle import torch_oracle as to
    torch.set_num_threads(os.cpu_count() or 1)
    svf, img = synth_inputs(n_items, 'cpu', seed)
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(n_items):
            to.headline_pipeline(svf[i:i + 1], img[i:i + 1], INT_STEPS)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    sample = '1 volume pair per step (of the B=%d per-GPU batch), restated reference on %d host threads' % (args.batch, cores)
    for _ in range(args.warmup):
        cpu_pipeline_seconds(1)
    t = 0.0
    for s in range(args.steps):
        t += cpu_pipeline_seconds(1, seed=s)
    ms = 1e3 * t / max(args.steps, 1)
    value = N_F / (ms * 1e-3)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                         'note': 'TensorFlow/voxelmorph/neurite cannot be installed (no network); oracle/torch_oracle.py '
                                 'restates their algorithm (torch %s CPU)' % torch.__version__},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {'workload': '3d_reg.py inference tail: VecInt(7 steps) @80x80x96 -> RescaleTransform(2) -> '
                        'linear SpatialTransformer of one 160x160x192 image (C=1), per volume pair',
            'volume_pairs_per_gpu_per_step': args.batch, 'int_steps': INT_STEPS,
            'l2': 'inputs larger than L2 (no flush): per step %.2f GB of fields and images per GPU'
                  % ((args.batch * (12 * N_H * 3 + 8 * N_F)) / 1e9),
            'sharding': 'volume pairs sharded across ranks, no data-path collective'}


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------
def run_own(args):
    import torch
    import torch.distributed as dist
    import multimodal_registration_b200 as mrb
    from multimodal_registration_b200 import ops

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the deformation engine has no CPU path '
                         '(use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    mrb._lib.load()

    B = args.batch
    svf_h, img_h = synth_inputs(B, 'cpu', int(os.environ.get('BENCH_SEED', rank)))      # BENCH_SEED: tuning aid (another rank's data on one GPU)
    svf_pin, img_pin = svf_h.pin_memory(), img_h.pin_memory()
    svf, img = svf_pin.to(dev), img_pin.to(dev)
    model = mrb.voxelmorph.networks.VxmDense(FULL, int_steps=INT_STEPS, svf_resolution=2, int_resolution=2)   # fuses rescale+warp at inference

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(events=None):
        # production inference tail: 7 SS steps, then RescaleTransform(2) + linear warp as ONE kernel (the
        # full-resolution field is an intermediate of VxmDense's inference call and never reaches HBM)
        if events is not None:
            events[0].record()
        flow = ops.vecint(svf, INT_STEPS)
        if events is not None:
            events[1].record()
        out = ops.rescale_warp(img, flow, 2)
        if events is not None:
            events[2].record()
        return out

    def step_unfused(events):
        # the two stand-alone kernels (what a caller that keeps the full-resolution field runs), reported beside
        flow = ops.vecint(svf, INT_STEPS)
        events[0].record()
        flow = ops.rescale_dense_transform(flow, 2)
        events[1].record()
        out = ops.warp(img, flow)
        events[2].record()
        return out

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_start.record()
        for k in range(args.steps):
            step(ev[k])
        t_end.record()
        barrier()
        total_ms = t_start.elapsed_time(t_end)
        stage_ms = [sum(e[i].elapsed_time(e[i + 1]) for e in ev) / args.steps for i in range(2)]      # ss, fused
        # the stand-alone up-sampler and warp (outside the headline region)
        n_u = max(3, min(args.steps, 10))
        evu = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_u)]
        step_unfused(evu[0])
        for k in range(n_u):
            step_unfused(evu[k])
        torch.cuda.synchronize()
        stage_ms += [sum(e[i].elapsed_time(e[i + 1]) for e in evu) / n_u for i in range(2)]           # rescale, warp

        # ---- e2e: numpy-in / numpy-out through the Keras-style call, pinned host buffers ----
        e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 10))
        for _ in range(0 if args.no_e2e else 2):
            model.predict_deform([img_pin, svf_pin], copy=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            moved, _pre = model.predict_deform([img_pin, svf_pin], copy=False)
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / max(e2e_steps, 1)
        clocks = sampler.stop() if rank == 0 else None

        # ---- e2e ceiling: the step's host<->device bytes with NO kernels (one H2D stream, one D2H stream, every
        # rank at once) -- what the PCIe / host-memory path of the box allows for this step ----
        ceil_ms = page_ms = float('nan')
        if not args.no_e2e:
            d_svf, d_img = torch.empty_like(svf), torch.empty_like(img)
            h_out = torch.empty(img.shape, dtype=torch.float32).pin_memory()
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

            def copy_only():
                with torch.cuda.stream(s_up):
                    d_svf.copy_(svf_pin, non_blocking=True)
                    d_img.copy_(img_pin, non_blocking=True)
                with torch.cuda.stream(s_dn):
                    h_out.copy_(d_img, non_blocking=True)
            copy_only()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                copy_only()
            barrier()
            ceil_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
            del d_svf, d_img, h_out
            # the pageable-numpy entry (what nibabel hands the reference: plain arrays in, fresh arrays out), once
            if rank == 0:
                svf_np, img_np = svf_h.numpy(), img_h.numpy()
                model.predict_deform([img_np, svf_np], copy=True)
                t0 = time.perf_counter()
                model.predict_deform([img_np, svf_np], copy=True)
                page_ms = 1e3 * (time.perf_counter() - t0)

    ms = torch.tensor([total_ms / args.steps, e2e_ms, ceil_ms] + stage_ms, device=dev, dtype=torch.float64)   # + ss, fused, rescale, warp
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step, e2e_ms, ceil_ms, ss_ms, fu_ms, rs_ms, wp_ms = [float(v) for v in ms.tolist()]

    # ---- the other BASELINE.json configs (3: training-step tail incl. the NCCL all-reduce, 4: two-step cascade
    # over 64 subjects, 5: Jacobian 256^3), every rank takes part; reported beside the headline ----
    configs = None
    if not (args.no_configs or args.no_e2e):
        del svf, img, model
        torch.cuda.empty_cache()
        configs = run_configs(rank, world, dev)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        voxels = world * B * N_F
        kernels = {
            K_SS: {'launches_per_step': INT_STEPS, 'ms_per_launch': ss_ms / INT_STEPS,
                   'algorithmic_bytes_per_launch': B * BYTES_SS_STEP, 'in_timed_region': True,
                   'note': 'average of the 7 steps; steps 6 and 7 launch a halo-2 and a halo-3 variant, the one not selected per item exits'},
            K_FU: {'launches_per_step': 1, 'ms_per_launch': fu_ms,
                   'algorithmic_bytes_per_launch': B * BYTES_FUSED, 'in_timed_region': True,
                   'note': 'bound by the texture pipe (tld4 write-back), not HBM: the fusion removes 24 B/voxel of field traffic, so '
                           'its own fraction of the HBM roofline is low by construction; equivalent_unfused_frac_of_peak rates the '
                           'same time against the bytes of the two kernels it replaces'},
            K_RS: {'launches_per_step': 0, 'ms_per_launch': rs_ms,
                   'algorithmic_bytes_per_launch': B * BYTES_RESCALE, 'in_timed_region': False},
            K_WP: {'launches_per_step': 0, 'ms_per_launch': wp_ms,
                   'algorithmic_bytes_per_launch': B * BYTES_WARP, 'in_timed_region': False},
        }
        for k in kernels.values():
            k['achieved_gbs'] = k['algorithmic_bytes_per_launch'] / (k['ms_per_launch'] * 1e-3) / 1e9
            k['frac_of_peak'] = k['achieved_gbs'] / peak
            k['share_of_step'] = k['ms_per_launch'] * k['launches_per_step'] / (ss_ms + fu_ms)
        kernels[K_FU]['equivalent_unfused_frac_of_peak'] = B * (BYTES_RESCALE + BYTES_WARP) / (fu_ms * 1e-3) / 1e9 / peak
        dom_name = max(kernels, key=lambda n: kernels[n]['share_of_step'])
        dom = kernels[dom_name]
        cores = os.cpu_count() or 1
        n_cpu = 24                                   # ~0.55 s per pair on 16 host threads -> 10-15 s of CPU work
        cpu_s = float('nan') if args.no_e2e else cpu_pipeline_seconds(n_cpu)
        line = {
            'metric': METRIC, 'value': voxels / (ms_per_step * 1e-3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(args),
            'e2e': {'value': voxels / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': int(svf_pin.numel() * 4 + img_pin.numel() * 4),
                    'd2h_bytes_per_step': int(img_pin.numel() * 4),
                    'note': 'chunked 3-stream pipeline (batch_size 2); the second output (pre-integration flow) is the '
                            'untouched input and is returned from the host copy, not re-downloaded',
                    'ceiling_ms_per_step': ceil_ms, 'ceiling_value': voxels / (ceil_ms * 1e-3),
                    'ceiling_gbs': {'h2d': world * (svf_pin.numel() + img_pin.numel()) * 4 / (ceil_ms * 1e-3) / 1e9,
                                    'd2h': world * img_pin.numel() * 4 / (ceil_ms * 1e-3) / 1e9},
                    'frac_of_ceiling': ceil_ms / e2e_ms,
                    'ceiling_note': 'copy-only probe: the same H2D + D2H bytes per step on two streams, no kernels, all ranks at once '
                                    '(aggregate GB/s over the ranks)',
                    'pageable_ms_per_step': page_ms,
                    'pageable_note': 'rank 0 alone, plain (pageable) numpy arrays in, fresh numpy arrays out (copy=True): adds a host '
                                     'memcpy into / out of the pinned staging buffers',
                    'api': 'voxelmorph.networks.VxmDense(...).predict_deform([source, flow]) on pinned host arrays'},
            # per step: 7 SS steps (the last two launch a halo-2 and a halo-3 variant of k_ss_march, each batch item runs
            # in one of them) + k_rescale_warp_tex = 10 kernels (profiles/r2_launches_bench_b32.csv)
            'gpu_launches': args.steps * (INT_STEPS + 2 + 1),
            'roofline': {'bound': 'hbm', 'kernel': dom_name, 'achieved': dom['achieved_gbs'], 'peak': peak,
                         'unit': 'GB/s', 'frac': dom['frac_of_peak'],
                         'traffic': TRAFFIC_NCU_B32_GB.get(dom_name) if B == 32 else None,
                         'traffic_unit': 'GB per launch: dram__bytes_read.sum + dram__bytes_write.sum of this kernel, ncu --set full '
                                         'capture of THIS command at B=32 (profiles/r2_b32_ncu_full_summary.csv)',
                         'achieved_bytes_per_launch_GB': dom['algorithmic_bytes_per_launch'] / 1e9,
                         'peak_source': peak_src,
                         'pipeline_achieved': B * (INT_STEPS * BYTES_SS_STEP + BYTES_FUSED) / (ms_per_step * 1e-3) / 1e9,
                         'pipeline_equivalent_unfused': B * (INT_STEPS * BYTES_SS_STEP + BYTES_RESCALE + BYTES_WARP)
                         / (ms_per_step * 1e-3) / 1e9,
                         'pipeline_note': 'pipeline_achieved: GB/s on the bytes the fused pipeline has to move (149.9 MB per pair); '
                                          'pipeline_equivalent_unfused: the same time against the 267.9 MB per pair of SURVEY 8(d), '
                                          'where the full-resolution field makes a round trip through HBM'},
            'kernels': kernels,
            'cpu_baseline': {'value': n_cpu * N_F / cpu_s, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '%d volume pairs of the same workload (%.1f s), restated reference '
                                       '(oracle/torch_oracle.py, torch-CPU fp32, %d threads)' % (n_cpu, cpu_s, cores)},
            'clocks': clocks,
            'configs': configs,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


K_SS = 'ss_step(k_ss_march)'
K_RS = 'rescale_x2(k_upsample3_march)'
K_WP = 'warp_linear(k_warp_tex)'
K_FU = 'rescale_warp_fused(k_rescale_warp_tex)'
# dram__bytes_read.sum + dram__bytes_write.sum per launch at B=32, from the committed `ncu --set full` capture of
# `python bench.py --steps 2 --warmup 3 --batch 32 --no-e2e` (profiles/r2_b32_ncu_full_summary.csv; the SS entry is
# a halo-2 step).  Every kernel moves slightly LESS than its algorithmic bytes: part of the writes is still in the
# 126 MB L2 when the launch ends.
TRAFFIC_NCU_B32_GB = {
    K_SS: 0.25637 + 0.19567,
    K_RS: 0.25414 + 1.82949,
    K_WP: 2.51435 + 0.61650,
    K_FU: 0.86307 + 0.59855,
}


def run_configs(rank, world, dev):
    """BASELINE.json configs 3-5 through the scripts that measure them (same code as the stand-alone runs)."""
    import importlib.util
    import torch

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, 'scripts', name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    out = {}
    out['config3_train_step_tail'] = load('bench_train_tail').measure(rank, world, dev, items=2, steps=5, warmup=3)
    torch.cuda.empty_cache()
    out['config4_two_step_64_subjects'] = load('bench_two_step').measure(rank, world, dev, subjects=64, steps=3, warmup=2)
    torch.cuda.empty_cache()
    out['config5_jacobian_256'] = load('bench_jacobian').measure(rank, world, dev, fields=8, steps=5, warmup=3)
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=32, help='volume pairs per GPU per step')
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--no-e2e', action='store_true', help='skip the e2e and cpu_baseline legs (profiling runs)')
    ap.add_argument('--no-configs', action='store_true', help='skip BASELINE configs 3-5 (profiling runs)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_own(args)


if __name__ == '__main__':
    main()
