#!/usr/bin/env python
"""bench.py -- images/sec of the DispNet hot path (BASELINE.json configs[1]: Disp_vgg_BN + L1 depth loss, synthetic
KITTI-shaped 128x416 batches, b=32 per GPU), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: disp_net forward -> 1/disp -> l1_loss + smooth_loss ->
backward -> Adam (what train.train() does per batch, reference train.py:420-522).
  value : steps timed with inputs already resident in HBM (CUDA events, max over ranks)
  e2e   : the same metric through the public API `supervised_dispnet_b200.train.train()` with pinned HOST batches:
          H2D copy of every batch and the D2H `loss.item()` of every step are inside the timed region
  roofline : the tcgen05 gather-convolution kernel (forward + data-gradient launches), algorithmic FLOPs / device time
          measured with CUDA events around every launch in a separate pass of the same step
  cpu_baseline : the oracle port (oracle/nets.py + oracle/losses.py, torch CPU fp32) on this box's host cores
  parity   : the benchmarked step at b=32, 128x416 against the fp32 CPU oracle (disparities, loss, gradients) and the Abs Rel
          figure (compute_errors(...)[1], loss_functions.py:444) of an eval-mode validation batch, reference vs this path
  parity_mode : the same step in precision 'tc32' (fp32 storage, split-bf16 tcgen05 GEMMs: the tensor-core mode that meets the
          north star's 1e-3) benched beside the headline mode, with its own parity figures
  gpu_baseline : the reference's own modules (staged baseline/_ref, else the oracle port) on the same B200 through
          PyTorch/cuDNN in fp32, TF32 and bf16-autocast -- the kernel to beat
--impl reference times that CPU implementation alone (bounded sample per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch  # noqa: E402

H, W, BATCH = 128, 416, 32
TRAIN_GFLOP_PER_IMG = 110.7        # BASELINE.md section 3 (3 x forward MACs x 2), Disp_vgg_BN @ 128x416
METRIC = 'images/sec (128x416, b=32/GPU)'
# forward + data-gradient gather-convolutions of Disp_vgg_BN at b=32: every layer reads its input and writes its output once
# (2 B/element) in each direction: 2 x 2 x (1.14 G conv-input + 1.10 G conv-output elements) ~ 4.4 GB per step
ALGO_CONV_BYTES_PER_STEP = 4.4e9


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


def synth_batch(b, seed, pinned=False):
    """images U(0,1) normalised to [-1,1]; KITTI-like sparse depth, ~5% valid, >=1 valid px in the crop."""
    import _inputs as I
    x = I.images(b, H, W, seed)
    gt = I.sparse_gt(b, H, W, seed + 1, 'kitti', density=0.05)
    if pinned:
        x, gt = x.pin_memory(), gt.pin_memory()
    return x, gt


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled during the timed region: NVML every 10 ms (in-process, ~50 us per sample), falling
    back to `nvidia-smi --query-gpu` (the profiling recipe's clocks line, ~60 ms per sample) when NVML cannot be loaded."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    BITS = [0x8, 0x40, 0x20, 0x4]      # nvmlClocksEventReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.source = index, [], False, 'nvidia-smi'
        self.nv = self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid) if not uuid.startswith('GPU-') else uuid)
            except Exception:  # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv, self.source = pynvml, 'nvml'
        except Exception:  # noqa: BLE001
            self.nv = None

    def sample(self):
        try:
            if self.nv is not None:
                mhz = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((mhz, self.max_mhz, [bool(mask & b) for b in self.BITS]))
            else:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                f = [t.strip() for t in out.strip().split(',')]
                if len(f) >= 6:
                    self.samples.append((float(f[0]), float(f[1]), [t.lower().startswith('active') for t in f[2:6]]))
        except Exception:  # noqa: BLE001
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.01 if self.nv is not None else 0.05)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unsampled'])
        sm = sorted(t[0] for t in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(t[2][i] for t in self.samples)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.samples[0][1], reasons=reasons, samples=len(sm), source=self.source)


# ------------------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the same step
# ------------------------------------------------------------------------------------------------------------
def cpu_step_fn():
    from oracle import nets as ON, losses as OL
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    params = [v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k]
    opt = torch.optim.Adam(params, lr=2e-4, betas=(0.9, 0.999))

    def step(x, gt):
        disp = ON.disp_vgg_bn(sd, x, True)
        depth = [1 / d for d in disp]
        loss = 1.0 * OL.l1_loss(gt, depth, 'kitti') + 0.0 * OL.smooth_loss(depth)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss)
    return step


def cpu_baseline(budget_s=20.0):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_fn()
    x, gt = synth_batch(1, 100)
    t0 = time.time()
    step(x, gt)
    t1 = time.time() - t0                      # one image incl. first-touch costs
    b = int(max(1, min(BATCH, budget_s / max(t1, 1e-3) * 1.3)))
    x, gt = synth_batch(b, 101)
    t0 = time.time()
    step(x, gt)
    dt = time.time() - t0
    n = 1
    while dt < 0.5 * budget_s and n < 8:        # ~10-20 s of CPU work in total
        step(x, gt)
        n += 1
        dt = time.time() - t0
    return dict(value=n * b / dt, unit='images/sec', cores=torch.get_num_threads(), kind='port',
                sample='%d step(s) of the oracle port (torch CPU fp32: Disp_vgg_BN fwd + L1 + smooth + bwd + Adam) at b=%d, %dx%d, '
                       'after a b=1 warm-up step' % (n, b, H, W))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_fn()
    x, gt = synth_batch(1, 100)
    t0 = time.time()
    step(x, gt)
    t1 = time.time() - t0
    total = args.steps + args.warmup
    b = int(max(1, min(BATCH, 150.0 / total / max(t1, 1e-3) * 1.3)))
    x, gt = synth_batch(b, 101)
    for _ in range(args.warmup):
        step(x, gt)
    t0 = time.time()
    for _ in range(args.steps):
        step(x, gt)
    dt = time.time() - t0
    v = args.steps * b / dt
    line = dict(metric=METRIC, value=v, unit='images/sec', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000 * dt / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference',
                config=dict(workload='Disp_vgg_BN + L1 depth loss, synthetic KITTI 128x416', per_step_batch=b,
                            note='reference CPU path (oracle port: the reference is pure PyTorch, no native code to compile); '
                                 'each step is a bounded sample of b=%d images of the b=32 workload' % b),
                cpu_baseline=dict(value=v, unit='images/sec', cores=torch.get_num_threads(), kind='port',
                                  sample='%d steps at b=%d' % (args.steps, b)),
                e2e=dict(value=v, unit='images/sec', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def make_step(model, opt, LF):
    def step(x, gt, m=None):
        disp = (m or model)(x)
        depth = [1 / d for d in disp]
        loss = 1.0 * LF.l1_loss(gt, depth, 'kitti') + 0.0 * LF.smooth_loss(depth)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss
    return step


def gpu_parity(precision, dev):
    """The benchmarked step against the fp32 CPU oracle at the benchmark's own size (b=32, 128x416): L2-relative error of the
    four disparity maps, relative error of the loss scalar, global / worst-parameter gradient error; plus Abs Rel
    (compute_errors(...)[1]) of an eval-mode validation batch computed by the reference arithmetic and by this path."""
    import _parity as P
    import _inputs as I
    import supervised_dispnet_b200 as S
    from supervised_dispnet_b200 import loss_functions as LF
    from oracle import nets as ON, losses as OL
    t0 = time.time()
    r = P.config2_step_case(precision, B=BATCH, H=H, W=W)
    out = dict(precision=precision, batch=BATCH, disp_rel=max(r['disp']), disp_rel_per_scale=r['disp'], loss_rel=r['loss'],
               grad_global_rel=r['grad_global'], grad_worst_param_rel=r['grad_worst'], grad_worst_param=r['grad_worst_name'])
    # Abs Rel on a validation batch (validate_with_gt, train.py:642-723: eval-mode forward, depth = 1/disp, compute_errors)
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    xv, gv = I.images(8, H, W, seed=400), I.sparse_gt(8, H, W, seed=401, dataset='kitti', density=0.05)
    with torch.no_grad():
        d_ref = ON.disp_vgg_bn({k: v.clone() for k, v in sd.items()}, xv, False)
        e_ref = OL.compute_errors(gv, 1 / d_ref[:, 0], 'kitti', True)
        c_ref = OL.error_counters(gv, 1 / d_ref[:, 0], 'kitti', True)
        m = S.models.Disp_vgg_BN()
        m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
        m.precision = precision
        m.to(dev).eval()
        d_our = m(xv.to(dev))
        e_our = LF.compute_errors(gv.to(dev), 1 / d_our[:, 0], 'kitti', True)
        c_our, _ = LF.error_counters(gv.to(dev), (1 / d_ref[:, 0]).to(dev), 'kitti', True)     # identical (gt, pred): bit-exact ints
    out.update(abs_rel_ref=e_ref[1], abs_rel_ours=e_our[1], abs_rel_rel_diff=abs(e_our[1] - e_ref[1]) / abs(e_ref[1]),
               error_counters_bit_exact_on_identical_inputs=bool((torch.as_tensor(c_ref).cpu() == torch.as_tensor(c_our).cpu()).all()),
               eval_disp_rel=float((d_our.cpu().double() - d_ref.double()).norm() / d_ref.double().norm()),
               seconds=round(time.time() - t0, 1))
    del m
    torch.cuda.empty_cache()
    return out


def gpu_baseline(dev, steps=20, warmup=5):
    """The reference's own Disp_vgg_BN + l1_loss + smooth_loss step (unmodified modules from baseline/_ref when staged, else the
    oracle port) on this B200 through PyTorch/cuDNN, cudnn.benchmark as train.py sets it: fp32 (TF32 off), TF32 (PyTorch's
    default for cuDNN convolutions) and bf16 autocast.  Also each mode's disparity error against the fp32 CPU oracle."""
    import _inputs as I
    from oracle import nets as ON, losses as OL, refshim as R
    res = dict(source=None, modes={})
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    x_h, gt_h = synth_batch(BATCH, 10)
    x, gt = x_h.to(dev), gt_h.to(dev)
    xs = I.images(4, H, W, seed=200)
    with torch.no_grad():
        d_cpu = ON.disp_vgg_bn({k: v.clone() for k, v in sd.items()}, xs, True)
    root = R.find_root()
    ref = None
    if root is not None:
        try:
            ref = R.import_reference(root, with_train=False)
            res['source'] = 'unmodified reference modules (%s)' % os.path.relpath(root, ROOT) if root.startswith(ROOT) else root
        except Exception as e:  # noqa: BLE001
            res['import_error'] = repr(e)[:200]
    if ref is None:
        res['source'] = 'oracle port (functional restatement on torch ops)'
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for mode in ('fp32', 'tf32', 'bf16_autocast'):
            torch.backends.cudnn.allow_tf32 = mode != 'fp32'
            torch.backends.cuda.matmul.allow_tf32 = mode != 'fp32'
            cast = torch.autocast('cuda', dtype=torch.bfloat16, enabled=(mode == 'bf16_autocast'))
            if ref is not None:
                net = ref.models.Disp_vgg_BN()
                net.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
                net = net.to(dev).train()
                params = [p for p in net.parameters() if p.requires_grad]
                fwd = lambda t: net(t)                                        # noqa: E731
                l1, sm = ref.loss_functions.l1_loss, ref.loss_functions.smooth_loss
            else:
                sdd = {k: v.clone().to(dev) for k, v in sd.items()}
                params = [v.requires_grad_(True) for k, v in sdd.items() if v.dtype.is_floating_point and 'running' not in k]
                fwd = lambda t: ON.disp_vgg_bn(sdd, t, True)                  # noqa: E731
                l1, sm = OL.l1_loss, OL.smooth_loss
            opt = torch.optim.Adam(params, lr=2e-4, betas=(0.9, 0.999))
            with torch.no_grad(), cast:
                d = fwd(xs.to(dev))
            err = max(float((a.float().cpu().double() - b.double()).norm() / b.double().norm()) for a, b in zip(d, d_cpu))

            def step(full=True):
                with cast:
                    disp = fwd(x)
                    depth = [1 / t.float() for t in disp]
                    loss = (1.0 * l1(gt, depth, 'kitti') + 0.0 * sm(depth)) if full else sum(t.mean() for t in depth)
                opt.zero_grad()
                loss.backward()
                opt.step()
            out = {}
            for tag, full in (('step', True), ('net_only', False)):
                for _ in range(warmup):
                    step(full)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    step(full)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                out[tag + '_ms'] = ms
                out[tag + '_images_per_sec'] = BATCH / (ms * 1e-3)
            out['disp_rel_vs_cpu_fp32'] = err
            res['modes'][mode] = out
            del opt, params
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    res['note'] = ('step = forward + l1_loss + 0*smooth_loss + backward + Adam at b=%d (the reference loss code syncs the host per '
                   'sample, loss_functions.py:104-129); net_only = the same with a trivial loss; %d timed steps, CUDA events' % (BATCH, steps))
    return res


def unchanged_loop_e2e(net, opt, x_h, gt_h, steps, dev):
    """images/sec through the reference's OWN train.train (unmodified train.py:394-539 from baseline/_ref) with this package's
    model and loss_functions swapped in: blocking .to(device) copies, 3+ .item() syncs and a CSV append per step included."""
    import tempfile
    import supervised_dispnet_b200 as S
    from oracle import refshim as R
    root = R.find_root()
    if root is None:
        return dict(unavailable='reference checkout not staged (python -m oracle.refshim)')
    ref = R.import_reference(root)
    T = ref.train
    T.device, T.n_iter = dev, 0
    T.loss_functions = S.loss_functions
    with tempfile.TemporaryDirectory() as d:
        a = R.reference_args(d, batch_size=BATCH)
        T.train(a, [(x_h, gt_h)] * 4, net, torch.nn.Identity(), opt, 4, R.NullLogger(), R.NullWriter())
        torch.cuda.synchronize()
        t0 = time.time()
        T.train(a, [(x_h, gt_h)] * steps, net, torch.nn.Identity(), opt, steps, R.NullLogger(), R.NullWriter())
        torch.cuda.synchronize()
        dt = time.time() - t0
    return dict(value=BATCH * steps / dt, unit='images/sec', ms_per_step=1000 * dt / steps, steps=steps,
                loop='unmodified reference train.train (%s/train.py:394-539), models + loss_functions swapped' % os.path.basename(root))


def run_ours(args):
    import torch.distributed as dist
    import supervised_dispnet_b200 as S
    from supervised_dispnet_b200 import _lib as L, loss_functions as LF, train as T

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if L.lib().dn_tc_available() != 1:
        raise RuntimeError('tcgen05 path unavailable on this device; bench.py measures the sm_100a kernels only')
    precision = args.precision

    def build(prec):
        torch.manual_seed(0)
        net = S.models.Disp_vgg_BN('kitti')
        net.init_weights()
        net.precision = prec
        return net.to(dev).train()

    net = build(precision)
    model = net
    dp_mode = os.environ.get('DISPNET_B200_DP', 'flat')
    if world > 1:
        from supervised_dispnet_b200 import dist as D
        if dp_mode == 'ddp':       # torch DistributedDataParallel wrapper (bucketed all-reduce)
            model = D.wrap_ddp(net, dev)
        else:                      # NCCL all-reduce of the module's flat gradient arena, overlapped with the backward
            D.attach(net)
    params = [p for p in net.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=2e-4, betas=(0.9, 0.999), fused=True)

    x_h, gt_h = synth_batch(BATCH, 10 + rank, pinned=True)
    x_d, gt_d = x_h.to(dev), gt_h.to(dev)
    step = make_step(model, opt, LF)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(stp, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss = stp(x_d, gt_d)
        e1.record()
        return e0, e1, loss

    # set-up: the engine builds its plan on the 1st step and captures the forward / backward CUDA graphs on the 3rd / 4th; two
    # priming steps here keep those one-time costs out of both the warm-up count and the timed region for any W >= 3
    for _ in range(2):
        step(x_d, gt_d)
    # (N > 1: the first all-reduces of a fresh NCCL communicator still set up channels / NVLS buffers; with a short timed region
    # -- the driver's scaling run times ~20 steps = 0.17 s -- that start-up would be measured instead of the steady state)
    n_warm = max(args.warmup, 3) if world == 1 else max(args.warmup, 20 if world < 4 else 60)
    for _ in range(n_warm):
        step(x_d, gt_d)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    calls0 = L.CALLS
    e0, e1, loss = timed_steps(step, args.steps)
    sampler.sample()          # the device is still working through the enqueued steps here: at least one sample under load
    barrier()
    calls = L.CALLS - calls0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    last_loss = float(loss)
    if args.minimal:
        if rank == 0:
            print(json.dumps(dict(metric=METRIC, value=BATCH * world * args.steps / (ms * 1e-3), ms_per_step=ms / args.steps,
                                  minimal=True, precision=precision)), flush=True)
        return

    # ---- e2e: public API with host batches
    targs = T.default_args(batch_size=BATCH, smooth_loss_weight=0.0)

    class Loader:
        def __iter__(self):
            for _ in range(args.steps):
                yield x_h, gt_h

    def e2e_ms(mdl, optim):
        T.train(targs, [(x_h, gt_h)] * 4, mdl, None, optim, 4)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        T.train(targs, Loader(), mdl, None, optim, args.steps)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    ms2 = e2e_ms(model, opt)
    sampler.stop_flag = True

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline pass (rank 0): CUDA events around every kernel launch of the same step
    peaks = load_peaks()
    net._dp_group = None            # rank 0 is alone from here on: no collective in the profiling pass
    L.PROFILE = []
    nprof = 3
    for _ in range(nprof):
        step(x_d, gt_d, net)          # the bare module: the other ranks have left, no DDP collective here
    torch.cuda.synchronize()
    agg = {}
    for name, tag, a, b in L.PROFILE:
        key = name if tag is None else '%s[%s,%s]' % (name, tag[0], 'tc' if tag[1] else 'cuda-core')
        t, n, f = agg.get(key, (0.0, 0, 0.0))
        agg[key] = (t + a.elapsed_time(b), n + 1, f + (tag[2] if tag else 0.0))
    if args.dump:
        per = {}
        for name, tag, a, b in L.PROFILE:
            if tag is None:
                continue
            k = '%s %s %s' % (tag[3], tag[0], 'tc' if tag[1] else 'cc')
            t, f = per.get(k, (0.0, 0.0))
            per[k] = (t + a.elapsed_time(b) / nprof, f + tag[2] / nprof)
        with open(args.dump, 'w') as fh:
            for k, (t, f) in sorted(per.items(), key=lambda kv: -kv[1][0]):
                fh.write('%-44s %8.3f ms %9.2f GFLOP %8.1f TFLOP/s\n' % (k, t, f / 1e9, f / (t * 1e-3) / 1e12 if t > 0 else 0))
    L.PROFILE = None

    def cls(prefix):
        t = sum(t for k, (t, n, f) in agg.items() if k.startswith(prefix) and 'tc' in k) / nprof
        f = sum(f for k, (t, n, f) in agg.items() if k.startswith(prefix) and 'tc' in k) / nprof
        n = sum(n for k, (t, n, f) in agg.items() if k.startswith(prefix) and 'tc' in k) / nprof
        return t, f, n
    tc_t, tc_f, tc_n = cls('dn_igemm_run')
    wg_t, wg_f, wg_n = cls('dn_wgrad_run')
    kern_ms = sum(t for (t, n, f) in agg.values()) / nprof
    achieved = tc_f / (tc_t * 1e-3) / 1e12 if tc_t > 0 else 0.0
    wg_achieved = wg_f / (wg_t * 1e-3) / 1e12 if wg_t > 0 else 0.0
    breakdown = {k: round(t / nprof, 3) for k, (t, n, f) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]}

    traffic, traffic_src = None, None
    for tp in ('r2_traffic.json', 'r1_traffic.json'):
        tp = os.path.join(ROOT, 'profiles', tp)
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic, traffic_src = tj['dram_bytes_per_launch'], tj['source']
            break
    total_imgs = BATCH * world * args.steps
    value = total_imgs / (ms * 1e-3)
    e2e = total_imgs / (ms2 * 1e-3)
    dtype_txt = {'mixed': 'fp16 operands, fp32 accumulate, bf16 gradient activations',
                 'tc32': 'fp32 storage; split-bf16 (3-term) tcgen05 operands, fp32 accumulate'}.get(precision, precision)
    line = dict(metric=METRIC, value=value, unit='images/sec', n_gpus=world, steps=args.steps, warmup=n_warm,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='%s (precision=%s)' % (dtype_txt, precision), data='synthetic',
                config=dict(workload='configs[1]: Disp_vgg_BN + L1 depth loss (+0*smooth as train.py does), synthetic KITTI '
                                     '128x416, b=32/GPU, fwd+loss+bwd+Adam', global_batch=BATCH * world, precision=precision,
                            parallelism='dp%d' % world, dp_mode=(dp_mode if world > 1 else None),
                            l2='per-step working set (activations+gradients ~4 GB) >> 126 MB L2; no explicit flush needed',
                            last_loss=last_loss),
                clocks=sampler.summary(),
                e2e=dict(value=e2e, unit='images/sec', h2d_bytes_per_step=x_h.numel() * 4 + gt_h.numel() * 4,
                         d2h_bytes_per_step=4, ms_per_step=ms2 / args.steps,
                         api='supervised_dispnet_b200.train.train() (pipelined mirror of train.py:394-539)'),
                gpu_launches=calls,
                roofline=dict(bound='tensor', kernel='igemm_tc_kernel (forward + data-gradient gather-convolutions)',
                              achieved=achieved, peak=peaks['tf_sust'], unit='TFLOP/s', frac=achieved / peaks['tf_sust'],
                              traffic=traffic, traffic_unit='DRAM bytes per launch (ncu dram__bytes_read+write, average over the launches of one step)',
                              traffic_source=traffic_src, algorithmic_bytes_per_launch=ALGO_CONV_BYTES_PER_STEP / max(tc_n, 1),
                              peak_source=peaks['src'] + ' sustained bf16 (kernel timed inside a long step)',
                              launches_per_step=tc_n, kernel_ms_per_step=tc_t, flops_per_step=tc_f,
                              wgrad=dict(kernel='wgrad_tc_kernel (weight-gradient GEMMs)', achieved=wg_achieved,
                                         frac=wg_achieved / peaks['tf_sust'], launches_per_step=wg_n, kernel_ms_per_step=wg_t,
                                         flops_per_step=wg_f),
                              step_fraction_of_tensor_roofline=(value / world * TRAIN_GFLOP_PER_IMG) / (peaks['tf_sust'] * 1e3)),
                kernel_breakdown_ms_per_step=breakdown, all_kernels_ms_per_step=kern_ms)
    if world == 1 and not args.no_extras:
        def guarded(fn, *a):
            try:
                return fn(*a)
            except Exception as e:  # noqa: BLE001
                import traceback
                traceback.print_exc()
                return dict(error=repr(e)[:300])
        line['e2e_unchanged_loop'] = guarded(unchanged_loop_e2e, net, opt, x_h, gt_h, args.steps, dev)
        line['parity'] = guarded(gpu_parity, precision, dev)
        other = 'tc32' if precision != 'tc32' else 'mixed'
        del step, model, opt, params
        net._plans = {}
        del net
        torch.cuda.empty_cache()

        def second_mode():
            net2 = build(other)
            opt2 = torch.optim.Adam([p for p in net2.parameters() if p.requires_grad], lr=2e-4, betas=(0.9, 0.999), fused=True)
            step2 = make_step(net2, opt2, LF)
            for _ in range(2 + n_warm):
                step2(x_d, gt_d)
            torch.cuda.synchronize()
            a, b, _ = timed_steps(step2, args.steps)
            torch.cuda.synchronize()
            m1 = a.elapsed_time(b)
            m2 = e2e_ms(net2, opt2)
            r = dict(precision=other, value=total_imgs / (m1 * 1e-3), unit='images/sec', ms_per_step=m1 / args.steps,
                     e2e=dict(value=total_imgs / (m2 * 1e-3), unit='images/sec', ms_per_step=m2 / args.steps),
                     step_fraction_of_tensor_roofline=(total_imgs / (m1 * 1e-3) * TRAIN_GFLOP_PER_IMG) / (peaks['tf_sust'] * 1e3))
            net2._plans = {}
            return r
        pm = guarded(second_mode)
        torch.cuda.empty_cache()
        if 'error' not in pm:
            pm['parity'] = guarded(gpu_parity, other, dev)
        line['parity_mode' if other == 'tc32' else 'fast_mode'] = pm
        line['gpu_baseline'] = guarded(gpu_baseline, dev)
    if world == 1 and not args.no_cpu:
        line['cpu_baseline'] = cpu_baseline()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip parity / second precision mode / cuDNN baseline (N=1 only)')
    ap.add_argument('--precision', default=os.environ.get('DISPNET_B200_PRECISION', 'mixed'),
                    help="headline precision: 'mixed' (fp16 operands) or 'tc32' (fp32-class split-bf16 tensor-core mode)")
    ap.add_argument('--dump', default=None, help='write per-layer conv kernel timings of the roofline pass here')
    ap.add_argument('--minimal', action='store_true', help='warm-up + timed steps only (for runs under ncu)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
