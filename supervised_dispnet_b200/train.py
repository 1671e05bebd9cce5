"""Host-side mirror of the reference's training loop `train.train(args, train_loader, disp_net, pose_exp_net,
optimizer, epoch_size, logger, train_writer)` (reference train.py:394-539) for the hot-path configurations.

Same signature, argument meaning and per-step call order as the reference (H2D copy -> disp_net -> 1/disp ->
loss_functions.* -> zero_grad / backward / step -> loss.item()), so that a reference checkout can also simply do
`import supervised_dispnet_b200.models as models; import supervised_dispnet_b200.loss_functions as loss_functions`
and keep its own train.py.  Differences, all forced by breakages documented in SURVEY.md 3.3:
  * a batch may be the reference's 2-tuple `(tgt_img, gt_depth)` (supervised) or the 5-tuple
    `(tgt_img, ref_imgs, intrinsics, intrinsics_inv, gt_depth)` the reference's commented-out line :418 used
    (the checked-in unsupervised branch reads unbound names);
  * every `--loss` choice of train.py:449-470 except DORN (which needs the DORN network) is wired;
  * tensorboard / csv side effects happen only when `train_writer` / `args.save_path` are given;
  * the `.to(device)` copies of batch i+1 (train.py:424-432) are issued on a copy stream while step i computes
    (`_DevicePrefetcher`): same tensors, same order, the 27 MB H2D copy just no longer sits between two steps;
  * the per-step `loss.item()` (:517) is an asynchronous 4-byte copy into pinned memory that is consumed two iterations
    later (`_LossReader`): every step's loss is still read back and averaged, but the host no longer drains the GPU in
    the middle of every step (between forward and backward).
"""
import csv
import os
import time
from types import SimpleNamespace

import torch

from . import loss_functions

n_iter = 0


class AverageMeter(object):
    """Running average of one or several values (reference logger.py:62-89)."""

    def __init__(self, i=1, precision=3):
        self.meters, self.precision = i, precision
        self.reset(i)

    def reset(self, i):
        self.val, self.avg, self.sum, self.count = [0] * i, [0] * i, [0] * i, 0

    def update(self, val, n=1):
        if not isinstance(val, list):
            val = [val]
        assert len(val) == self.meters
        self.count += n
        for i, v in enumerate(val):
            self.val[i] = v
            self.sum[i] += v * n
            self.avg[i] = self.sum[i] / self.count

    def __repr__(self):
        val = ' '.join(['{:.{}f}'.format(v, self.precision) for v in self.val])
        avg = ' '.join(['{:.{}f}'.format(a, self.precision) for a in self.avg])
        return '{} ({})'.format(val, avg)


class _DevicePrefetcher(object):
    """Iterates `loader`, handing out batches whose tensors already live on `device`.  The host->device copies of the
    NEXT batch are enqueued on a private copy stream the moment the current batch is handed out, so they overlap the
    current step's kernels; the compute stream waits on the copy's event before it touches the batch.  The device side
    is a ring of four preallocated buffer sets (no allocator traffic in the loop): the copy into a slot waits for the
    event that marks the end of the enqueued work of the batch that used the slot four hand-outs earlier, so a handed
    out batch stays valid until three more batches have been requested.  At most `limit` batches are pulled from the
    loader (the reference's loop breaks after `epoch_size` batches, train.py:536)."""
    SLOTS = 4          # _LossReader.DEPTH + 2: the host runs at most DEPTH steps ahead of the device

    _STATE = {}        # device -> (copy stream, ring buffers, slot events): kept across calls - a new stream per epoch would
                       # strand the previous ring in another stream's allocator pool and pay cudaMalloc / cudaFree again

    def __init__(self, loader, device, limit):
        self.it, self.device, self.limit, self.pulled = iter(loader), torch.device(device), limit, 0
        self.cuda = self.device.type == 'cuda'
        if self.cuda:
            key = self.device.index if self.device.index is not None else torch.cuda.current_device()
            st = self._STATE.get(key)
            if st is None:
                st = (torch.cuda.Stream(self.device), [dict() for _ in range(self.SLOTS)], [None] * self.SLOTS)
                self._STATE[key] = st
            # slot -> {position in the batch: device tensor};  slot -> event: the work enqueued on its batch has finished
            self.stream, self.bufs, self.done = st
        else:
            self.stream, self.bufs, self.done = None, [dict() for _ in range(self.SLOTS)], [None] * self.SLOTS
        self.handed = None                                       # slot of the batch the caller is working on
        self._next = None
        self._preload()

    def _move(self, obj, slot, path):
        if torch.is_tensor(obj):
            if not self.cuda or obj.device == self.device:
                return obj.to(self.device)
            buf = self.bufs[slot].get(path)
            if buf is None or buf.shape != obj.shape or buf.dtype != obj.dtype:
                buf = torch.empty(obj.shape, dtype=obj.dtype, device=self.device)
                self.bufs[slot][path] = buf
            buf.copy_(obj, non_blocking=True)
            return buf
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._move(o, slot, path + (k,)) for k, o in enumerate(obj))
        return obj

    def _preload(self):
        self._next = None
        if self.pulled >= self.limit:
            return
        try:
            batch = next(self.it)
        except StopIteration:
            return
        slot = self.pulled % self.SLOTS
        self.pulled += 1
        if not self.cuda:
            self._next = (self._move(batch, slot, ()), None, slot)
            return
        with torch.cuda.stream(self.stream):
            if self.done[slot] is not None:
                self.stream.wait_event(self.done[slot])          # the slot's previous batch is no longer read by any kernel
            b = self._move(batch, slot, ())
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self._next = (b, ev, slot)

    def __iter__(self):
        return self

    def __next__(self):
        if self.cuda and self.handed is not None:                # everything the caller enqueued on the previous batch
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.done[self.handed] = ev
        if self._next is None:
            raise StopIteration
        batch, ev, slot = self._next
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        self.handed = slot
        self._preload()
        return batch


_PINNED_SCALARS = []      # two pinned 4-byte landing buffers, allocated once per process (cudaHostAlloc costs milliseconds)


class _LossReader(object):
    """`losses.update(loss.item(), n)` with the device->host read deferred by DEPTH steps: the 4-byte copy is enqueued
    right where the reference calls `.item()`, its value is consumed DEPTH iterations later, so the host stays up to
    DEPTH steps ahead of the device and a few milliseconds of host jitter never drain the GPU queue."""
    DEPTH = 2

    def __init__(self, meter, device):
        self.meter, self.cuda = meter, torch.device(device).type == 'cuda'
        if self.cuda and len(_PINNED_SCALARS) < self.DEPTH + 1:
            _PINNED_SCALARS.extend(torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(self.DEPTH + 1 - len(_PINNED_SCALARS)))
        self.pending, self.k = [], 0

    def push(self, loss, n):
        if not self.cuda:
            self.meter.update(loss.item(), n)
            return
        buf = _PINNED_SCALARS[self.k]
        self.k = (self.k + 1) % (self.DEPTH + 1)
        buf.copy_(loss.detach(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.pending.append((buf, ev, n))
        while len(self.pending) > self.DEPTH:
            self._pop()

    def _pop(self):
        buf, ev, n = self.pending.pop(0)
        ev.synchronize()
        self.meter.update(buf.item(), n)

    def flush(self):
        while self.pending:
            self._pop()


def default_args(**kw):
    """argparse defaults of the reference that the loop reads (train.py:28-91)."""
    a = dict(photo_loss_weight=1.0, mask_loss_weight=0.0, smooth_loss_weight=0.0, unsupervised=False, dataset='kitti',
             loss='L1', monodepth2=False, diff_lr=False, rotation_mode='euler', padding_mode='zeros', print_freq=10,
             training_output_freq=0, batch_size=4, network='disp_vgg_BN', save_path=None, log_full='progress_log_full.csv')
    a.update(kw)
    return SimpleNamespace(**a)


def train(args, train_loader, disp_net, pose_exp_net, optimizer, epoch_size, logger=None, train_writer=None,
          device=None):
    global n_iter
    device = device or next(disp_net.parameters()).device
    batch_time, data_time, losses = AverageMeter(), AverageMeter(), AverageMeter(precision=4)
    w1, w2, w3 = args.photo_loss_weight, args.mask_loss_weight, args.smooth_loss_weight
    disp_net.train()
    if pose_exp_net is not None:
        pose_exp_net.train()
    if getattr(args, 'diff_lr', False):
        for m in disp_net.modules():
            if m.__class__.__name__.find('BatchNorm') != -1:
                m.eval()
    end = time.time()
    if logger is not None:
        logger.train_bar.update(0)

    reader = _LossReader(losses, device)
    for i, batch in enumerate(_DevicePrefetcher(train_loader, device, epoch_size)):
        data_time.update(time.time() - end)
        if len(batch) == 2:
            tgt_img, gt_depth = batch
            ref_imgs = intrinsics = intrinsics_inv = None
        else:
            tgt_img, ref_imgs, intrinsics, intrinsics_inv, gt_depth = batch
        tgt_img = tgt_img.to(device, non_blocking=True)

        if args.unsupervised:
            ref_imgs = [img.to(device, non_blocking=True) for img in ref_imgs]
            intrinsics = intrinsics.to(device, non_blocking=True)
            intrinsics_inv = intrinsics_inv.to(device, non_blocking=True)
            explainability_mask, pose = pose_exp_net(tgt_img, ref_imgs)

        if gt_depth is not None:
            gt_depth = gt_depth.to(device, non_blocking=True)
            if args.dataset == 'nyu' and gt_depth.dim() == 4:
                gt_depth = torch.squeeze(gt_depth[:, 0, :, :])

        disparities = disp_net(tgt_img)
        scale = 5.4 if getattr(args, 'monodepth2', False) else 1
        depth = [scale / disp for disp in disparities]

        if not args.unsupervised:           # the `--loss` switch of train.py:449-470 (DORN needs its own network: out of scope)
            if args.loss == 'Multi_L1':
                loss_1 = loss_functions.Multiscale_L1_loss(gt_depth, depth)
            elif args.loss == 'Multi_full_L1':
                loss_1 = loss_functions.Multiscale_FULL_L1_loss(gt_depth, depth)
            elif args.loss == 'Multi_berhu':
                loss_1 = loss_functions.Multiscale_berhu_loss(gt_depth, depth)
            elif args.loss == 'Multi_L2':
                loss_1 = loss_functions.Multiscale_L2_loss(gt_depth, depth)
            elif args.loss == 'L1':
                loss_1 = loss_functions.l1_loss(gt_depth, depth, args.dataset)
            elif args.loss == 'berhu':
                loss_1 = loss_functions.berhu_loss(gt_depth, depth, args.dataset)
            elif args.loss == 'L2':
                loss_1 = loss_functions.l2_loss(gt_depth, depth, args.dataset)
            elif args.loss == 'scale_inv':
                loss_1 = loss_functions.Scale_invariant_loss(gt_depth, depth, args.dataset)
            elif args.loss == 'Multi_scale_inv':
                loss_1 = loss_functions.Multiscale_scale_inv_loss(gt_depth, depth)
            else:
                raise TypeError('undefined loss')      # the reference does `raise "undefined loss"` (a TypeError in py3)
        else:
            loss_1 = loss_functions.photometric_reconstruction_loss(tgt_img, ref_imgs, intrinsics, intrinsics_inv, depth,
                                                                    explainability_mask, pose, args.rotation_mode,
                                                                    args.padding_mode)
        loss_2 = loss_functions.explainability_loss(explainability_mask) if w2 > 0 else 0
        loss_3 = loss_functions.smooth_loss(depth)
        loss = w1 * loss_1 + w2 * loss_2 + w3 * loss_3

        if train_writer is not None and i > 0 and n_iter % args.print_freq == 0:
            train_writer.add_scalar('photometric_error', loss_1.item(), n_iter)
            if w2 > 0:
                train_writer.add_scalar('explanability_loss', loss_2.item(), n_iter)
            train_writer.add_scalar('disparity_smoothness_loss', loss_3.item(), n_iter)
            train_writer.add_scalar('total_loss', loss.item(), n_iter)

        reader.push(loss, args.batch_size)      # the per-step D2H read of the reference (:517), consumed two steps later

        optimizer.zero_grad()
        loss.backward()
        optimizer.step()

        batch_time.update(time.time() - end)
        end = time.time()
        if getattr(args, 'save_path', None):
            with open(os.path.join(str(args.save_path), args.log_full), 'a') as csvfile:
                csv.writer(csvfile, delimiter='\t').writerow([loss.item(), loss_1.item(), loss_2.item() if w2 > 0 else 0,
                                                              loss_3.item()])
        if logger is not None:
            logger.train_bar.update(i + 1)
            if i % args.print_freq == 0:
                logger.train_writer.write('Train: Time {} Data {} Loss {}'.format(batch_time, data_time, losses))
        if i >= epoch_size - 1:
            break
        n_iter += 1
    reader.flush()
    return losses.avg[0]


def upsample_bilinear_ac(x, size):
    """nn.UpsamplingBilinear2d(size=size) (align_corners=True) of [B,h,w] on the device kernel (train.py:696-700)."""
    from . import _lib as L
    L.require_cuda(x)
    x = x.contiguous().float()
    B, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    out = torch.empty((B, H, W), dtype=torch.float32, device=x.device)
    L.call('dn_resize_bilinear_ac', L.ptr(x), B, h, w, H, W, L.ptr(out), L.stream_ptr())
    return out


@torch.no_grad()
def validate_with_gt(args, val_loader, disp_net, epoch, logger=None, output_writers=[], device=None):
    """Mirror of the reference's validate_with_gt (train.py:642-723): eval-mode forward (BatchNorm folded into the
    convolutions by the engine), depth = 1/disp (x5.4 under --monodepth2), the NYU branch's UpsamplingBilinear2d to the
    ground truth's size, loss_functions.compute_errors per batch, AverageMeter over batches.  Returns (errors.avg, names)."""
    device = device or next(disp_net.parameters()).device
    batch_time = AverageMeter()
    error_names = ['abs_diff', 'abs_rel', 'sq_rel', 'rmse', 'rmse_log', 'a1', 'a2', 'a3']
    errors = AverageMeter(i=len(error_names))
    disp_net.eval()
    end = time.time()
    if logger is not None:
        logger.valid_bar.update(0)
    n = 0
    for i, (tgt_img, depth) in enumerate(val_loader):
        tgt_img = tgt_img.to(device)
        depth = depth.to(device)
        if args.dataset == 'nyu':
            depth = torch.squeeze(depth[:, 0, :, :])
        if getattr(args, 'loss', 'L1') == 'DORN':
            raise TypeError('DORN is outside the accelerated path')
        output_disp = disp_net(tgt_img)
        output_depth = 1 / output_disp[:, 0]
        if getattr(args, 'monodepth2', False):
            output_depth = output_depth * 5.4
        if args.dataset == 'nyu':
            d3 = depth if depth.dim() == 3 else depth.unsqueeze(0)
            output_depth = torch.squeeze(upsample_bilinear_ac(output_depth, d3.shape[1:]))
            depth = d3 if output_depth.dim() == 3 else depth
            if output_depth.dim() == 2:
                output_depth, depth = output_depth.unsqueeze(0), d3
        errors.update(loss_functions.compute_errors(depth, output_depth, dataset=args.dataset,
                                                    unsupervised=getattr(args, 'unsupervised', False)))
        batch_time.update(time.time() - end)
        end = time.time()
        n = i + 1
        if logger is not None:
            logger.valid_bar.update(i + 1)
            if i % args.print_freq == 0:
                logger.valid_writer.write('valid: Time {} Abs Error {:.4f} ({:.4f})'.format(batch_time, errors.val[0], errors.avg[0]))
    if logger is not None:
        logger.valid_bar.update(n)
    return errors.avg, error_names
