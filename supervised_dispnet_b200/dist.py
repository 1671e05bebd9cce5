"""Multi-GPU plumbing: one process per GPU, batches sharded by image, weights replicated, one
gradient all-reduce per step through DistributedDataParallel over NCCL (NVLink 5 / NVSwitch).  The reference's own
multi-GPU path is single-process nn.DataParallel (train.py:316-317); per-rank BatchNorm statistics and rank-0 running
buffers (broadcast_buffers=True) reproduce its per-replica semantics."""
import numpy as np
import torch
import torch.distributed as dist


def wrap_ddp(net, device=None):
    """DistributedDataParallel over the trainable parameters only.  The dead VGG classifier (123.6 M parameters that never
    receive a gradient) is frozen by the model class, so the reducer does not wait for it (SURVEY.md 2.3 'DDP trap')."""
    ids = [device.index] if device is not None and device.type == 'cuda' else None
    return torch.nn.parallel.DistributedDataParallel(net, device_ids=ids, broadcast_buffers=True, gradient_as_bucket_view=True)


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items independent images for this rank (weak scaling uses n_items = b * world)."""
    per = (n_items + world - 1) // world
    return min(rank * per, n_items), min((rank + 1) * per, n_items)


def allreduce_error_counters(counters, sums, group=None):
    """Global compute_errors from per-rank (counters int64 [B,4], sums float64 [B,5]) tables: the per-sample metrics are
    formed locally, summed over all samples of all ranks with SUM all-reduces, then divided by the global batch.
    Integer counters stay exact."""
    n = counters[:, 0].astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        per = np.stack([sums[:, 0] / n, sums[:, 1] / n, sums[:, 2] / n, np.sqrt(sums[:, 3] / n), np.sqrt(sums[:, 4] / n),
                        counters[:, 1] / n, counters[:, 2] / n, counters[:, 3] / n], 1).sum(0)
    t = torch.tensor(list(per) + [float(counters.shape[0])], dtype=torch.float64)
    c = torch.tensor(counters.sum(0), dtype=torch.int64)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, group=group)
        dist.all_reduce(c, group=group)
    return [float(v) / float(t[8]) for v in t[:8]], c.tolist()


def attach(net, group=None):
    """Data parallelism without the DDP wrapper: broadcast rank 0's parameters and buffers once, then let the module's own
    backward all-reduce (average) its flat fp32 gradient arena with ONE NCCL call per step (engine.Plan.run_backward: three parameter ranges, each all-reduced while the ranges below it are still in their backward).
    The gradients already live in one contiguous buffer, so DDP's bucket copies, hooks and per-step buffer broadcasts buy
    nothing here; BatchNorm running statistics stay per rank and rank 0's are the ones checkpointed, which is what the
    reference's nn.DataParallel keeps (replica 0; train.py:316-317, 378)."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError('torch.distributed is not initialised')
    with torch.no_grad():
        for t in list(net.parameters()) + list(net.buffers()):
            dist.broadcast(t, src=0, group=group)
    net._dp_group = group if group is not None else dist.group.WORLD
    return net
