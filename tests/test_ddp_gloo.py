"""CPU, world_size 2, gloo: the host-side logic of the data-parallel path -- image sharding, the SUM all-reduce of the
depth-error tables (integer counters stay exact), and DistributedDataParallel construction over the drop-in module with
its dead parameters frozen.  (The kernels themselves need a GPU; their N>1 run is bench.py under torchrun.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _inputs as I


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import losses as OL
        import supervised_dispnet_b200 as S
        from supervised_dispnet_b200 import dist as D
        B = 6
        gt = I.sparse_gt(B, 64, 96, seed=300, dataset='kitti', density=0.2)
        pred = I.depth_map(B, 64, 96, seed=301, lo=0.0005, hi=95.0)
        lo, hi = D.shard_range(B, rank, world)
        g, p = gt[lo:hi], pred[lo:hi]
        cnt = OL.error_counters(g, p, 'kitti', True)
        # per-rank float sums through the oracle (stand-in for the CUDA kernel's table on a GPU box)
        y1, y2, x1, x2 = OL.garg_crop(64, 96)
        sums = np.zeros((hi - lo, 5))
        for b in range(hi - lo):
            valid = (g[b] > 0) & (g[b] < 80)
            cm = torch.zeros_like(valid); cm[y1:y2, x1:x2] = True
            valid &= cm
            vg, vp = g[b][valid], p[b][valid].clamp(1e-3, 80)
            d = vg - vp
            sums[b] = [float(d.abs().double().sum()), float((d.abs() / vg).double().sum()), float((d * d / vg).double().sum()),
                       float((d * d).double().sum()), float(((vg.log() - vp.log()) ** 2).double().sum())]
        errs, tot = D.allreduce_error_counters(cnt, sums)
        # DDP construction over the drop-in module: only trainable parameters take part
        net = S.models.DispNetS()
        ddp = D.wrap_ddp(net)
        n_ddp = sum(p.numel() for p in ddp.parameters() if p.requires_grad)
        vgg = S.models.Disp_vgg_BN()
        n_train = sum(p.numel() for p in vgg.parameters() if p.requires_grad)
        # flat data-parallel mode: attach() makes every rank start from rank 0's parameters and buffers
        net2 = S.models.DispNetS()
        torch.manual_seed(100 + rank)
        net2.init_weights()
        D.attach(net2)
        probe = float(net2.conv1[0].weight.double().sum())
        gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.tensor([probe], dtype=torch.float64))
        assert all(float(g) == float(gathered[0]) for g in gathered) and net2._dp_group is not None
        # gradient averaging semantics: mean of per-rank means == global mean for equal shards
        w = torch.nn.Parameter(torch.zeros(3))
        lin = torch.nn.parallel.DistributedDataParallel(torch.nn.Linear(3, 1, bias=False))
        x = torch.arange(12.).view(4, 3)[2 * rank:2 * rank + 2]
        lin(x).mean().backward()
        gavg = lin.module.weight.grad.clone()
        q.put((rank, errs, tot, n_ddp, n_train, gavg.tolist(), (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharding_and_metric_allreduce():
    from oracle import losses as OL
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    B = 6
    gt = I.sparse_gt(B, 64, 96, seed=300, dataset='kitti', density=0.2)
    pred = I.depth_map(B, 64, 96, seed=301, lo=0.0005, hi=95.0)
    ref = OL.compute_errors(gt, pred, 'kitti', True)
    ref_cnt = OL.error_counters(gt, pred, 'kitti', True).sum(0).tolist()
    assert res[0][6] == (0, 3) and res[1][6] == (3, 6)
    for rank, errs, tot, n_ddp, n_train, gavg, _ in res:
        assert tot == ref_cnt                                   # integer counters: exact after the SUM all-reduce
        for a, b in zip(errs, ref):
            assert a == pytest.approx(b, rel=1e-5)
        assert n_ddp == 31_596_900
        assert 1.9e7 < n_train < 2.1e7                          # 143.5 M total minus the frozen 123.6 M classifier
        assert gavg[0] == pytest.approx([4.5, 5.5, 6.5])         # global mean of x over the 4 rows
