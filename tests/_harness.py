"""Shared helpers for the GPU parity tests: single-layer modules on the product engine, error metrics, and the
oracle-vs-product comparisons used by tests/test_gpu_*.py, tools/gpu_check.py and __graft_entry__.smoke()."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

import supervised_dispnet_b200 as S
from supervised_dispnet_b200 import engine as E
from supervised_dispnet_b200 import _lib as L


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def maxabs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


class OneConv(E.PlannedModule):
    """A single nn.Conv2d / nn.ConvTranspose2d (+ fused activation) on the engine, output returned as NCHW."""

    def __init__(self, cin, cout, k, stride=1, pad=None, transposed=False, out_pad=0, act=L.ACT_NONE, bias=True,
                 crop=None, precision='fp32'):
        super().__init__()
        self.precision = precision
        pad = (k - 1) // 2 if pad is None else pad
        self.cfg = (cin, cout, k, stride, pad, transposed, out_pad, act, bias, crop)
        if transposed:
            self.conv = nn.ConvTranspose2d(cin, cout, k, stride, pad, out_pad, bias=bias)
        else:
            self.conv = nn.Conv2d(cin, cout, k, stride, pad, bias=bias)

    def out_hw(self, H, W):
        cin, cout, k, stride, pad, transposed, out_pad, act, bias, crop = self.cfg
        if transposed:
            ho, wo = (H - 1) * stride - 2 * pad + k + out_pad, (W - 1) * stride - 2 * pad + k + out_pad
        else:
            ho, wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        if crop:
            ho, wo = min(ho, crop[0]), min(wo, crop[1])
        return ho, wo

    def _build_plan(self, plan, shapes):
        cin, cout, k, stride, pad, transposed, out_pad, act, bias, crop = self.cfg
        N, _, H, W = shapes[0]
        # the input goes through a 1x1 identity-free path: InputOp then the conv under test (needs_dx=True so the
        # data gradient is produced into the input buffer's gradient arena, read back by the test)
        inp = plan.add(E.InputOp(plan, shapes))
        ho, wo = self.out_hw(H, W)
        out = plan.new_buf(N, ho, wo, cout).view()
        self.op = plan.add(E.ConvOp(plan, 'conv', inp.out, out, k, stride=stride, pad=pad, transposed=transposed, bias=bias,
                                    act=act, needs_dx=True))
        self.tail = plan.add(DumpOp(plan, out))
        self.inp = inp

    def forward(self, x):
        return self._run([x])[0]

    def input_grad(self, x):
        plan = self._plan_for([x])
        return self.inp.out.grad_view(plan.prec.grad).to_nchw() / plan.prec.gscale

    def backends(self, x):
        plan = self._plan_for([x])
        return dict(fwd=[b[1] for b in self.op._fwd_built], wgrad=[b[1] for b in self.op.wg], dgrad=[b[1] for b in self.op.dg])


class DumpOp(E.Op):
    """Test-only op: exposes an NHWC view as an fp32 NCHW output and feeds the incoming gradient back into the
    view's gradient arena (scaled like the heads do)."""

    def __init__(self, plan, v):
        self.v = v
        self.idx = plan.add_output((v.N, v.C, v.H, v.W))

    def fwd(self, plan):
        plan.outputs[self.idx].copy_(self.v.to_nchw())

    def plan_bwd(self, plan):
        self.gv = self.v.grad_view(plan.prec.grad)
        assert not self.gv.claim_grad_write()

    def bwd(self, plan):
        g = plan.gouts[self.idx] * plan.prec.gscale
        buf = self.gv.buf.t
        buf[..., self.gv.c0:self.gv.c0 + self.gv.C] = g.permute(0, 2, 3, 1).to(buf.dtype)


def torch_conv_ref(m, x, gout):
    """fp32 torch reference of OneConv on the same device: returns out, dx, dw, db."""
    cin, cout, k, stride, pad, transposed, out_pad, act, bias, crop = m.cfg
    x = x.detach().clone().requires_grad_(True)
    w = m.conv.weight.detach().clone().requires_grad_(True)
    b = m.conv.bias.detach().clone().requires_grad_(True) if bias else None
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        if transposed:
            y = F.conv_transpose2d(x, w, b, stride, pad, out_pad)
        else:
            y = F.conv2d(x, w, b, stride, pad)
        if crop:
            y = y[:, :, :crop[0], :crop[1]]
        if act == L.ACT_RELU:
            y = F.relu(y)
        elif act == L.ACT_LRELU:
            y = F.leaky_relu(y, 0.1)
        (y * gout).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    return y.detach(), x.grad, w.grad, (b.grad if bias else None)


def conv_case(cfg, shape, precision='fp32', seed=0, device='cuda'):
    """Runs one OneConv config on the engine and against torch; returns dict of relative errors + backends used."""
    torch.manual_seed(seed)
    m = OneConv(precision=precision, **cfg).to(device)
    m.train()
    x = torch.randn(shape, device=device)
    xr = x.clone().requires_grad_(False)
    out = m(xr)
    gout = torch.randn_like(out)
    (out * gout).sum().backward()
    dx = m.input_grad(xr)
    y, rdx, rdw, rdb = torch_conv_ref(m, x, gout)
    res = dict(fwd=rel(out, y), dx=rel(dx, rdx), dw=rel(m.conv.weight.grad, rdw))
    if rdb is not None:
        res['db'] = rel(m.conv.bias.grad, rdb)
    res['backends'] = m.backends(xr)
    return res
