"""Host-side layer engine: static per-shape execution plans over NHWC views, driving the C-ABI kernels.

This is tensor plumbing only -- every arithmetic step is a call into libdispnet_b200.so.  A `Plan` is built
once per (network, input shape, mode); it owns the activation / gradient arenas in HBM (sized for the
180 GB part: nothing is re-materialised), the packed-weight workspaces and the ordered op list.  Concatenation
(`torch.cat` in the reference decoders) never happens: producers write straight into channel slices of the
consumer's buffer, transposed convolutions write 2x2 phase sub-lattices, crops are views.

Gradient flow mirrors the reference's autograd graph: ops run in reverse order; the first writer of a gradient
region overwrites, later writers accumulate (decided statically at plan-build time).
"""
import ctypes as C
import os

import torch

from . import _lib as L

_DT = {torch.float32: L.DN_F32, torch.float16: L.DN_F16, torch.bfloat16: L.DN_BF16}


def _ru(x, m):
    return (x + m - 1) // m * m


class Precision:
    """Storage / compute types of a plan.  'fp32': CUDA-core kernels, fp32 everywhere (exact-parity mode).
    'tc32': fp32 storage, split-bf16 (3-term) tcgen05 GEMMs -- the parity mode on the tensor cores.
    'fp16': fp16 activations and weights, gradient activations in `grad` dtype (scaled by `gscale` when
    fp16), fp32 accumulation everywhere, tcgen05 kernels where the problem shape allows."""

    def __init__(self, name):
        self.name = name
        self.split = None
        if name == 'tc32':
            # fp32 storage and fp32-class arithmetic ON the tensor cores: every GEMM operand is split into two bf16 terms
            # (hi = bf16(x), lo = bf16(x - hi); x = hi + lo to 2^-18) and a convolution runs as the three tcgen05 terms
            # hi*w_hi + lo*w_hi + hi*w_lo with fp32 accumulation in TMEM (the dropped lo*w_lo term is 2^-18 relative).
            # This is the mode that meets the reference's fp32 results to 1e-3 / 1e-5 (models/Disp_vgg_BN.py:136-191).
            self.act, self.grad, self.gscale, self.split = torch.float32, torch.float32, 1.0, torch.bfloat16
        elif name == 'fp32':
            self.act, self.grad, self.gscale = torch.float32, torch.float32, 1.0
        elif name == 'fp16':
            self.act, self.grad, self.gscale = torch.float16, torch.float16, 4096.0
        elif name in ('mixed', 'fp16_bf16grad'):
            # fp16 forward (11-bit significand keeps disparities within ~1e-3 of fp32), bf16 gradient activations
            # (fp32 exponent range: no loss scaling).  tcgen05 kind::f16 cannot mix the two operand formats, so the
            # weight-gradient GEMM reads a bf16 copy of the saved activation made just in time (ConvOp.bwd).
            self.act, self.grad, self.gscale = torch.float16, torch.bfloat16, 1.0
        elif name == 'bf16':
            self.act, self.grad, self.gscale = torch.bfloat16, torch.bfloat16, 1.0
        else:
            raise ValueError('unknown precision %r' % name)


def default_precision():
    return os.environ.get('DISPNET_B200_PRECISION', 'mixed')


def graphs_enabled():
    return os.environ.get('DISPNET_B200_GRAPHS', '1') != '0' and L.PROFILE is None


def tc_enabled():
    return os.environ.get('DISPNET_B200_BACKEND', 'auto') != 'generic' and L.lib().dn_tc_available() == 1


class Buf:
    """NHWC storage with channel pitch rounded to 8 elements (16-byte pixel alignment for TMA / vector access)."""

    def __init__(self, N, H, W, Cc, dtype, device, zero=True, storage=None):
        self.N, self.H, self.W, self.C = N, H, W, Cc
        # channel pitch: whole TMA rows for the odd channel counts (the 3-channel image, the 17- / 97- / 193- / 385-channel
        # concatenation buffers): 16 / 32 / 64-channel rows below 64 channels, multiples of 64 above; everything else is
        # already a multiple of 8.  The padding is zero-filled once and never written.
        self.Cp = _pitch(Cc)
        self.readable_pad = storage is None and zero
        self.dtype = dtype
        if storage is not None:     # carve from a shared scratch tensor
            self.t = storage[:N * H * W * self.Cp * torch.empty((), dtype=dtype).element_size()].view(dtype).view(N, H, W, self.Cp)
        else:
            f = torch.zeros if zero else torch.empty
            self.t = f((N, H, W, self.Cp), dtype=dtype, device=device)
        self.grad = None
        self.written = []      # channel intervals of .grad already written during the backward walk (plan time)

    def view(self):
        return View(self, 0, self.C, self.H, self.W, 0, self.W * self.Cp, self.Cp)

    def split_planes(self, dtype):
        """(hi, lo) 16-bit planes of an fp32 buffer (precision 'tc32'); same N/H/W/C, hence the same element strides."""
        if getattr(self, 'planes', None) is None:
            self.planes = (Buf(self.N, self.H, self.W, self.C, dtype, self.t.device), Buf(self.N, self.H, self.W, self.C, dtype, self.t.device))
            self.split_done = []       # channel intervals already split earlier in the forward op order (plan time)
        return self.planes

    def grad_buf(self, gdtype):
        if self.grad is None:
            self.grad = Buf(self.N, self.H, self.W, self.C, gdtype, self.t.device)
        return self.grad


class View:
    """Channel slice / phase sub-lattice / crop of a Buf (element strides; channel stride 1)."""

    def __init__(self, buf, c0, Cc, H, W, off, sH, sW):
        self.buf, self.c0, self.C, self.H, self.W = buf, c0, Cc, H, W
        self.off, self.sH, self.sW = off, sH, sW
        self.N = buf.N
        self.sN = buf.H * buf.W * buf.Cp
        self._dn = None

    def channels(self, c0, Cc):
        assert c0 + Cc <= self.C
        return View(self.buf, self.c0 + c0, Cc, self.H, self.W, self.off, self.sH, self.sW)

    def crop(self, H, W):
        assert H <= self.H and W <= self.W
        return View(self.buf, self.c0, self.C, H, W, self.off, self.sH, self.sW)

    def phase(self, a, b):
        H = (self.H - a + 1) // 2
        W = (self.W - b + 1) // 2
        return View(self.buf, self.c0, self.C, H, W, self.off + a * self.sH + b * self.sW, 2 * self.sH, 2 * self.sW)

    def phase_h(self, a):
        """every second row starting at row a (all columns)"""
        return View(self.buf, self.c0, self.C, (self.H - a + 1) // 2, self.W, self.off + a * self.sH, 2 * self.sH, self.sW)

    def on(self, buf):
        """The same region of another buffer with identical geometry (a gradient, shadow or split plane)."""
        return View(buf, self.c0, self.C, self.H, self.W, self.off, self.sH, self.sW)

    def planes(self, dtype):
        hi, lo = self.buf.split_planes(dtype)
        return self.on(hi), self.on(lo)

    def claim_split(self, dtype):
        """Plan time: True when this region's (hi, lo) planes are not yet produced by an earlier op of the forward order
        (the caller then launches dn_split_bf16 for the whole view before its GEMM)."""
        self.buf.split_planes(dtype)
        lo, hi = self.c0, self.c0 + self.C
        if any(a <= lo and hi <= b for a, b in self.buf.split_done):
            return False
        self.buf.split_done.append((lo, hi))
        return True

    def dn(self):
        if self._dn is None:
            es = self.buf.t.element_size()
            # c_ext: what follows the view's channels inside the pixel record is zero padding of the buffer or other slices of a
            # concatenation buffer -- finite values a gather-convolution may read against zero weight rows (dn_view.c_ext);
            # scratch-backed buffers are not zero-initialised, so they declare nothing beyond C
            c_ext = (self.buf.Cp - self.c0) if (self.buf.readable_pad and self.sW % self.buf.Cp == 0) else self.C
            self._dn = L.DnView(self.buf.t.data_ptr() + (self.off + self.c0) * es, _DT[self.buf.dtype], self.N, self.H,
                                self.W, self.C, c_ext, self.sN, self.sH, self.sW)
        return self._dn

    def ref(self):
        return C.byref(self.dn())

    def grad_view(self, gdtype):
        g = self.buf.grad_buf(gdtype)
        return View(g, self.c0, self.C, self.H, self.W, self.off, self.sH, self.sW)

    # ---- plan-time gradient bookkeeping: returns True if a write to this region must accumulate
    def claim_grad_write(self):
        lo, hi = self.c0, self.c0 + self.C
        w = self.buf.written
        covered = any(a <= lo and hi <= b for a, b in w)
        partial = any(a < hi and lo < b for a, b in w)
        if not covered:
            if partial:
                raise RuntimeError('partially overlapping gradient writes are not supported')
            w.append((lo, hi))
        return covered

    def to_nchw(self):
        """Debug helper: NCHW fp32 copy of the view."""
        t = self.buf.t.view(-1)
        idx_n = torch.arange(self.N, device=t.device).view(-1, 1, 1, 1) * self.sN
        idx_h = torch.arange(self.H, device=t.device).view(1, -1, 1, 1) * self.sH
        idx_w = torch.arange(self.W, device=t.device).view(1, 1, -1, 1) * self.sW
        idx_c = torch.arange(self.C, device=t.device).view(1, 1, 1, -1)
        return t[idx_n + idx_h + idx_w + idx_c + self.off + self.c0].permute(0, 3, 1, 2).float().contiguous()


def _pitch(Cc):
    """Channel pitch of a buffer with Cc channels (see Buf)."""
    if Cc % 8 != 0 and os.environ.get('DISPNET_B200_PAD_PITCH', '1') != '0':
        return 16 if Cc <= 16 else 32 if Cc <= 32 else _ru(Cc, 64)
    return _ru(Cc, 8)


def _i32arr(vals):
    return (C.c_int32 * len(vals))(*vals)


def _backend(kind, prob):
    if not tc_enabled():
        return 0
    fn = L.lib().dn_igemm_tc_supported if kind == 'igemm' else L.lib().dn_wgrad_tc_supported
    return 1 if fn(C.byref(prob)) == 1 else 0


def _pad_ok(v):
    """True when the channels between v.C and the next multiple of 8 are padding of v's buffer."""
    return int(v.C % 8 != 0 and v.c0 % 8 == 0 and v.c0 + v.C == v.buf.C)


def _mk_igemm(ins, out, w, w_dtype, cin_pad, cout_pad, bias, act, accumulate, stride, taps, out_scale=1.0, out2=None):
    p = L.DnIgemm()
    for i, v in enumerate(ins):
        p.inp[i] = v.dn()
    p.nsrc = len(ins)
    p.out = out.dn()
    p.w = w.data_ptr()
    p.w_dtype = _DT[w_dtype]
    p.cin_pad, p.cout_pad = cin_pad, cout_pad
    p.bias = bias
    p.act, p.accumulate, p.stride = act, int(accumulate), stride
    p.ntaps = len(taps)
    for i, (s, dh, dw, wt) in enumerate(taps):
        p.taps[i] = L.DnTap(s, dh, dw, wt)
    p.out_scale = out_scale
    p.out_pad_ok = _pad_ok(out)
    if out2 is not None:
        d2 = out2.dn()
        p.out2, p.out2_dtype = d2.ptr, d2.dtype
    return p


def _mk_wgrad(ps, q, dw, cp_pad, cq_pad, stride, taps, scale):
    p = L.DnWgrad()
    for i, v in enumerate(ps):
        p.p[i] = v.dn()
    p.nsrc = len(ps)
    p.q = q.dn()
    p.dw = dw.data_ptr()
    p.cp_pad, p.cq_pad, p.stride = cp_pad, cq_pad, stride
    p.ntaps = len(taps)
    for i, (s, dh, dw_, wt) in enumerate(taps):
        p.taps[i] = L.DnTap(s, dh, dw_, wt)
    p.scale = scale
    return p


def _igemm_flops(p):
    """Algorithmic FLOPs of a gather-convolution: 2 * output pixels * Cout * (taps * Cin), true (unpadded) sizes."""
    return 2.0 * p.out.N * p.out.H * p.out.W * p.out.C * p.ntaps * p.inp[0].C


def _wgrad_flops(p):
    return 2.0 * p.p[0].N * p.p[0].H * p.p[0].W * p.p[0].C * p.q.C * p.ntaps


def _split3_igemm(ins, taps, T, dtype):
    """Split-precision form of a gather-convolution (precision 'tc32'): every source view becomes its (hi, lo) planes
    and every tap the three terms  hi * w_hi + lo * w_hi + hi * w_lo ; the packed weights hold the T hi matrices followed by
    the T lo matrices."""
    ins3 = []
    for v in ins:
        ins3 += list(v.planes(dtype))
    taps3 = []
    for (s_, dh, dw, wt) in taps:
        taps3 += [(2 * s_, dh, dw, wt), (2 * s_ + 1, dh, dw, wt), (2 * s_, dh, dw, T + wt)]
    return ins3, taps3


def _full_extent(v):
    """All pixels of v's buffer for v's channel slice (what dn_split_bf16 converts: phases and crops of it are views)."""
    b = v.buf
    return View(b, v.c0, v.C, b.H, b.W, 0, b.W * b.Cp, b.Cp)


class Op:
    def fwd(self, plan):
        raise NotImplementedError

    def plan_bwd(self, plan):
        pass

    def bwd(self, plan):
        pass


class InputOp(Op):
    """NCHW fp32 input tensor(s) -> channel slices of one NHWC buffer (replaces torch.cat of PoseExpNet inputs)."""

    def __init__(self, plan, shapes):
        N, _, H, W = shapes[0]
        self.chans = [s[1] for s in shapes]
        self.buf = Buf(N, H, W, sum(self.chans), plan.prec.act, plan.device)
        self.buf.is_input = True
        self.out = self.buf.view()
        self.shadow = plan.shadow_of(self.out) if plan.training else None

    def fwd(self, plan):
        c0 = 0
        for x, c in zip(plan.inputs, self.chans):
            L.call('dn_pack_input', L.ptr(x), x.shape[0], c, x.shape[2], x.shape[3], self.out.ref(), c0, plan.stream)
            if self.shadow is not None and getattr(self.buf, 'shadow_needed', False):    # (a row-expanded first conv has its own copy)
                L.call('dn_pack_input', L.ptr(x), x.shape[0], c, x.shape[2], x.shape[3], self.shadow.ref(), c0, plan.stream)
            c0 += c


class ConvOp(Op):
    """nn.Conv2d (any k / stride / pad) or nn.ConvTranspose2d (stride 2), with fused bias + activation."""

    def __init__(self, plan, name, x, out, k, stride=1, pad=None, transposed=False, bias=True, act=L.ACT_NONE,
                 needs_dx=True, bn_follows=False, fold_bn=None):
        self.name, self.x, self.out = name, x, out
        # eval mode (validate_with_gt, train.py:642-723): the BatchNorm behind this convolution normalises with its running
        # statistics, so conv + BN is one affine map -- gamma / sqrt(var + eps) goes into the packed weights (row_scale), the
        # rest into the epilogue's bias vector, and the separate normalise pass over the activation disappears
        self.fold_bn = fold_bn
        if fold_bn is not None:
            assert not plan.training and not transposed
            self.fold_ss = torch.zeros(2 * out.C, dtype=torch.float32, device=plan.device)
            self.fold_mi = torch.zeros(2 * out.C, dtype=torch.float32, device=plan.device)
            self.fold_bias = torch.zeros(out.C, dtype=torch.float32, device=plan.device)
            for suffix in ('.weight', '.bias'):
                plan.register_param(fold_bn + suffix)
        # a bias in front of BatchNorm has an analytically zero gradient (BN subtracts the batch mean); the reference
        # computes rounding noise around 0 there.  We write exact zeros and skip the extra pass over dy.
        self.bias_grad_zero = bn_follows
        self.k, self.stride, self.transposed, self.act = k, stride, transposed, act
        self.pad = (k - 1) // 2 if pad is None else pad
        self.needs_dx = needs_dx
        self.Cin, self.Cout = x.C, out.C
        self.has_bias = bias
        self.cin_pad, self.cout_pad = _ru(self.Cin, 64), _ru(self.Cout, 16)
        self.cinT_pad, self.coutT_pad = _ru(self.Cin, 16), _ru(self.Cout, 64)   # dgrad roles swapped
        T = k * k
        dev = plan.device
        # first layer (input = the image, no data gradient): kernel columns folded into the channel dimension, k taps
        # over k * Cin channels of a row-expanded copy of the image instead of k * k taps over 3 real channels each
        self.rowx = (not transposed and not needs_dx and stride in (1, 2) and k > 1 and k * x.C <= 128 and tc_enabled()
                     and (plan.prec.act != torch.float32 or plan.prec.split is not None) and getattr(x.buf, 'is_input', False) and x.c0 == 0
                     and x.C == x.buf.C and os.environ.get('DISPNET_B200_ROWX', '1') != '0')
        if getattr(x.buf, 'is_input', False) and not self.rowx:
            x.buf.shadow_needed = True          # this layer's weight gradient reads the gradient-dtype image of the input
        if self.rowx:
            cx = k * self.Cin
            self.xr = Buf(x.N, x.H, out.W, cx, plan.prec.act, dev).view()
            self.xr_g = None
            if plan.training and plan.prec.grad != plan.prec.act:
                self.xr_g = Buf(x.N, x.H, out.W, cx, plan.prec.grad, dev).view()
            self.cin_pad = _ru(cx, 64)
            if plan.prec.split is not None:
                self.wp = torch.zeros((2 * k, self.cout_pad, self.cin_pad), dtype=plan.prec.split, device=dev)     # [hi | lo]
            else:
                self.wp = torch.zeros((k, self.cout_pad, self.cin_pad), dtype=plan.prec.act, device=dev)
        elif plan.prec.split is not None:
            self.wp = torch.zeros((2 * T, self.cout_pad, self.cin_pad), dtype=plan.prec.split, device=dev)     # [hi | lo]
        else:
            self.wp = torch.zeros((T, self.cout_pad, self.cin_pad), dtype=plan.prec.act, device=dev)
        # split-precision operands: this op converts its input to (hi, lo) planes unless an earlier consumer already did
        self.split_x = None
        if plan.prec.split is not None:
            xs = _full_extent(self.xr if self.rowx else x)
            self.split_x = xs if xs.claim_split(plan.prec.split) else False
        self.kh = _i32arr([t // k for t in range(T)])
        self.kw = _i32arr([t % k for t in range(T)])
        # source strides of the torch parameter seen as [co][ci][kh][kw]
        if transposed:      # nn.ConvTranspose2d weight is [Cin, Cout, k, k]
            self.s_co, self.s_ci = T, self.Cout * T
        else:
            self.s_co, self.s_ci = self.Cin * T, T
        # ---- forward problems
        self.fwd_probs = []
        if self.rowx:
            if stride == 1:
                ins, taps = [self.xr], [(0, kh - self.pad, 0, kh) for kh in range(k)]
            else:       # input row 2*ho + kh - pad lives in row phase (kh - pad) & 1 at row ho + ((kh - pad) >> 1)
                ins = [self.xr.phase_h(0), self.xr.phase_h(1)]
                taps = [((kh - self.pad) & 1, (kh - self.pad) >> 1, 0, kh) for kh in range(k)]
            self.fwd_probs.append(dict(ins=ins, out=out, taps=taps, stride=1))
        elif not transposed and stride == 2 and x.H >= 2 and x.W >= 2:      # (odd sizes: the odd phases are one row / column shorter,
            # what a tap reads beyond them is outside the image and zero-filled like any padding)
            # input pixel (2*ho + dh, 2*wo + dw) lives in phase (dh & 1, dw & 1) of x at (ho + (dh >> 1), wo + (dw >> 1)):
            # a strided convolution is a stride-1 gather over the four 2x2 phase views (which the tcgen05 kernel serves)
            ins = [x.phase(a, b) for a in range(2) for b in range(2)]
            taps = [(((kh - self.pad) & 1) * 2 + ((kw - self.pad) & 1), (kh - self.pad) >> 1, (kw - self.pad) >> 1, kh * k + kw)
                    for kh in range(k) for kw in range(k)]
            self.fwd_probs.append(dict(ins=ins, out=out, taps=taps, stride=1))
        elif not transposed:
            taps = [(0, kh - self.pad, kw - self.pad, kh * k + kw) for kh in range(k) for kw in range(k)]
            self.fwd_probs.append(dict(ins=[x], out=out, taps=taps, stride=stride))
        else:
            assert stride == 2
            for a in range(2):
                for b in range(2):
                    ov = out.phase(a, b)
                    if ov.H == 0 or ov.W == 0:
                        continue
                    taps = [(0, (a + self.pad - kh) // 2, (b + self.pad - kw) // 2, kh * k + kw)
                            for kh in range(k) if (a + self.pad - kh) % 2 == 0
                            for kw in range(k) if (b + self.pad - kw) % 2 == 0]
                    self.fwd_probs.append(dict(ins=[x], out=ov, taps=taps, stride=1, phase=(a, b)))
        self._fwd_built = None
        # activations that a later convolution consumes get a gradient-dtype shadow written by the same epilogue
        self.out_shadow = plan.shadow_of(out) if (plan.training and act != L.ACT_NONE and out.buf.dtype != torch.float32) else None
        plan.register_param(name + '.weight')
        if bias:
            plan.register_param(name + '.bias')

    def _build_fwd(self, plan):
        built = []
        merged = self._build_fwd_stacked(plan) or self._build_fwd_merged(plan)
        if merged is not None:
            return merged
        for pr in self.fwd_probs:
            o2 = None
            if self.out_shadow is not None:
                o2 = self.out_shadow.phase(*pr['phase']) if 'phase' in pr else self.out_shadow
            ins, taps, wdt, div = pr['ins'], pr['taps'], plan.prec.act, 1.0
            if plan.prec.split is not None:
                ins, taps = _split3_igemm(ins, taps, self.k if self.rowx else self.k * self.k, plan.prec.split)
                wdt, div = plan.prec.split, 3.0
            p = _mk_igemm(ins, pr['out'], self.wp, wdt, self.cin_pad, self.cout_pad, None, self.act, False,
                          pr['stride'], taps, out2=o2)
            built.append((p, _backend('igemm', p), _igemm_flops(p) / div))
        return built

    def _build_fwd_stacked(self, plan):
        """Thin transposed convolution (4x4 / stride 2, at most 64 input channels, 4 * Cout_pad <= 128): the four output phases stacked
        along the output channels of ONE 3x3 gather-convolution over the input (dn_igemm.phase_cout).  Every input tile is fetched
        once and nine taps x N = 4 * Cout MMAs replace sixteen taps x N = Cout ones; the weights are packed as nine
        [4 * Cout_pad][Cin_pad] matrices that are zero where a phase does not use the tap.  None when not applicable."""
        prs = self.fwd_probs
        cpp = self.cout_pad
        if not (self.transposed and len(prs) == 4 and tc_enabled() and plan.prec.split is None and self.x.C <= 64 and 4 * cpp <= 128
                and plan.prec.act != torch.float32 and os.environ.get('DISPNET_B200_STACK_PHASES', '1') != '0'):
            return None
        if len({(pr['out'].H, pr['out'].W) for pr in prs}) != 1 or (prs[0]['out'].H, prs[0]['out'].W) != (self.x.H, self.x.W):
            return None
        union = sorted({(dh, dw) for pr in prs for (_, dh, dw, _) in pr['taps']})
        if any(abs(dh) > 1 or abs(dw) > 1 for dh, dw in union) or len(union) != 9:
            return None
        if getattr(self, 'wp9', None) is None:
            self.wp9 = torch.zeros((9, 4 * cpp, self.cin_pad), dtype=plan.prec.act, device=plan.device)
        base = prs[0]['out']
        taps = [(0, dh, dw, (dh + 1) * 3 + (dw + 1)) for dh, dw in union]
        o2 = self.out_shadow.phase(*prs[0]['phase']) if self.out_shadow is not None else None
        p = _mk_igemm([self.x], base, self.wp9, plan.prec.act, self.cin_pad, 4 * cpp, None, self.act, False, 1, taps, out2=o2)
        p.nphase, p.phase_cout = 4, cpp
        for i, pr in enumerate(prs):
            p.phase_off[i] = pr['out'].off - base.off
        if _backend('igemm', p) != 1:
            return None
        # one pack job per (phase, tap): the k x k parameter plane (kh, kw) goes to rows [phase * cpp, ...) of matrix (dh, dw)
        self.stack_jobs = []
        W = plan.param(self.name + '.weight')
        for i, pr in enumerate(prs):
            for (_, dh, dw, wt) in pr['taps']:
                # (source offset: (kh * k + kw) floats into the parameter; destination offset in bytes)
                self.stack_jobs.append((wt * 4, (((dh + 1) * 3 + (dw + 1)) * 4 * cpp + i * cpp) * self.cin_pad * self.wp9.element_size()))
        flops = sum(_igemm_flops(_mk_igemm([self.x], pr['out'], self.wp, plan.prec.act, self.cin_pad, self.cout_pad, None, self.act, False, 1,
                                           pr['taps'])) for pr in prs)
        return [(p, 1, flops)]

    def _build_fwd_merged(self, plan):
        """The output phases of a transposed convolution as ONE gather-convolution launch (dn_igemm.nphase): same input view and
        weights, the taps ordered phase by phase, every phase writing its 2x2 sub-lattice of `out` through an element offset.  Four
        13-tile launches of upconv4 become one 104-tile launch; None when the phases differ in extent or the tensor-core kernel
        does not take the problem."""
        prs = self.fwd_probs
        if not (self.transposed and len(prs) > 1 and tc_enabled() and os.environ.get('DISPNET_B200_MERGE_PHASES', '1') != '0'):
            return None
        if len({(pr['out'].H, pr['out'].W, len(pr['taps'])) for pr in prs}) != 1:
            return None
        base = prs[0]['out']
        ins, taps, wdt, div = prs[0]['ins'], [t for pr in prs for t in pr['taps']], plan.prec.act, 1.0
        if plan.prec.split is not None:
            ins, taps = _split3_igemm(ins, taps, self.k * self.k, plan.prec.split)
            wdt, div = plan.prec.split, 3.0
        if len(taps) > L.MAX_TAPS:
            return None
        o2 = self.out_shadow.phase(*prs[0]['phase']) if self.out_shadow is not None else None
        p = _mk_igemm(ins, base, self.wp, wdt, self.cin_pad, self.cout_pad, None, self.act, False, 1, taps, out2=o2)
        p.nphase = len(prs)
        for i, pr in enumerate(prs):
            p.phase_off[i] = pr['out'].off - base.off
        if _backend('igemm', p) != 1:
            return None
        flops = _igemm_flops(p)       # (ntaps counts every phase's taps, out = one phase's pixels)
        return [(p, 1, flops / div)]

    # ---- weight (un)packing is batched over all layers of the plan: one dn_pack_jobs launch each (Plan._run_jobs)
    def jobs(self, plan, which):
        if which == 'fwd' and self.transposed:
            if self._fwd_built is None:      # (decides whether the forward runs channel-stacked: different packed weights)
                self._fwd_built = self._build_fwd(plan)
            if getattr(self, 'stack_jobs', None):
                W = plan.param(self.name + '.weight')
                out = []
                for src_off, dst_off in self.stack_jobs:
                    j = L.DnPackJob()
                    j.T, j.k, j.s_kh, j.s_kw = 1, self.k, self.k, 1
                    j.src, j.dst = W.data_ptr() + src_off, self.wp9.data_ptr() + dst_off
                    j.dst_dtype, j.unpack = _DT[plan.prec.act], 0
                    j.R, j.Cc, j.R_pad, j.C_pad, j.s_r, j.s_c = self.Cout, self.Cin, self.cout_pad, self.cin_pad, self.s_co, self.s_ci
                    out.append(j)
                return out
        j = self.job(plan, which)
        if j is None:
            return []
        if plan.prec.split is None or which == 'unpack':
            return [j]
        # split precision: the same pack once more for the residual matrices (planes T .. 2T-1 of the packed tensor)
        j2 = self.job(plan, which)
        j.dst_dtype, j2.dst_dtype = L.DN_BF16, L.DN_BF16_LO
        T = self.k * self.k
        j2.dst = j.dst + T * j.R_pad * j.C_pad * 2
        return [j, j2]

    def job(self, plan, which):
        if self.rowx:
            return None          # packs / unpacks its weights itself (different column layout)
        T = self.k * self.k
        j = L.DnPackJob()
        j.T, j.k, j.s_kh, j.s_kw = T, self.k, self.k, 1
        W = plan.param(self.name + '.weight')
        if which == 'fwd':
            j.src, j.dst, j.dst_dtype, j.unpack = W.data_ptr(), self.wp.data_ptr(), _DT[plan.prec.split or plan.prec.act], 0
            j.R, j.Cc, j.R_pad, j.C_pad, j.s_r, j.s_c = self.Cout, self.Cin, self.cout_pad, self.cin_pad, self.s_co, self.s_ci
            if self.fold_bn is not None:
                j.row_scale = self.fold_ss.data_ptr()
        elif which == 'dgrad':
            if not self.needs_dx:
                return None
            j.src, j.dst, j.dst_dtype, j.unpack = W.data_ptr(), self.wpT.data_ptr(), _DT[plan.prec.split or plan.prec.grad], 0
            j.R, j.Cc, j.R_pad, j.C_pad, j.s_r, j.s_c = self.Cin, self.Cout, self.cinT_pad, self.coutT_pad, self.s_ci, self.s_co
        else:
            j.src, j.dst, j.unpack = self.dwp.data_ptr(), plan.grad_of(self.name + '.weight').data_ptr(), 1
            j.R, j.Cc, j.R_pad, j.C_pad, j.s_r, j.s_c = self.Cout, self.Cin, self.cout_pad, self.cin_pad, self.s_co, self.s_ci
            j.scale = 1.0 / plan.prec.gscale
        return j

    def pre_fwd(self, plan):
        """Runs before the batched weight pack of the step: the folded BatchNorm's scale / bias vectors (eval mode)."""
        if self.fold_bn is None:
            return
        n = self.fold_bn
        L.call('dn_bn_finalize', None, 1.0, L.ptr(plan.param(n + '.weight')), L.ptr(plan.param(n + '.bias')),
               L.ptr(plan.buffer(n + '.running_mean')), L.ptr(plan.buffer(n + '.running_var')), 0.1, 1e-5, 0, 0,
               L.ptr(self.fold_mi), L.ptr(self.fold_ss), self.Cout, plan.stream)
        cb = plan.param(self.name + '.bias') if self.has_bias else None
        L.call('dn_bn_fold_bias', L.ptr(cb), L.ptr(self.fold_ss), self.Cout, L.ptr(self.fold_bias), plan.stream)

    def fwd(self, plan):
        if self._fwd_built is None:
            self._fwd_built = self._build_fwd(plan)
        if self.rowx:
            L.call('dn_rowx_expand', self.x.ref(), self.k, self.stride, self.pad, self.xr.ref(),
                   self.xr_g.ref() if self.xr_g is not None else None, plan.stream)
            rs = L.ptr(self.fold_ss) if self.fold_bn is not None else None
            if plan.prec.split is not None:         # hi matrices, then the residual matrices
                for half, code in ((0, L.DN_BF16), (1, L.DN_BF16_LO)):
                    L.call('dn_rowx_pack_weight', L.ptr(plan.param(self.name + '.weight')), self.Cout, self.Cin, self.k,
                           L.ptr(self.wp[half * self.k:]), code, self.cout_pad, self.cin_pad, rs, plan.stream)
            else:
                L.call('dn_rowx_pack_weight', L.ptr(plan.param(self.name + '.weight')), self.Cout, self.Cin, self.k, L.ptr(self.wp),
                       _DT[plan.prec.act], self.cout_pad, self.cin_pad, rs, plan.stream)
        if self.split_x:
            hi, lo = self.split_x.planes(plan.prec.split)
            L.call('dn_split_bf16', self.split_x.ref(), hi.ref(), lo.ref(), plan.stream)
        b = plan.param(self.name + '.bias') if self.has_bias else None
        if self.fold_bn is not None:
            b = self.fold_bias
        for p, be, fl in self._fwd_built:
            p.bias = b.data_ptr() if b is not None else None
        plan.run_group([(p, be, ('fwd', be, fl, self.name)) for p, be, fl in self._fwd_built])

    def plan_bwd(self, plan):
        g = plan.prec.grad
        sp = plan.prec.split
        T = self.k * self.k
        dev = plan.device
        self.gout = self.out.grad_view(g)
        if self.rowx:
            T = self.k
        self.dwp = plan.dwp_alloc(T * self.cout_pad * self.cin_pad).view(T, self.cout_pad, self.cin_pad)
        k, pad = self.k, self.pad
        # ---- weight-gradient problems
        def wg_probs(q, gout=None, div=1.0):
            gout = self.gout if gout is None else gout
            probs = []
            if self.rowx:
                if self.stride == 1:
                    probs.append(_mk_wgrad([gout], q, self.dwp, self.cout_pad, self.cin_pad, 1,
                                           [(0, kh - pad, 0, kh) for kh in range(k)], 1.0))
                else:
                    for a in range(2):
                        taps = [(0, (kh - pad) >> 1, 0, kh) for kh in range(k) if ((kh - pad) & 1) == a]
                        if taps:
                            probs.append(_mk_wgrad([gout], q.phase_h(a), self.dwp, self.cout_pad, self.cin_pad, 1, taps, 1.0))
            elif not self.transposed and self.stride == 2 and q.H >= 2 and q.W >= 2:
                for a in range(2):          # one problem per input phase (see the forward tables)
                    for b in range(2):
                        taps = [(0, (kh - pad) >> 1, (kw - pad) >> 1, kh * k + kw) for kh in range(k) for kw in range(k)
                                if ((kh - pad) & 1) == a and ((kw - pad) & 1) == b]
                        if taps:
                            probs.append(_mk_wgrad([gout], q.phase(a, b), self.dwp, self.cout_pad, self.cin_pad, 1, taps,
                                                   1.0))
            elif not self.transposed:
                taps = [(0, kh - pad, kw - pad, kh * k + kw) for kh in range(k) for kw in range(k)]
                probs.append(_mk_wgrad([gout], q, self.dwp, self.cout_pad, self.cin_pad, self.stride, taps, 1.0))
            else:       # transposed: the output phases are the dy sources of ONE problem, taps ordered phase by phase
                srcs, taps = [], []
                for pr in self.fwd_probs:
                    a, b = pr['phase']
                    srcs.append(gout.phase(a, b))
                    taps += [(len(srcs) - 1, dh, dw, wt) for (_, dh, dw, wt) in pr['taps']]
                if len({(v.H, v.W) for v in srcs}) == 1 and len(taps) <= L.MAX_TAPS:
                    probs.append(_mk_wgrad(srcs, q, self.dwp, self.cout_pad, self.cin_pad, 1, taps, 1.0))
                else:   # (odd output sizes: the phases differ in extent)
                    for v, pr in zip(srcs, self.fwd_probs):
                        probs.append(_mk_wgrad([v], q, self.dwp, self.cout_pad, self.cin_pad, 1, pr['taps'], 1.0))
            return [(p, _backend('wgrad', p), _wgrad_flops(p) / div) for p in probs]

        self.xq = None
        self.gsplit = None
        if sp is not None:
            # split precision: dw = g_hi^T x_hi + g_lo^T x_hi + g_hi^T x_lo, three launches that add into the same fp32 dw
            # (the weight-gradient kernel accumulates with red.global.add); x's planes are the ones the forward made
            self.gsplit = _full_extent(self.gout)
            gh, gl = self.gout.planes(sp)
            xh, xl = (self.xr if self.rowx else self.x).planes(sp)
            self.wg = wg_probs(xh, gh, 3.0) + wg_probs(xh, gl, 3.0) + wg_probs(xl, gh, 3.0)
        else:
            self.wg = wg_probs(self.xr_g if (self.rowx and self.xr_g is not None) else (self.xr if self.rowx else self.x))
        if sp is None and not self.rowx and plan.prec.act != g and plan.prec.act != torch.float32 and tc_enabled():
            # kind::f16 MMAs need both operands in one format: use a just-in-time copy of x in the gradient dtype
            sh = plan.shadow_lookup(self.x)
            xq = sh if sh is not None else plan.scratch_buf(self.x.N, self.x.H, self.x.W, self.x.C, g).view()
            alt = wg_probs(xq)
            if all(be == 1 for _, be, _ in alt):
                self.wg = alt
                self.xq = None if sh is not None else xq     # a shadow is already filled by its producer
                if self.xq is not None:
                    # copy whole 16-byte channel groups (vector path) when the tail channels are buffer padding
                    cpad = _ru(self.x.C, 8) if _pad_ok(self.x) else self.x.C
                    self.x_cp = View(self.x.buf, self.x.c0, cpad, self.x.H, self.x.W, self.x.off, self.x.sH, self.x.sW)
                    self.xq_cp = View(xq.buf, xq.c0, cpad, xq.H, xq.W, xq.off, xq.sH, xq.sW)
        # ---- data-gradient problems
        self.dg = []
        if self.needs_dx:
            self.wpT = torch.zeros(((2 if sp is not None else 1) * T, self.cinT_pad, self.coutT_pad), dtype=sp or g, device=dev)
            gx = self.x.grad_view(g)
            acc = gx.claim_grad_write()
            self.dx_zero_first = None
            if not self.transposed and self.stride == 1:
                taps = [(0, pad - kh, pad - kw, kh * k + kw) for kh in range(k) for kw in range(k)]
                probs = [dict(ins=[self.gout], out=gx, taps=taps)]
            elif not self.transposed:
                assert self.stride == 2
                probs = []
                for a in range(2):
                    for b in range(2):
                        gv = gx.phase(a, b)
                        if gv.H == 0 or gv.W == 0:
                            continue
                        taps = [(0, (a + pad - kh) // 2, (b + pad - kw) // 2, kh * k + kw)
                                for kh in range(k) if (a + pad - kh) % 2 == 0
                                for kw in range(k) if (b + pad - kw) % 2 == 0]
                        if taps:
                            probs.append(dict(ins=[self.gout], out=gv, taps=taps))
                        elif not acc:
                            self.dx_zero_first = gx
            else:
                ins, taps = [], []
                for pr in self.fwd_probs:
                    a, b = pr['phase']
                    ins.append(self.gout.phase(a, b))
                    taps += [(len(ins) - 1, -dh, -dw, wt) for (_, dh, dw, wt) in pr['taps']]
                probs = [dict(ins=ins, out=gx, taps=taps)]
            for pr in probs:
                ins, taps, wdt, div = pr['ins'], pr['taps'], g, 1.0
                if sp is not None:
                    ins, taps = _split3_igemm(ins, taps, T, sp)
                    wdt, div = sp, 3.0
                p = _mk_igemm(ins, pr['out'], self.wpT, wdt, self.coutT_pad, self.cinT_pad, None, L.ACT_NONE, acc, 1, taps)
                self.dg.append((p, _backend('igemm', p), _igemm_flops(p) / div))

    def bwd(self, plan):
        inv = 1.0 / plan.prec.gscale
        gb = plan.grad_of(self.name + '.bias') if (self.has_bias and not self.bias_grad_zero) else None
        if self.act != L.ACT_NONE or gb is not None:
            L.call('dn_act_bwd', self.gout.ref(), self.out.ref(), self.act, L.ptr(gb), inv, L.ptr(plan.reduce_ws(self.Cout)),
                   plan.stream)
        if self.gsplit is not None:       # (hi, lo) planes of the output gradient: operands of both gradient GEMMs
            gh, gl = self.gsplit.planes(plan.prec.split)
            L.call('dn_split_bf16', self.gsplit.ref(), gh.ref(), gl.ref(), plan.stream)
        # the weight gradient only feeds the final unpack: it runs on the plan's side stream, concurrently with the data
        # gradient and the (HBM-bound) BatchNorm / activation backward passes of the layers below on the main stream
        side = plan.side_stream()
        if side is not None:
            ev = torch.cuda.Event()
            ev.record()
            side.wait_event(ev)
            with torch.cuda.stream(side):
                st = L.stream_ptr()
                if self.xq is not None:
                    L.call('dn_copy_view', self.x_cp.ref(), self.xq_cp.ref(), 0, st)
                for p, be, fl in self.wg:
                    L.call('dn_wgrad_run', C.byref(p), be, st, tag=('wgrad', be, fl, self.name))
                self._rowx_unpack(plan, st)
        else:
            if self.xq is not None:
                L.call('dn_copy_view', self.x_cp.ref(), self.xq_cp.ref(), 0, plan.stream)
            for p, be, fl in self.wg:
                L.call('dn_wgrad_run', C.byref(p), be, plan.stream, tag=('wgrad', be, fl, self.name))
            self._rowx_unpack(plan, plan.stream)
        if self.needs_dx:
            if self.dx_zero_first is not None:
                v = self.dx_zero_first
                if v.c0 == 0 and v.C == v.buf.C:
                    v.buf.t.zero_()
                else:       # a channel slice of a shared (concat) gradient buffer: clear the slice only
                    v.buf.t[..., v.c0:v.c0 + v.C].zero_()
            plan.run_group([(p, be, ('dgrad', be, fl, self.name)) for p, be, fl in self.dg])


    def _rowx_unpack(self, plan, st):
        if self.rowx:
            L.call('dn_rowx_unpack_wgrad', L.ptr(self.dwp), L.ptr(plan.grad_of(self.name + '.weight')), self.Cout, self.Cin, self.k,
                   self.cout_pad, self.cin_pad, 1.0 / plan.prec.gscale, st)


def conv_bn(plan, conv_name, bn_name, x, y_shape, out, k, act=L.ACT_RELU, pool=False, stride=1, pad=None, bias=True,
            needs_dx=True):
    """nn.Conv2d -> nn.BatchNorm2d -> activation (-> MaxPool2d(2, 2)).  Training: the convolution writes y, BNOp normalises with
    batch statistics (fused activation / pool).  Eval: BatchNorm is folded into the convolution (ConvOp fold_bn) and only
    the optional pool remains as a pass of its own.  y_shape = (N, H, W, C) of the convolution output."""
    N, H, W, Cc = y_shape
    fold = (not plan.training) and os.environ.get('DISPNET_B200_FOLD_BN', '1') != '0'
    if not fold:
        y = plan.new_buf(N, H, W, Cc).view()
        conv = plan.add(ConvOp(plan, conv_name, x, y, k, stride=stride, pad=pad, bias=bias, needs_dx=needs_dx, bn_follows=True))
        plan.add(BNOp(plan, bn_name, y, out, act, pool=pool, conv=conv))
        return
    if pool:
        y = plan.new_buf(N, H, W, Cc).view()
        plan.add(ConvOp(plan, conv_name, x, y, k, stride=stride, pad=pad, bias=bias, act=act, needs_dx=needs_dx, fold_bn=bn_name))
        plan.add(MaxPoolOp(plan, y, out, 2, 2, 0))
    else:
        plan.add(ConvOp(plan, conv_name, x, out, k, stride=stride, pad=pad, bias=bias, act=act, needs_dx=needs_dx, fold_bn=bn_name))


class HeadConvOp(Op):
    """predict_disp's nn.Conv2d(C, 1, 3, padding=1): a 9*C-term dot product per pixel on CUDA cores (HBM-bound), with a
    fused backward (data gradient + weight gradient + bias gradient in one pass)."""

    def __init__(self, plan, name, x, z):
        self.name, self.x, self.z = name, x, z
        plan.register_param(name + '.weight')
        plan.register_param(name + '.bias')

    def fwd(self, plan):
        L.call('dn_head_conv_fwd', self.x.ref(), L.ptr(plan.param(self.name + '.weight')), L.ptr(plan.param(self.name + '.bias')),
               self.z.ref(), plan.stream)

    def plan_bwd(self, plan):
        g = plan.prec.grad
        self.gz = self.z.grad_view(g)
        self.gx = self.x.grad_view(g)
        self.acc = int(self.gx.claim_grad_write())

    def bwd(self, plan):
        L.call('dn_head_conv_bwd', self.x.ref(), L.ptr(plan.param(self.name + '.weight')), self.gz.ref(), self.gx.ref(), self.acc,
               L.ptr(plan.grad_of(self.name + '.weight')), L.ptr(plan.grad_of(self.name + '.bias')), 1.0 / plan.prec.gscale,
               L.ptr(plan.reduce_ws(self.x.C * 5)), plan.stream)


class BNOp(Op):
    """nn.BatchNorm2d (+ residual add) + activation (+ MaxPool2d(2,2)).  `out=None`: statistics only (the dead bn1
    of Disp_res_50, models/Disp_res_50.py:143-145, whose running buffers still update)."""

    def __init__(self, plan, name, y, out, act=L.ACT_RELU, pool=False, residual=None, conv=None):
        self.name, self.y, self.out, self.act, self.pool, self.res = name, y, out, act, int(pool), residual
        self.conv = conv          # the ConvOp that produced y
        self.out2 = plan.shadow_of(out) if (out is not None and plan.training) else None
        Cc = y.C
        dev = plan.device
        self.sums = torch.zeros(2 * Cc, dtype=torch.float64, device=dev)
        self.red = torch.zeros(2 * Cc, dtype=torch.float64, device=dev)
        self.mean_invstd = torch.zeros(2 * Cc, dtype=torch.float32, device=dev)
        self.scale_shift = torch.zeros(2 * Cc, dtype=torch.float32, device=dev)
        self.count = float(y.N * y.H * y.W)
        if out is not None:
            plan.register_param(name + '.weight')
            plan.register_param(name + '.bias')

    def fwd(self, plan):
        st = plan.stream
        g, b = plan.param(self.name + '.weight'), plan.param(self.name + '.bias')
        rm, rv = plan.buffer(self.name + '.running_mean'), plan.buffer(self.name + '.running_var')
        if plan.training and not plan.bn_frozen:      # statistics, finalize and num_batches_tracked += 1 in one launch
            L.call('dn_bn_train_stats', self.y.ref(), L.ptr(g), L.ptr(b), L.ptr(rm), L.ptr(rv),
                   L.ptr(plan.buffer(self.name + '.num_batches_tracked')), 0.1, 1e-5, 1, None, L.ptr(self.mean_invstd),
                   L.ptr(self.scale_shift), L.ptr(plan.reduce_ws(self.y.C)), st)
        else:
            L.call('dn_bn_finalize', L.ptr(self.sums), self.count, L.ptr(g), L.ptr(b), L.ptr(rm), L.ptr(rv), 0.1, 1e-5,
                   0, 0, L.ptr(self.mean_invstd), L.ptr(self.scale_shift), self.y.C, st)
        if self.out is not None:
            L.call('dn_bn_apply', self.y.ref(), L.ptr(self.scale_shift), self.res.ref() if self.res else None, self.act,
                   self.pool, self.out.ref(), self.out2.ref() if self.out2 is not None else None, st)

    def plan_bwd(self, plan):
        if self.out is None:
            return
        g = plan.prec.grad
        self.gout = self.out.grad_view(g)
        self.gy = self.y.grad_view(g)
        assert not self.gy.claim_grad_write(), 'conv output feeding BN must have a single consumer'
        self.gres, self.gres_acc = None, 0
        if self.res is not None:
            self.gres = self.res.grad_view(g)
            self.gres_acc = int(self.gres.claim_grad_write())

    def bwd(self, plan):
        if self.out is None:
            return
        st = plan.stream
        # frozen statistics (plan.bn_frozen): mean / invstd are constants, dy = gamma * invstd * g -- the batch-mean terms of the
        # training formula are sum / count, so an infinite count removes them while dgamma / dbeta keep their sums
        gm, bt = plan.param(self.name + '.weight'), plan.param(self.name + '.bias')
        res = self.res.ref() if self.res else None
        L.call('dn_bn_bwd_reduce', self.gout.ref(), self.y.ref(), res, L.ptr(self.mean_invstd), L.ptr(gm), L.ptr(bt),
               self.act, self.pool, L.ptr(self.red), L.ptr(plan.reduce_ws(self.y.C)), st)
        L.call('dn_bn_bwd_apply', self.gout.ref(), self.y.ref(), res, L.ptr(self.mean_invstd), L.ptr(gm), L.ptr(bt),
               self.act, self.pool, L.ptr(self.red), (1e300 if plan.bn_frozen else self.count), 1.0 / plan.prec.gscale,
               L.ptr(plan.grad_of(self.name + '.weight')), L.ptr(plan.grad_of(self.name + '.bias')), self.gy.ref(),
               self.gres.ref() if self.gres else None, self.gres_acc, st)


class MaxPoolOp(Op):
    def __init__(self, plan, x, out, k, stride, pad):
        self.x, self.out, self.k, self.s, self.p = x, out, k, stride, pad

    def fwd(self, plan):
        L.call('dn_maxpool_fwd', self.x.ref(), self.out.ref(), self.k, self.s, self.p, plan.stream)

    def plan_bwd(self, plan):
        g = plan.prec.grad
        self.gout = self.out.grad_view(g)
        self.gx = self.x.grad_view(g)
        self.acc = int(self.gx.claim_grad_write())

    def bwd(self, plan):
        L.call('dn_maxpool_bwd', self.gout.ref(), self.x.ref(), self.gx.ref(), self.k, self.s, self.p, self.acc, plan.stream)


class ActOp(Op):
    """out = act(x) as its own pass (Disp_res_50 relu1 = relu(conv1), models/Disp_res_50.py:145)."""

    def __init__(self, plan, x, out, act, needs_dx=True):
        self.x, self.out, self.act, self.needs_dx = x, out, act, needs_dx

    def fwd(self, plan):
        L.call('dn_act_fwd', self.x.ref(), self.act, self.out.ref(), plan.stream)

    def plan_bwd(self, plan):
        g = plan.prec.grad
        self.gout = self.out.grad_view(g)
        self.gx = self.x.grad_view(g)
        self.acc = int(self.gx.claim_grad_write())

    def bwd(self, plan):
        L.call('dn_add_act_bwd', self.gout.ref(), self.out.ref(), self.act, self.gx.ref(), self.acc, None, 0, plan.stream)


class HeadOp(Op):
    """alpha * sigmoid(z) + beta -> fp32 NCHW output, plus the x2 up-sampled copy written into the next iconv's
    input slot (models/Disp_vgg_BN.py:168-171 nearest; models/DispNetS.py:119-120 bilinear + crop_like)."""

    def __init__(self, plan, z, alpha, beta, up=None, up_mode=0):
        self.z, self.alpha, self.beta, self.up, self.up_mode = z, float(alpha), float(beta), up, up_mode
        self.up_shadow = plan.shadow_of(up) if (up is not None and plan.training) else None
        self.idx = plan.add_output((z.N, 1, z.H, z.W))

    def fwd(self, plan):
        out = plan.outputs[self.idx]
        # (bit 4: the slot is the last slice of its concatenation buffer -- what follows it in the pixel record is zero padding,
        # so the kernel may write whole 32-byte sectors)
        tail_pad = 16 if (self.up is not None and self.up.c0 + self.up.C == self.up.buf.C and self.up.buf.readable_pad) else 0
        L.call('dn_head_fwd2', self.z.ref(), self.alpha, self.beta, L.ptr(out), self.up.ref() if self.up else None,
               self.up_shadow.ref() if self.up_shadow is not None else None, self.up_mode | tail_pad, plan.stream)

    def plan_bwd(self, plan):
        g = plan.prec.grad
        self.gz = self.z.grad_view(g)
        assert not self.gz.claim_grad_write()
        self.gup = self.up.grad_view(g) if self.up is not None else None

    def bwd(self, plan):
        go = plan.gouts[self.idx]
        L.call('dn_head_bwd', L.ptr(go), self.gup.ref() if self.gup else None, self.up_mode, self.z.ref(), self.alpha,
               plan.prec.gscale, self.gz.ref(), plan.stream)


class SigmoidOutOp(Op):
    """PoseExpNet explainability masks: sigmoid(conv) -> fp32 NCHW (models/PoseExpNet.py:82-85)."""

    def __init__(self, plan, z):
        self.z = z
        self.idx = plan.add_output((z.N, z.C, z.H, z.W))

    def fwd(self, plan):
        L.call('dn_sigmoid_nchw_fwd', self.z.ref(), L.ptr(plan.outputs[self.idx]), plan.stream)

    def plan_bwd(self, plan):
        self.gz = self.z.grad_view(plan.prec.grad)
        assert not self.gz.claim_grad_write()

    def bwd(self, plan):
        go = plan.gouts[self.idx]
        L.call('dn_sigmoid_nchw_bwd', L.ptr(go), L.ptr(plan.saved_outputs[self.idx]), plan.prec.gscale, self.gz.ref(),
               plan.stream)


class MeanOutOp(Op):
    """pose = scale * spatial mean of the 1x1 pose_pred conv (models/PoseExpNet.py:71-73) -> fp32 [N, C]."""

    def __init__(self, plan, z, scale):
        self.z, self.scale = z, float(scale)
        self.idx = plan.add_output((z.N, z.C))

    def fwd(self, plan):
        L.call('dn_spatial_mean_fwd', self.z.ref(), self.scale, L.ptr(plan.outputs[self.idx]), plan.stream)

    def plan_bwd(self, plan):
        self.gz = self.z.grad_view(plan.prec.grad)
        assert not self.gz.claim_grad_write()

    def bwd(self, plan):
        go = plan.gouts[self.idx]
        if go is None:
            self.gz.buf.t.zero_()
            return
        L.call('dn_spatial_mean_bwd', L.ptr(go), self.scale * plan.prec.gscale, self.gz.ref(), plan.stream)


class Plan:
    def __init__(self, module, device, precision, training, bn_frozen=False, only_train_dec=False):
        self.module = module
        self.device = device
        self.prec = Precision(precision)
        self.training = training
        # train.py:405-411 `--diff_lr`: BatchNorm modules put in eval() inside a training network normalise with their running
        # statistics (which are then constants of the graph) and leave the buffers alone; gamma / beta still train
        self.bn_frozen = bool(bn_frozen)
        # models/Disp_vgg_BN.py:148-153, models/Disp_res_50.py:154-160 `only_train_dec`: the encoder outputs are detached, so no
        # op before `encoder_end` (set by the model's _build_plan) runs a backward and its parameters receive no gradient
        self.only_train_dec = bool(only_train_dec)
        self.encoder_end = 0
        self.ops = []
        self.param_names = []
        self.out_shapes = []
        self.outputs = []
        self.saved_outputs = []
        self.gouts = []
        self.inputs = []
        self._params = {}
        self._grads = {}
        self.stream = None
        self._bwd_planned = False
        self._scratch = None
        self._ws = None
        self._gflat = None
        self._side = None
        self._dwp_arena, self._dwp_used = None, 0
        self._job_tables = {}
        self._fwd_graph, self._bwd_graphs, self._graph_key = None, {}, None
        self._fwd_warm = self._bwd_warm = 0
        self.generation = 0

    # ---- construction
    def add(self, op):
        self.ops.append(op)
        return op

    def register_param(self, name):
        if name not in self.param_names:
            self.param_names.append(name)

    def add_output(self, shape):
        self.out_shapes.append(tuple(shape))
        return len(self.out_shapes) - 1

    def new_buf(self, N, H, W, Cc, dtype=None):
        return Buf(N, H, W, Cc, dtype or self.prec.act, self.device)

    def shadow_of(self, v):
        """View (same geometry) into a gradient-dtype shadow of v's buffer, created on first use; None when activations
        and gradients share a dtype or the tensor cores are off.  Producers that can write two outputs in one pass
        (BatchNorm apply) fill it; ConvOp.bwd then skips the just-in-time conversion of its input."""
        if self.prec.act == self.prec.grad or self.prec.act == torch.float32 or not tc_enabled():
            return None
        b = v.buf
        if getattr(b, 'shadow', None) is None:
            b.shadow = Buf(b.N, b.H, b.W, b.C, self.prec.grad, self.device)
            b.shadow_valid = []
        b.shadow_valid.append((v.c0, v.c0 + v.C))
        return View(b.shadow, v.c0, v.C, v.H, v.W, v.off, v.sH, v.sW)

    def shadow_lookup(self, v):
        b = v.buf
        if getattr(b, 'shadow', None) is None:
            return None
        need, covered = v.c0, False
        for lo, hi in sorted(b.shadow_valid):       # union of the slices that producers fill
            if lo <= need < hi:
                need = hi
            if need >= v.c0 + v.C:
                covered = True
                break
        if covered:
            return View(b.shadow, v.c0, v.C, v.H, v.W, v.off, v.sH, v.sW)
        return None

    def scratch_buf(self, N, H, W, Cc, dtype):
        """Buf carved from the plan's shared scratch storage (valid only between consecutive launches)."""
        need = N * H * W * _pitch(Cc) * 4
        if self._scratch is None or self._scratch.numel() < need:
            raise RuntimeError('scratch storage too small')
        return Buf(N, H, W, Cc, dtype, self.device, storage=self._scratch)

    def reduce_ws(self, Cc):
        need = int(L.lib().dn_reduce_ws_floats(Cc))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.zeros(need, dtype=torch.float32, device=self.device)     # arrival counters start at zero
        return self._ws

    # ---- run-time lookups
    def param(self, name):
        return self._params[name]

    def buffer(self, name):
        return self._params[name]

    def grad_of(self, name):
        return self._grads[name]

    def bind(self, tensors):
        self._params = tensors

    # ---- eager execution -----------------------------------------------------------------------------------------
    def side_stream(self):
        if os.environ.get('DISPNET_B200_SIDE_STREAM', '1') == '0' or L.PROFILE is not None:
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def run_group(self, items):
        """Gather-convolutions that are independent of each other (the four output phases of a transposed convolution, the four
        input phases of a stride-2 data gradient).  Small ones (fewer tiles than SMs each) are launched on separate streams so
        that they share the GPU instead of running one under-filled persistent grid after the other; inside a CUDA-graph
        capture the fork / join becomes parallel branches of the graph."""
        small = len(items) > 1 and all(it[0].out.N * it[0].out.H * it[0].out.W <= 128 * 160 for it in items)
        if not small or L.PROFILE is not None or os.environ.get('DISPNET_B200_PHASE_STREAMS', '1') == '0':
            for p, be, tag in items:
                L.call('dn_igemm_run', C.byref(p), be, self.stream, tag=tag)
            return
        if getattr(self, '_phase_streams', None) is None:
            self._phase_streams = [torch.cuda.Stream(device=self.device) for _ in range(3)]
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        joins = []
        for i, (p, be, tag) in enumerate(items):
            if i == 0:
                L.call('dn_igemm_run', C.byref(p), be, self.stream, tag=tag)
                continue
            st = self._phase_streams[(i - 1) % 3]
            st.wait_event(fork)
            L.call('dn_igemm_run', C.byref(p), be, C.c_void_p(st.cuda_stream), tag=tag)
            ev = torch.cuda.Event()
            ev.record(st)
            joins.append(ev)
        for ev in joins:
            main.wait_event(ev)

    def dwp_alloc(self, numel):
        numel = _ru(numel, 64)
        out = self._dwp_arena[self._dwp_used:self._dwp_used + numel]
        self._dwp_used += numel
        assert out.numel() == numel, 'packed weight-gradient arena exhausted'
        return out

    def _run_jobs(self, which, rng=None, launch=True):
        """One batched dn_pack_jobs launch for the convolutions of ops[rng] (all ops of the direction when rng is None).
        launch=False only builds the device-side job table (a host-to-device copy, which a stream capture must not contain)."""
        tkey = (which, rng)
        key = (which, rng, self._ptr_key())
        ent = self._job_tables.get(tkey)
        if ent is None or ent[0] != key:
            lo, hi = rng if rng is not None else ((0 if which == 'fwd' else self._bwd_cut()), len(self.ops))
            jobs = [j for op in self.ops[lo:hi] if isinstance(op, ConvOp) for j in op.jobs(self, which)]
            dev = None
            if jobs:
                arr = (L.DnPackJob * len(jobs))(*jobs)
                dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
            ent = (key, dev, len(jobs))
            self._job_tables[tkey] = ent
        if ent[2] and launch:
            L.call('dn_pack_jobs', L.ptr(ent[1]), ent[2], self.stream)

    def _forward_impl(self, inputs):
        self.inputs = inputs
        self.stream = L.stream_ptr()
        self.outputs = [torch.empty(s, dtype=torch.float32, device=self.device) for s in self.out_shapes]
        for op in self.ops:
            if isinstance(op, ConvOp):
                op.pre_fwd(self)
        self._run_jobs('fwd')
        for op in self.ops:
            op.fwd(self)
        self.saved_outputs = self.outputs
        return tuple(self.outputs)

    def _ensure_gflat(self):
        if self._gflat is None:      # persistent fp32 gradient arena, one slice per parameter
            total = sum(self._params[n].numel() for n in self.param_names)
            self._gflat = torch.zeros(total, dtype=torch.float32, device=self.device)
            self._grads, self._goffs, o = {}, {}, 0
            for n in self.param_names:
                p = self._params[n]
                self._grads[n] = self._gflat[o:o + p.numel()].view(p.shape)
                self._goffs[n] = (o, o + p.numel())
                o += p.numel()

    def _backward_segment(self, gouts, lo, hi, first):
        """Backward of ops[lo:hi] (reverse order); `first`: this is the first segment of the step (clears the arenas and
        packs the data-gradient weights).  Leaves the gradients of the segment's parameters final in the flat arena."""
        self.stream = L.stream_ptr()
        self.gouts = gouts
        self._ensure_gflat()
        if first:
            self._gflat.zero_()
            self._dwp_arena.zero_()
            self._run_jobs('dgrad')
        if self.side_stream() is not None:          # the side stream must see the zeroed arenas / packed weights
            self._side.wait_stream(torch.cuda.current_stream())
        for op in reversed(self.ops[lo:hi]):
            op.bwd(self)
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        self._run_jobs('unpack', (lo, hi))
        return self._gflat

    def _backward_impl(self, gouts):
        return self._backward_segment(gouts, self._bwd_cut(), len(self.ops), True)

    def _segments(self, distributed):
        """Op ranges (forward order) whose backward runs as one unit.  Data parallel: three ranges, cut where the cumulative
        parameter count (forward order) passes 10 % and 60 % -- the all-reduce of a range is issued as soon as its backward is
        enqueued and overlaps the backward of the ranges below it, so only the smallest one (the first layers, ~10 % of the
        bytes) is exposed at the end of the step."""
        cut, n = self._bwd_cut(), len(self.ops)
        if not distributed or os.environ.get('DISPNET_B200_DP_SEGMENTS', '3') == '1':
            return [(cut, n)]
        sizes = []
        for op in self.ops[cut:]:
            nm = getattr(op, 'name', None)
            sizes.append(sum(self._params[k].numel() for k in ((nm + '.weight', nm + '.bias') if nm else ()) if k in self.param_names))
        total = float(sum(sizes)) or 1.0
        marks, acc, want = [], 0.0, [0.10, 0.60]
        for i, sz in enumerate(sizes):
            acc += sz
            while want and acc / total >= want[0]:
                if 0 < i + 1 < len(sizes) and (not marks or marks[-1] != cut + i + 1):
                    marks.append(cut + i + 1)
                want.pop(0)
        edges = [cut] + marks + [n]
        return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]

    def _segment_span(self, lo, hi):
        """[begin, end) of the flat gradient arena that belongs to the parameters of ops[lo:hi] (contiguous: parameters are
        registered in op order)."""
        names = []
        for op in self.ops[lo:hi]:
            nm = getattr(op, 'name', None)
            if nm:
                names += [k for k in (nm + '.weight', nm + '.bias') if k in self._goffs]
        if not names:
            return None
        b, e = min(self._goffs[k][0] for k in names), max(self._goffs[k][1] for k in names)
        assert sum(self._goffs[k][1] - self._goffs[k][0] for k in set(names)) == e - b, 'segment parameters are not contiguous'
        return b, e

    # ---- CUDA-graph execution: the whole forward (and backward) op list is captured once per plan and replayed, so a
    # step costs two graph launches instead of ~300 host-side launches (the reference's loop is launch-bound)
    def _ptr_key(self):
        return tuple(self._params[n].data_ptr() for n in self.param_names) + tuple(
            v.data_ptr() for k, v in self._params.items() if k.endswith(('running_mean', 'running_var', 'num_batches_tracked')))

    def run_forward(self, inputs):
        for x in inputs:
            L.require_cuda(x)
        inputs = [x.contiguous().float() for x in inputs]
        self.generation += 1
        if not graphs_enabled():
            return self._forward_impl(inputs)
        key = self._ptr_key()
        if self._fwd_graph is not None and key != self._graph_key:
            self._fwd_graph, self._bwd_graphs, self._fwd_warm, self._bwd_warm = None, {}, 0, 0     # storage moved: re-capture
        self._graph_key = key
        if self._fwd_graph is None:
            if self._fwd_warm < 2:          # lazy one-time initialisation must happen outside a capture
                self._fwd_warm += 1
                return self._forward_impl(inputs)
            self._static_in = [x.clone() for x in inputs]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            calls0 = L.CALLS
            with torch.cuda.graph(g, capture_error_mode='thread_local'):
                self._static_out = self._forward_impl(self._static_in)
            self._fwd_graph, self._fwd_calls = g, L.CALLS - calls0
        for s_, x in zip(self._static_in, inputs):
            s_.copy_(x)
        self.inputs = self._static_in
        self.outputs = self.saved_outputs = list(self._static_out)
        self._fwd_graph.replay()
        L.CALLS += self._fwd_calls
        return tuple(o.clone() for o in self._static_out)

    def plan_backward(self):
        if self._bwd_planned:
            return
        convs = [op for op in self.ops if isinstance(op, ConvOp)]
        big = max([op.x.buf.t.numel() for op in convs] + [1])
        self._scratch = torch.empty(big * 4, dtype=torch.uint8, device=self.device)
        self._dwp_arena = torch.zeros(sum(_ru(op.k * op.k * op.cout_pad * op.cin_pad, 64) for op in convs) + 64,
                                      dtype=torch.float32, device=self.device)
        self._dwp_used = 0
        for op in reversed(self.ops[self._bwd_cut():]):
            op.plan_bwd(self)
        self._bwd_planned = True

    def _bwd_cut(self):
        return self.encoder_end if self.only_train_dec else 0

    def frozen_param_names(self):
        """Parameters that receive no gradient in this plan (the encoder under `only_train_dec`)."""
        if not self.only_train_dec:
            return set()
        live = set()
        for op in self.ops[self._bwd_cut():]:
            n = getattr(op, 'name', None)
            if n is not None:
                live.update((n + '.weight', n + '.bias'))
        return set(self.param_names) - live

    def run_backward(self, gouts):
        self.plan_backward()
        gouts = [None if g is None else g.contiguous().float() for g in gouts]
        group = getattr(self.module, '_dp_group', None)
        segs = list(reversed(self._segments(group is not None)))          # execution order: last layers first
        use_graph = graphs_enabled() and self._fwd_graph is not None
        pattern = (tuple(g is not None for g in gouts), len(segs))
        ent = self._bwd_graphs.get(pattern) if use_graph else None
        if use_graph and ent is None:
            if self._bwd_warm < 1:
                self._bwd_warm += 1
                use_graph = False
            else:
                static_g = [None if g is None else g.clone() for g in gouts]
                self._ensure_gflat()
                self._run_jobs('dgrad', launch=False)
                for rng in segs:
                    self._run_jobs('unpack', rng, launch=False)
                torch.cuda.synchronize()
                graphs, calls0 = [], L.CALLS
                for si, (lo, hi) in enumerate(segs):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode='thread_local'):
                        self._backward_segment(static_g, lo, hi, si == 0)
                    graphs.append(g)
                ent = (graphs, static_g, L.CALLS - calls0)
                self._bwd_graphs[pattern] = ent
                # the captures only recorded the work: run it now through the common replay path below
        if use_graph:
            graphs, static_g, ncalls = ent
            for s_, x in zip(static_g, gouts):
                if s_ is not None:
                    s_.copy_(x)
            L.CALLS += ncalls
        self._ensure_gflat()
        out = torch.empty_like(self._gflat)     # autograd may keep / accumulate into what we return: hand out a private copy
        works = []
        for si, (lo, hi) in enumerate(segs):
            if use_graph:
                graphs[si].replay()
            else:
                self._backward_segment(gouts, lo, hi, si == 0)
            span = self._segment_span(lo, hi)
            if span is None:
                continue
            seg = out[span[0]:span[1]]
            seg.copy_(self._gflat[span[0]:span[1]])
            if group is not None:
                # data parallel: the NCCL all-reduce (mean) of this range's gradients starts as soon as the range's backward
                # is enqueued and runs on NCCL's stream, concurrently with the backward of the remaining ranges
                import torch.distributed as dist
                works.append(dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=group, async_op=True))
        for w in works:
            w.wait()                            # stream-level wait: the compute stream resumes after the last all-reduce
        covered = self._segment_span(self._bwd_cut(), len(self.ops))
        res = {}
        for n in self.param_names:
            b, e = self._goffs[n]
            if covered is not None and covered[0] <= b and e <= covered[1]:
                res[n] = out[b:e].view(self._params[n].shape)
        return res


class _NetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, n_inputs, *tensors):
        inputs = tensors[:n_inputs]
        outs = plan.run_forward(inputs)
        ctx.plan = plan
        ctx.generation = plan.generation
        ctx.n_inputs = n_inputs
        ctx.param_order = plan._fn_param_order
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        plan = ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError('dispnet_b200: backward() after a newer forward() of the same module/shape: the activation '
                               'arena has been overwritten (run forward -> backward in lockstep, as train.py does)')
        grads = plan.run_backward(gouts)
        res = [None, None] + [None] * ctx.n_inputs
        dead = plan.frozen_param_names()
        for name in ctx.param_order:
            res.append(None if name in dead else grads.get(name))
        return tuple(res)


class PlannedModule(torch.nn.Module):
    """Base of the drop-in model classes: owns parameters/buffers with the reference's state_dict keys and runs
    the forward through a cached Plan inside one autograd node."""

    precision = None

    def _build_plan(self, plan, shapes):
        raise NotImplementedError

    def _plan_for(self, inputs):
        dev = inputs[0].device
        prec = self.precision or default_precision()
        bn_frozen = self.training and any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and not m.training
                                          for m in self.modules())
        only_dec = self.training and bool(getattr(self, 'only_train_dec', False))
        key = (tuple(tuple(x.shape) for x in inputs), self.training, prec, str(dev), bn_frozen, only_dec)
        plans = self.__dict__.setdefault('_plans', {})
        if key not in plans:
            plan = Plan(self, dev, prec, self.training, bn_frozen, only_dec)
            with torch.no_grad():
                self._build_plan(plan, [tuple(x.shape) for x in inputs])
            plans[key] = plan
        return plans[key]

    def _run(self, inputs):
        if not inputs[0].is_cuda:
            raise RuntimeError('dispnet_b200 models run on CUDA only (no CPU fallback): move the module and inputs to a '
                               'B200 device')
        plan = self._plan_for(inputs)
        named = dict(self.named_parameters())
        tensors = dict(named)
        tensors.update(dict(self.named_buffers()))
        plan.bind(tensors)
        order = [n for n in plan.param_names]
        plan._fn_param_order = order
        with torch.cuda.device(inputs[0].device):       # kernels go to the current stream of the tensors' device
            return _NetFn.apply(plan, len(inputs), *inputs, *[named[n] for n in order])
