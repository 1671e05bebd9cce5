"""CPU: the host-side plan logic (tap tables of forward / data-gradient / weight-gradient problems for every convolution
flavour, phase views of transposed convolutions, crops, concat slices, gradient-accumulation bookkeeping) checked by
*interpreting* the gather-convolution problems the engine emits with plain torch indexing and comparing with
F.conv2d / F.conv_transpose2d and their autograd gradients.  No kernel is launched; the C ABI is not called."""
import pytest
import torch
import torch.nn.functional as F

from supervised_dispnet_b200 import engine as E
from supervised_dispnet_b200 import _lib as L


class FakePlan(E.Plan):
    def __init__(self):
        super().__init__(None, torch.device('cpu'), 'fp32', True)


def fill_view(v, nchw):
    """write an NCHW tensor into a view (test helper, CPU)."""
    t = v.buf.t.view(-1)
    N, Cc, H, W = nchw.shape
    idx = (torch.arange(N).view(-1, 1, 1, 1) * v.sN + torch.arange(H).view(1, -1, 1, 1) * v.sH
           + torch.arange(W).view(1, 1, -1, 1) * v.sW + torch.arange(Cc).view(1, 1, 1, -1) + v.off + v.c0)
    t[idx] = nchw.permute(0, 2, 3, 1)


def interp_igemm(ins, out, taps, stride, weight_tco_ci, accumulate=False):
    """reference interpreter of the dn_igemm definition (include/dispnet_b200.h) on views."""
    src = [v.to_nchw().double() for v in ins]
    N, Co, Ho, Wo = out.N, out.C, out.H, out.W
    res = torch.zeros(N, Co, Ho, Wo, dtype=torch.double)
    for (s, dh, dw, wt) in taps:
        x = src[s]
        Wm = weight_tco_ci[wt][:Co, :x.shape[1]].double()          # [co][ci]
        for ho in range(Ho):
            hi = ho * stride + dh
            if hi < 0 or hi >= x.shape[2]:
                continue
            for wo in range(Wo):
                wi = wo * stride + dw
                if wi < 0 or wi >= x.shape[3]:
                    continue
                res[:, :, ho, wo] += x[:, :, hi, wi] @ Wm.t()
    if accumulate:
        res += out.to_nchw().double()
    fill_view(out, res.float())


def interp_wgrad(ps, q, taps, stride, T, Co, Ci):
    P = [v.to_nchw().double() for v in ps]
    Q = q.to_nchw().double()
    R = torch.zeros(T, Co, Ci, dtype=torch.double)
    for (s, dh, dw, wt) in taps:
        p = P[s]
        for h in range(p.shape[2]):
            hi = h * stride + dh
            if hi < 0 or hi >= Q.shape[2]:
                continue
            for w in range(p.shape[3]):
                wi = w * stride + dw
                if wi < 0 or wi >= Q.shape[3]:
                    continue
                R[wt] += p[:, :, h, w].t() @ Q[:, :, hi, wi]
    return R


def taps_of(p):
    return [(p.taps[i].src, p.taps[i].dh, p.taps[i].dw, p.taps[i].wt) for i in range(p.ntaps)]


CASES = [
    dict(cin=5, cout=4, k=3, stride=1, pad=1, transposed=False, hw=(6, 7)),
    dict(cin=3, cout=4, k=7, stride=2, pad=3, transposed=False, hw=(9, 10)),
    dict(cin=3, cout=4, k=7, stride=2, pad=3, transposed=False, hw=(10, 12)),
    dict(cin=4, cout=3, k=5, stride=2, pad=2, transposed=False, hw=(8, 8)),
    dict(cin=4, cout=3, k=3, stride=2, pad=1, transposed=False, hw=(4, 13)),      # DispNetS conv6 / conv7 shapes: odd widths
    dict(cin=4, cout=3, k=3, stride=2, pad=1, transposed=False, hw=(2, 7)),
    dict(cin=3, cout=4, k=3, stride=2, pad=1, transposed=False, hw=(5, 3)),
    dict(cin=4, cout=6, k=1, stride=2, pad=0, transposed=False, hw=(6, 6)),
    dict(cin=4, cout=3, k=4, stride=2, pad=1, transposed=True, out_pad=0, hw=(3, 5)),
    dict(cin=3, cout=5, k=3, stride=2, pad=1, transposed=True, out_pad=1, hw=(2, 4)),
    dict(cin=3, cout=5, k=3, stride=2, pad=1, transposed=True, out_pad=1, hw=(1, 4), crop=(2, 7)),
]


@pytest.mark.parametrize('c', CASES)
def test_conv_problem_tables_match_torch(c):
    torch.manual_seed(0)
    N, (H, W) = 2, c['hw']
    plan = FakePlan()
    xin = plan.new_buf(N, H, W, c['cin']).view()
    if c['transposed']:
        Ho = (H - 1) * 2 - 2 * c['pad'] + c['k'] + c.get('out_pad', 0)
        Wo = (W - 1) * 2 - 2 * c['pad'] + c['k'] + c.get('out_pad', 0)
    else:
        Ho = (H + 2 * c['pad'] - c['k']) // c['stride'] + 1
        Wo = (W + 2 * c['pad'] - c['k']) // c['stride'] + 1
    if 'crop' in c:
        Ho, Wo = min(Ho, c['crop'][0]), min(Wo, c['crop'][1])
    out = plan.new_buf(N, Ho, Wo, c['cout']).view()
    op = E.ConvOp(plan, 'conv', xin, out, c['k'], stride=c['stride'], pad=c['pad'], transposed=c['transposed'], bias=False)
    x = torch.randn(N, c['cin'], H, W)
    k = c['k']
    if c['transposed']:
        w = torch.randn(c['cin'], c['cout'], k, k)
        w_t = w.permute(2, 3, 1, 0).reshape(k * k, c['cout'], c['cin'])       # [t][co][ci]
    else:
        w = torch.randn(c['cout'], c['cin'], k, k)
        w_t = w.permute(2, 3, 0, 1).reshape(k * k, c['cout'], c['cin'])
    fill_view(xin, x)
    # ---- forward
    for pr in op.fwd_probs:
        interp_igemm(pr['ins'], pr['out'], pr['taps'], pr['stride'], w_t)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    if c['transposed']:
        y = F.conv_transpose2d(xr, wr, None, 2, c['pad'], c.get('out_pad', 0))[:, :, :Ho, :Wo]
    else:
        y = F.conv2d(xr, wr, None, c['stride'], c['pad'])
    assert torch.allclose(out.to_nchw(), y.detach(), atol=1e-4)
    # ---- backward problems
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    plan._dwp_arena = torch.zeros(1 << 16)      # normally sized by Plan.plan_backward()
    op.plan_bwd(plan)
    fill_view(op.gout, gy)
    w_tT = w_t.transpose(1, 2).contiguous()                                  # [t][ci][co]
    if op.dx_zero_first is not None:
        op.dx_zero_first.buf.t.zero_()
    for p, be, fl in op.dg:
        ins = [E.View(op.gout.buf, op.gout.c0, op.gout.C, p.inp[i].H, p.inp[i].W,
                      (p.inp[i].ptr - op.gout.buf.t.data_ptr()) // 4 - op.gout.c0, p.inp[i].sH, p.inp[i].sW) for i in range(p.nsrc)]
        gxbuf = xin.grad_view(torch.float32)
        outv = E.View(gxbuf.buf, gxbuf.c0, gxbuf.C, p.out.H, p.out.W, (p.out.ptr - gxbuf.buf.t.data_ptr()) // 4 - gxbuf.c0,
                      p.out.sH, p.out.sW)
        interp_igemm(ins, outv, taps_of(p), p.stride, w_tT, accumulate=bool(p.accumulate))
    assert torch.allclose(xin.grad_view(torch.float32).to_nchw(), xr.grad, atol=1e-4)
    R = torch.zeros(k * k, c['cout'], c['cin'], dtype=torch.double)
    for p, be, fl in op.wg:
        ps = [E.View(op.gout.buf, op.gout.c0, op.gout.C, p.p[i].H, p.p[i].W, (p.p[i].ptr - op.gout.buf.t.data_ptr()) // 4 - op.gout.c0,
                     p.p[i].sH, p.p[i].sW) for i in range(p.nsrc)]
        qv = E.View(xin.buf, xin.c0, xin.C, p.q.H, p.q.W, (p.q.ptr - xin.buf.t.data_ptr()) // 4 - xin.c0, p.q.sH, p.q.sW)
        R += interp_wgrad(ps, qv, taps_of(p), p.stride, k * k, c['cout'], c['cin'])
    if c['transposed']:
        gw = R.view(k, k, c['cout'], c['cin']).permute(3, 2, 0, 1)
    else:
        gw = R.view(k, k, c['cout'], c['cin']).permute(2, 3, 0, 1)
    assert torch.allclose(gw.float(), wr.grad, atol=1e-3)


def test_gradient_accumulation_bookkeeping():
    """first writer of a gradient region overwrites, later writers accumulate; partial overlaps are rejected."""
    plan = FakePlan()
    cat = plan.new_buf(1, 4, 4, 24)
    whole, a, b = cat.view(), cat.view().channels(0, 8), cat.view().channels(8, 16)
    assert whole.grad_view(torch.float32).claim_grad_write() is False      # iconv dgrad writes the whole concat buffer
    assert b.grad_view(torch.float32).claim_grad_write() is True           # the next encoder conv then adds into its slice
    assert a.grad_view(torch.float32).claim_grad_write() is True
    other = plan.new_buf(1, 4, 4, 24)
    assert other.view().channels(0, 8).grad_view(torch.float32).claim_grad_write() is False
    with pytest.raises(RuntimeError):
        other.view().channels(4, 8).grad_view(torch.float32).claim_grad_write()


def test_views_phase_crop_slice_addressing():
    plan = FakePlan()
    b = plan.new_buf(2, 5, 7, 10)
    x = torch.randn(2, 10, 5, 7)
    fill_view(b.view(), x)
    assert torch.equal(b.view().channels(3, 4).to_nchw(), x[:, 3:7])
    assert torch.equal(b.view().crop(2, 3).to_nchw(), x[:, :, :2, :3])
    for a in range(2):
        for c in range(2):
            assert torch.equal(b.view().phase(a, c).to_nchw(), x[:, :, a::2, c::2])
    assert E._pad_ok(b.view()) == 1 and E._pad_ok(b.view().channels(0, 8)) == 0 and E._pad_ok(b.view().channels(8, 2)) == 1


@pytest.mark.parametrize('cls,shape', [('Disp_vgg_BN', (1, 3, 64, 96)), ('DispNetS', (1, 3, 128, 160)), ('Disp_res_50', (1, 3, 64, 96))])
def test_plans_build_on_cpu_and_backward_plan_is_consistent(cls, shape):
    """Plan construction (buffers, views, op list, gradient bookkeeping) runs without a GPU; every parameter that the
    reference trains is registered exactly once and nothing dead is."""
    import supervised_dispnet_b200 as S
    m = getattr(S.models, cls)()
    plan = E.Plan(m, torch.device('cpu'), 'fp32', True)
    m._build_plan(plan, [shape])
    plan.plan_backward()
    named = dict(m.named_parameters())
    assert len(set(plan.param_names)) == len(plan.param_names)
    trained = {n for n, p in named.items() if p.requires_grad and not n.startswith('bn1.')}
    assert set(plan.param_names) == trained
    assert len(plan.out_shapes) == 4


# ---- train-loop helpers (host side of supervised_dispnet_b200.train) ---------------------------------------------------------
def test_device_prefetcher_order_and_limit_cpu():
    """_DevicePrefetcher hands out the loader's batches in order, keeps nested list / tuple structure, and never pulls more
    than `limit` batches (the reference's loop breaks after epoch_size batches, train.py:536)."""
    import torch
    from supervised_dispnet_b200.train import _DevicePrefetcher
    pulled = []

    def gen():
        for i in range(10):
            pulled.append(i)
            yield (torch.full((2,), float(i)), [torch.full((1,), i + 0.5), torch.full((1,), i + 0.25)], None)
    got = list(_DevicePrefetcher(gen(), 'cpu', 4))
    assert len(got) == 4 and pulled == [0, 1, 2, 3]
    for i, (a, refs, none) in enumerate(got):
        assert none is None and isinstance(refs, list)
        assert a.tolist() == [float(i)] * 2 and refs[0].item() == i + 0.5 and refs[1].item() == i + 0.25
    assert list(_DevicePrefetcher(iter([]), 'cpu', 3)) == []


def test_loss_reader_average_cpu():
    """_LossReader feeds every pushed loss into the meter exactly once (CPU path reads immediately)."""
    import torch
    from supervised_dispnet_b200.train import AverageMeter, _LossReader
    m = AverageMeter(precision=4)
    r = _LossReader(m, 'cpu')
    for v in (1.0, 2.0, 6.0):
        r.push(torch.tensor(v), 4)
    r.flush()
    assert m.count == 12 and abs(m.avg[0] - 3.0) < 1e-6


def test_dp_segments_partition_the_gradient_arena():
    """Host logic of the overlapped data-parallel all-reduce (engine.Plan._segments / _segment_span): three op ranges that
    cover the op list, whose parameter spans tile the flat gradient arena, the range that finishes last (the first layers)
    holding about a tenth of the bytes."""
    import torch
    import supervised_dispnet_b200 as S
    m = S.models.Disp_vgg_BN()
    x = torch.zeros(1, 3, 64, 96)
    plan = m._plan_for([x])
    tensors = dict(m.named_parameters())
    tensors.update(dict(m.named_buffers()))
    plan.bind(tensors)
    plan.plan_backward()
    plan._ensure_gflat()
    segs = plan._segments(True)
    assert len(segs) == 3 and segs[0][0] == 0 and segs[-1][1] == len(plan.ops)
    assert all(a[1] == b[0] for a, b in zip(segs[:-1], segs[1:]))
    spans = [plan._segment_span(lo, hi) for lo, hi in segs]
    assert spans[0][0] == 0 and spans[-1][1] == plan._gflat.numel()
    assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
    frac = (spans[0][1] - spans[0][0]) / plan._gflat.numel()
    assert 0.05 < frac < 0.2, frac
    assert plan._segments(False) == [(0, len(plan.ops))]


@pytest.mark.parametrize('name', ['Disp_vgg_BN', 'DispNetS', 'Disp_res_50'])
def test_checkpoint_roundtrip_with_reference_format(tmp_path, name):
    """SURVEY 8(f4) / utils.py:79-93: a checkpoint written by this package's save_checkpoint loads into the REFERENCE model
    class the way train.py:280-281 does (strict), and a checkpoint written by the reference's own save_checkpoint loads into
    this package's model; the two files, their keys and the model_best copy follow the reference layout."""
    import pathlib
    import torch
    import supervised_dispnet_b200 as S
    from oracle import refshim as R
    root = R.find_root()
    if root is None:
        pytest.skip('reference checkout not available')
    ref = R.import_reference(root, with_train=False)
    ours = getattr(S.models, name)()
    torch.manual_seed(3)
    ours.init_weights()
    pose = S.models.PoseExpNet(2, False)
    pose.init_weights()
    opt = torch.optim.Adam([p for p in ours.parameters() if p.requires_grad], lr=1e-4)
    d1 = tmp_path / 'ours'
    d1.mkdir()
    S.utils.save_checkpoint(d1, {'epoch': 4, 'state_dict': ours.state_dict(), 'optimizer': opt.state_dict()},
                            {'epoch': 4, 'state_dict': pose.state_dict()}, is_best=True, epoch=3, record=True)
    assert sorted(p.name for p in d1.iterdir()) == ['dispnet_checkpoint.pth.tar', 'dispnet_model_best.pth.tar',
                                                    'exp_pose_checkpoint.pth.tar', 'exp_pose_model_best.pth.tar', 'weights_3']
    rnet = getattr(ref.models, name)()
    weights = torch.load(d1 / 'dispnet_checkpoint.pth.tar', weights_only=False)
    rnet.load_state_dict(weights['state_dict'])                              # strict, as train.py:281
    assert weights['epoch'] == 4 and 'optimizer' in weights
    for k, v in rnet.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k
    rpose = ref.models.PoseExpNet(2, False)
    rpose.load_state_dict(torch.load(d1 / 'exp_pose_model_best.pth.tar', weights_only=False)['state_dict'], strict=False)
    # the other direction, written by the reference's own utils.save_checkpoint (it needs path.py's Path: pathlib + makedirs_p)
    d2 = tmp_path / 'ref'
    d2.mkdir()
    torch.manual_seed(4)
    rnet.init_weights()
    ref.utils.save_checkpoint(pathlib.Path(d2), {'epoch': 7, 'state_dict': rnet.state_dict()}, {'epoch': 7, 'state_dict': rpose.state_dict()},
                              is_best=False, epoch=6)
    assert S.utils.load_checkpoint(d2 / 'dispnet_checkpoint.pth.tar', ours) == 7
    for k, v in ours.state_dict().items():
        assert torch.equal(v, rnet.state_dict()[k]), k


def test_channel_pitch_and_readable_extent():
    """Odd channel counts get whole-TMA-row pitches (16 / 32 / multiples of 64) whose zero padding a gather-convolution may read:
    dn_view.c_ext = channels readable behind the view inside a pixel record.  Buffers carved from scratch storage are not
    zero-initialised and declare nothing beyond their channels."""
    assert [E._pitch(c) for c in (3, 9, 16, 17, 24, 32, 33, 64, 97, 193, 385, 512)] == [16, 16, 16, 32, 24, 32, 64, 64, 128, 256, 448, 512]
    dev = torch.device('cpu')
    b = E.Buf(2, 4, 6, 97, torch.float16, dev)
    assert b.Cp == 128 and b.t.shape == (2, 4, 6, 128) and float(b.t.abs().sum()) == 0.0
    v = b.view()
    assert (v.dn().C, v.dn().c_ext, v.dn().sW) == (97, 128, 128)
    s = v.channels(32, 64)                  # a slice of the concatenation buffer: the neighbouring slice + padding are readable
    assert (s.dn().C, s.dn().c_ext) == (64, 96)
    ph = v.phase(1, 0)                      # stride-2 sub-lattice: same pixel records, doubled strides
    assert (ph.dn().c_ext, ph.dn().sW, ph.dn().H) == (128, 256, 2)
    scratch = torch.empty(2 * 4 * 6 * 128 * 2, dtype=torch.uint8)
    sb = E.Buf(2, 4, 6, 97, torch.float16, dev, storage=scratch)
    assert sb.view().dn().c_ext == 97


def test_igemm_struct_has_the_phase_fields():
    """dn_igemm carries the merged / channel-stacked phase description of a transposed convolution (include/dispnet_b200.h)."""
    p = L.DnIgemm()
    p.nphase, p.phase_cout = 4, 16
    for i in range(4):
        p.phase_off[i] = 1000 * i
    assert [int(p.phase_off[i]) for i in range(4)] == [0, 1000, 2000, 3000] and (p.nphase, p.phase_cout) == (4, 16)
