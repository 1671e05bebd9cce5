"""DispNetS (reference: models/DispNetS.py) -- same constructor, `init_weights`, 64 state_dict keys
(`convK.{0,2}.*`, `upconvK.0.*`, `iconvK.0.*`, `predict_dispK.0.*`); forward (:93-140) runs on
libdispnet_b200.so: 14 strided/unit convs + ReLU, 7 ConvTranspose2d(k3,s2,p1,op1) + ReLU as phase
convolutions cropped (`crop_like`) by writing only the kept region, concat-free iconvs, bilinear x2
(align_corners=False) disparity up-sampling fused into the heads."""
import torch

from .. import engine as E
from ._common import ACT_NONE, ACT_RELU, alpha_beta, conv_block, predict_disp, upconv_block, xavier_init
import torch.nn as nn


def _downsample_conv(c_in, c_out, k):
    return nn.Sequential(nn.Conv2d(c_in, c_out, k, 2, (k - 1) // 2), nn.ReLU(inplace=True),
                         nn.Conv2d(c_out, c_out, k, 1, (k - 1) // 2), nn.ReLU(inplace=True))


class DispNetS(E.PlannedModule):

    def __init__(self, datasets='kitti'):
        super().__init__()
        self.alpha, self.beta = alpha_beta(datasets)
        cp = [32, 64, 128, 256, 512, 512, 512]
        ks = [7, 5, 3, 3, 3, 3, 3]
        c_in = 3
        for i, (c, k) in enumerate(zip(cp, ks)):
            setattr(self, 'conv%d' % (i + 1), _downsample_conv(c_in, c, k))
            c_in = c
        up = [512, 512, 256, 128, 64, 32, 16]
        c_in = cp[6]
        for i, c in zip(range(7, 0, -1), up):
            setattr(self, 'upconv%d' % i, upconv_block(c_in, c, 3, 1, 1, nn.ReLU))
            c_in = c
        iin = [up[0] + cp[5], up[1] + cp[4], up[2] + cp[3], up[3] + cp[2], 1 + up[4] + cp[1], 1 + up[5] + cp[0], 1 + up[6]]
        for i, ci, c in zip(range(7, 0, -1), iin, up):
            setattr(self, 'iconv%d' % i, conv_block(ci, c, 3, 1, nn.ReLU))
        self.predict_disp4 = predict_disp(up[3])
        self.predict_disp3 = predict_disp(up[4])
        self.predict_disp2 = predict_disp(up[5])
        self.predict_disp1 = predict_disp(up[6])

    def init_weights(self, use_pretrained_weights=False):
        xavier_init(self, with_linear=False)

    def _build_plan(self, plan, shapes):
        N, Cin, H, W = shapes[0]
        assert Cin == 3
        nb = plan.new_buf
        inp = plan.add(E.InputOp(plan, shapes))
        cp = [32, 64, 128, 256, 512, 512, 512]
        ks = [7, 5, 3, 3, 3, 3, 3]
        up = [512, 512, 256, 128, 64, 32, 16]
        # spatial sizes of conv1..conv7 outputs
        hs, ws = [], []
        h, w = H, W
        for k in ks:
            p = (k - 1) // 2
            h, w = (h + 2 * p - k) // 2 + 1, (w + 2 * p - k) // 2 + 1
            hs.append(h); ws.append(w)
        # iconv input buffers at the resolution of conv6..conv1 and the input: [upconv | skip | disp_up]
        cat = {7: nb(N, hs[5], ws[5], up[0] + cp[5]), 6: nb(N, hs[4], ws[4], up[1] + cp[4]),
               5: nb(N, hs[3], ws[3], up[2] + cp[3]), 4: nb(N, hs[2], ws[2], up[3] + cp[2]),
               3: nb(N, hs[1], ws[1], up[4] + cp[1] + 1), 2: nb(N, hs[0], ws[0], up[5] + cp[0] + 1),
               1: nb(N, H, W, up[6] + 1)}
        skip_dst = {6: cat[7].view().channels(up[0], cp[5]), 5: cat[6].view().channels(up[1], cp[4]),
                    4: cat[5].view().channels(up[2], cp[3]), 3: cat[4].view().channels(up[3], cp[2]),
                    2: cat[3].view().channels(up[4], cp[1]), 1: cat[2].view().channels(up[5], cp[0])}
        x = inp.out
        for i in range(7):
            mid = nb(N, hs[i], ws[i], cp[i]).view()
            plan.add(E.ConvOp(plan, 'conv%d.0' % (i + 1), x, mid, ks[i], stride=2, act=ACT_RELU, needs_dx=(i > 0)))
            out = skip_dst.get(i + 1) or nb(N, hs[i], ws[i], cp[i]).view()
            plan.add(E.ConvOp(plan, 'conv%d.2' % (i + 1), mid, out, ks[i], act=ACT_RELU))
            x = out
        heads = {}
        slot = {4: cat[3].view().channels(up[4] + cp[1], 1), 3: cat[2].view().channels(up[5] + cp[0], 1),
                2: cat[1].view().channels(up[6], 1)}
        for lvl in range(7, 0, -1):
            c = up[7 - lvl]
            plan.add(E.ConvOp(plan, 'upconv%d.0' % lvl, x, cat[lvl].view().channels(0, c), 3, stride=2, pad=1,
                              transposed=True, act=ACT_RELU))
            o = nb(N, cat[lvl].H, cat[lvl].W, c).view()
            plan.add(E.ConvOp(plan, 'iconv%d.0' % lvl, cat[lvl].view(), o, 3, act=ACT_RELU))
            x = o
            if lvl <= 4:
                z = nb(N, o.H, o.W, 1, torch.float32).view()
                plan.add(E.HeadConvOp(plan, 'predict_disp%d.0' % lvl, o, z))
                heads[lvl] = plan.add(E.HeadOp(plan, z, self.alpha, self.beta, slot.get(lvl), 1))
        plan.out_order = [heads[1].idx, heads[2].idx, heads[3].idx, heads[4].idx]

    def forward(self, x):
        outs = self._run([x])
        plan = self._plan_for([x])
        d = [outs[i] for i in plan.out_order]
        if self.training:
            return d[0], d[1], d[2], d[3]
        return d[0]
