"""PoseExpNet (reference: models/PoseExpNet.py) -- same constructor (`nb_ref_imgs`, `output_exp`),
`init_weights()`, state_dict keys (`convK.0.*`, `pose_pred.*`, `upconvK.0.*`, `predict_maskK.*`) and return
convention (:92-95).  The channel concat of target + reference frames (:60-62) is done by the input packer;
pose = 0.01 * spatial mean of the 1x1 pose_pred conv; explainability masks are sigmoid heads on the cropped
ConvTranspose2d(k4,s2,p1) decoder."""
import torch
import torch.nn as nn

from .. import engine as E
from ._common import ACT_NONE, ACT_RELU, upconv_block, xavier_init


def _conv(c_in, c_out, k=3):
    return nn.Sequential(nn.Conv2d(c_in, c_out, kernel_size=k, padding=(k - 1) // 2, stride=2), nn.ReLU(inplace=True))


class PoseExpNet(E.PlannedModule):

    def __init__(self, nb_ref_imgs=2, output_exp=False):
        super().__init__()
        self.nb_ref_imgs = nb_ref_imgs
        self.output_exp = output_exp
        cp = [16, 32, 64, 128, 256, 256, 256]
        ks = [7, 5, 3, 3, 3, 3, 3]
        c_in = 3 * (1 + nb_ref_imgs)
        for i, (c, k) in enumerate(zip(cp, ks)):
            setattr(self, 'conv%d' % (i + 1), _conv(c_in, c, k))
            c_in = c
        self.pose_pred = nn.Conv2d(cp[6], 6 * nb_ref_imgs, kernel_size=1, padding=0)
        if output_exp:
            up = [256, 128, 64, 32, 16]
            c_in = cp[4]
            for i, c in zip(range(5, 0, -1), up):
                setattr(self, 'upconv%d' % i, upconv_block(c_in, c, 4, 1, 0, nn.ReLU))
                c_in = c
            self.predict_mask4 = nn.Conv2d(up[1], nb_ref_imgs, kernel_size=3, padding=1)
            self.predict_mask3 = nn.Conv2d(up[2], nb_ref_imgs, kernel_size=3, padding=1)
            self.predict_mask2 = nn.Conv2d(up[3], nb_ref_imgs, kernel_size=3, padding=1)
            self.predict_mask1 = nn.Conv2d(up[4], nb_ref_imgs, kernel_size=3, padding=1)

    def init_weights(self):
        xavier_init(self, with_linear=False)

    def _build_plan(self, plan, shapes):
        N, _, H, W = shapes[0]
        nb = plan.new_buf
        inp = plan.add(E.InputOp(plan, shapes))
        cp = [16, 32, 64, 128, 256, 256, 256]
        ks = [7, 5, 3, 3, 3, 3, 3]
        x = inp.out
        convs = []
        h, w = H, W
        for i, (c, k) in enumerate(zip(cp, ks)):
            p = (k - 1) // 2
            h, w = (h + 2 * p - k) // 2 + 1, (w + 2 * p - k) // 2 + 1
            o = nb(N, h, w, c).view()
            plan.add(E.ConvOp(plan, 'conv%d.0' % (i + 1), x, o, k, stride=2, act=ACT_RELU, needs_dx=(i > 0)))
            convs.append(o)
            x = o
        R = self.nb_ref_imgs
        z = nb(N, h, w, 6 * R, torch.float32).view()
        plan.add(E.ConvOp(plan, 'pose_pred', x, z, 1, pad=0, act=ACT_NONE))
        pose = plan.add(E.MeanOutOp(plan, z, 0.01))
        masks = {}
        if self.output_exp:
            up = [256, 128, 64, 32, 16]
            sizes = [(convs[3].H, convs[3].W), (convs[2].H, convs[2].W), (convs[1].H, convs[1].W), (convs[0].H, convs[0].W),
                     (H, W)]
            x = convs[4]
            for lvl, c, (hh, ww) in zip(range(5, 0, -1), up, sizes):
                o = nb(N, hh, ww, c).view()          # cropped region only (reference :76-80)
                plan.add(E.ConvOp(plan, 'upconv%d.0' % lvl, x, o, 4, stride=2, pad=1, transposed=True, act=ACT_RELU))
                x = o
                if lvl <= 4:
                    zm = nb(N, hh, ww, R, torch.float32).view()
                    plan.add(E.ConvOp(plan, 'predict_mask%d' % lvl, o, zm, 3, act=ACT_NONE))
                    masks[lvl] = plan.add(E.SigmoidOutOp(plan, zm))
        plan.pose_idx = pose.idx
        plan.mask_idx = [masks[l].idx for l in (1, 2, 3, 4)] if masks else None

    def forward(self, target_image, ref_imgs):
        assert len(ref_imgs) == self.nb_ref_imgs
        inputs = [target_image] + list(ref_imgs)
        outs = self._run(inputs)
        plan = self._plan_for(inputs)
        pose = outs[plan.pose_idx].view(target_image.size(0), self.nb_ref_imgs, 6)
        if plan.mask_idx is not None:
            m1, m2, m3, m4 = [outs[i] for i in plan.mask_idx]
        else:
            m1 = m2 = m3 = m4 = None
        if self.training:
            return [m1, m2, m3, m4], pose
        return m1, pose
