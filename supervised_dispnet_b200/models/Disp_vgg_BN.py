"""Disp_vgg_BN -- VGG16-BN encoder + 5-stage up-convolutional decoder (reference: models/Disp_vgg_BN.py).

Same constructor, attributes (`alpha`, `beta`, `only_train_dec`), `init_weights`, `load_vgg_params` and the
same 125 state_dict keys (`features.features.N.*`, `features.classifier.N.*`, `upconvK.0.*`, `iconvK.0.*`,
`dispK.0.*`) as the reference, so its checkpoints load unchanged (train.py:281).  forward() (reference
:136-191) is executed by libdispnet_b200.so: 13x conv3x3 + BatchNorm(train) + ReLU with the five 2x2 max-pools
fused into the normalise pass, 5x ConvTranspose2d(k4,s2,p1) as four 2x2-tap phase convolutions writing
straight into the iconv input buffers (no torch.cat), 5x iconv + LeakyReLU(0.1), 4x alpha*sigmoid+beta heads
whose nearest x2 up-sampling lands in the next iconv's input slot.
"""
import torch
import torch.nn as nn

from .. import engine as E
from ._common import (ACT_LRELU, ACT_NONE, ACT_RELU, LeakyReLU01, alpha_beta, conv_block, predict_disp, upconv_block,
                      xavier_init)

_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']
_BLOCKS = [2, 2, 3, 3, 3]          # convs per block as sliced at reference :137-141


class _VGG16BN(nn.Module):
    """Container with torchvision.models.vgg16_bn's module tree (features / avgpool / classifier) so the keys
    match `models.vgg16_bn(pretrained=False)` (reference :84).  The classifier is dead weight there too
    (123.6 M parameters that never receive a gradient); it is kept for checkpoint compatibility and frozen so
    DistributedDataParallel does not wait for it."""

    def __init__(self):
        super().__init__()
        layers, c_in = [], 3
        for v in _CFG:
            if v == 'M':
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(c_in, v, 3, padding=1), nn.BatchNorm2d(v), nn.ReLU(inplace=True)]
                c_in = v
        self.features = nn.Sequential(*layers)
        self.avgpool = nn.AdaptiveAvgPool2d((7, 7))
        self.classifier = nn.Sequential(nn.Linear(512 * 7 * 7, 4096), nn.ReLU(True), nn.Dropout(),
                                        nn.Linear(4096, 4096), nn.ReLU(True), nn.Dropout(), nn.Linear(4096, 1000))
        for p in self.classifier.parameters():
            p.requires_grad_(False)


class Disp_vgg_BN(E.PlannedModule):

    def __init__(self, datasets='kitti'):
        super().__init__()
        self.only_train_dec = False
        self.alpha, self.beta = alpha_beta(datasets)
        self.features = _VGG16BN()
        self.upconv4 = upconv_block(512, 256, 4, 1, 0, LeakyReLU01)
        self.iconv4 = conv_block(256 + 512, 256, 3, 1, LeakyReLU01)
        self.upconv3 = upconv_block(256, 128, 4, 1, 0, LeakyReLU01)
        self.iconv3 = conv_block(128 + 256, 128, 3, 1, LeakyReLU01)
        self.upconv2 = upconv_block(128, 64, 4, 1, 0, LeakyReLU01)
        self.iconv2 = conv_block(64 + 128 + 1, 64, 3, 1, LeakyReLU01)
        self.upconv1 = upconv_block(64, 32, 4, 1, 0, LeakyReLU01)
        self.iconv1 = conv_block(32 + 64 + 1, 32, 3, 1, LeakyReLU01)
        self.upconv0 = upconv_block(32, 16, 4, 1, 0, LeakyReLU01)
        self.iconv0 = conv_block(16 + 1, 16, 3, 1, LeakyReLU01)
        self.disp3 = predict_disp(128)
        self.disp2 = predict_disp(64)
        self.disp1 = predict_disp(32)
        self.disp0 = predict_disp(16)

    def init_weights(self, use_pretrained_weights=False):
        xavier_init(self)
        if use_pretrained_weights:
            raise RuntimeError('pretrained VGG weights must be supplied with load_vgg_params(state_dict); this '
                               'environment has no network access')

    def load_vgg_params(self, params):
        model_dict = self.features.state_dict()
        model_dict.update({k: v for k, v in params.items() if k in model_dict})
        self.features.load_state_dict(model_dict)

    # ---- plan: reference forward :136-191 as a static op list
    def _build_plan(self, plan, shapes):
        N, Cin, H, W = shapes[0]
        assert Cin == 3 and H % 32 == 0 and W % 32 == 0, 'Disp_vgg_BN needs 3xHxW input with H, W multiples of 32'
        nb = plan.new_buf
        inp = plan.add(E.InputOp(plan, shapes))
        planes = [64, 128, 256, 512, 512]
        up_planes = [256, 128, 64, 32, 16]
        # decoder input buffers: [upconv | skip | upsampled disparity]
        cat4 = nb(N, H // 16, W // 16, 256 + 512)
        cat3 = nb(N, H // 8, W // 8, 128 + 256)
        cat2 = nb(N, H // 4, W // 4, 64 + 128 + 1)
        cat1 = nb(N, H // 2, W // 2, 32 + 64 + 1)
        cat0 = nb(N, H, W, 16 + 1)
        skip_dst = [cat1.view().channels(32, 64), cat2.view().channels(64, 128), cat3.view().channels(128, 256),
                    cat4.view().channels(256, 512), None]
        x = inp.out
        h, w = H, W
        conv_idx = [i for i, v in enumerate(self.features.features) if isinstance(v, nn.Conv2d)]
        k = 0
        for b, nconv in enumerate(_BLOCKS):
            for j in range(nconv):
                ci = conv_idx[k]
                yshape = (N, h, w, planes[b])
                last = j == nconv - 1
                if last:
                    h, w = h // 2, w // 2
                    out = skip_dst[b] if skip_dst[b] is not None else nb(N, h, w, planes[b]).view()
                else:
                    out = nb(N, h, w, planes[b]).view()
                E.conv_bn(plan, 'features.features.%d' % ci, 'features.features.%d' % (ci + 1), x, yshape, out, 3, ACT_RELU,
                          pool=last, needs_dx=(k > 0))
                x = out
                k += 1
        c5 = x
        plan.encoder_end = len(plan.ops)          # `only_train_dec` (reference :148-153) detaches conv1..conv5 here

        def up(name, src, dst):
            plan.add(E.ConvOp(plan, name + '.0', src, dst, 4, stride=2, pad=1, transposed=True, act=ACT_LRELU))

        def iconv(name, cat, cout):
            o = nb(N, cat.H, cat.W, cout).view()
            plan.add(E.ConvOp(plan, name + '.0', cat.view(), o, 3, act=ACT_LRELU))
            return o

        def head(name, src, up_view):
            z = nb(N, src.H, src.W, 1, torch.float32).view()
            plan.add(E.HeadConvOp(plan, name + '.0', src, z))
            return plan.add(E.HeadOp(plan, z, self.alpha, self.beta, up_view, 0))

        up('upconv4', c5, cat4.view().channels(0, 256))
        i4 = iconv('iconv4', cat4, 256)
        up('upconv3', i4, cat3.view().channels(0, 128))
        i3 = iconv('iconv3', cat3, 128)
        d3 = head('disp3', i3, cat2.view().channels(192, 1))
        up('upconv2', i3, cat2.view().channels(0, 64))
        i2 = iconv('iconv2', cat2, 64)
        d2 = head('disp2', i2, cat1.view().channels(96, 1))
        up('upconv1', i2, cat1.view().channels(0, 32))
        i1 = iconv('iconv1', cat1, 32)
        d1 = head('disp1', i1, cat0.view().channels(16, 1))
        up('upconv0', i1, cat0.view().channels(0, 16))
        i0 = iconv('iconv0', cat0, 16)
        d0 = head('disp0', i0, None)
        plan.out_order = [d0.idx, d1.idx, d2.idx, d3.idx]

    def forward(self, x):
        outs = self._run([x])
        plan = self._plan_for([x])
        d = [outs[i] for i in plan.out_order]
        if self.training:
            return d[0], d[1], d[2], d[3]
        return d[0]
