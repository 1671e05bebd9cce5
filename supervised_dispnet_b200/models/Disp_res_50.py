"""Disp_res_50 -- hand-rolled ResNet-50 bottleneck encoder + 5-stage decoder (reference: models/Disp_res_50.py).

Same constructor, `init_weights`, `load_res_params`, 346 state_dict keys.  forward (:139-198) on
libdispnet_b200.so.  Quirk preserved: `bn1` is evaluated and discarded (`relu1 = relu(conv1)`, :143-145), so
it only updates its running statistics and its weight/bias never receive a gradient."""
import torch
import torch.nn as nn

from .. import engine as E
from ._common import (ACT_LRELU, ACT_NONE, ACT_RELU, LeakyReLU01, alpha_beta, conv_block, predict_disp, upconv_block,
                      xavier_init)


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)


class Bottleneck(nn.Module):
    """Parameter container with the reference's attribute names (:212-247); executed by the plan."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = conv1x1(inplanes, planes)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = conv3x3(planes, planes, stride)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = conv1x1(planes, planes * 4)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


_BLOCKS = [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]


class Disp_res_50(E.PlannedModule):

    def __init__(self, datasets='kitti'):
        super().__init__()
        self.alpha, self.beta = alpha_beta(datasets)
        self.only_train_dec = False
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        # bn1's output is discarded by the reference's forward (:143-145): its affine parameters never receive a gradient.  Frozen
        # like the dead VGG classifier so that DistributedDataParallel's reducer does not wait for them (buffers still update).
        for p in self.bn1.parameters():
            p.requires_grad_(False)
        self.relu = nn.ReLU(inplace=True)
        self.pool1 = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self.resblock(64, 3)
        self.layer2 = self.resblock(128, 4, stride=2)
        self.layer3 = self.resblock(256, 6, stride=2)
        self.layer4 = self.resblock(512, 3, stride=2)
        up = [512, 256, 128, 64, 32, 16]
        self.upconv5 = upconv_block(2048, up[1], 3, 1, 1, LeakyReLU01)
        self.upconv4 = upconv_block(up[1], up[2], 3, 1, 1, LeakyReLU01)
        self.upconv3 = upconv_block(up[2], up[3], 3, 1, 1, LeakyReLU01)
        self.upconv2 = upconv_block(up[3], up[4], 3, 1, 1, LeakyReLU01)
        self.upconv1 = upconv_block(up[4], up[5], 3, 1, 1, LeakyReLU01)
        self.iconv5 = conv_block(up[1] + 1024, up[1], 3, 1, LeakyReLU01)
        self.iconv4 = conv_block(up[2] + 512, up[2], 3, 1, LeakyReLU01)
        self.iconv3 = conv_block(1 + up[3] + 256, up[3], 3, 1, LeakyReLU01)
        self.iconv2 = conv_block(1 + up[4] + 64, up[4], 3, 1, LeakyReLU01)
        self.iconv1 = conv_block(1 + up[5], up[5], 3, 1, LeakyReLU01)
        self.predict_disp4 = predict_disp(up[2])
        self.predict_disp3 = predict_disp(up[3])
        self.predict_disp2 = predict_disp(up[4])
        self.predict_disp1 = predict_disp(up[5])

    def resblock(self, planes, num_blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * 4:
            downsample = nn.Sequential(conv1x1(self.inplanes, planes * 4, stride), nn.BatchNorm2d(planes * 4))
        layers = [Bottleneck(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * 4
        for _ in range(1, num_blocks):
            layers.append(Bottleneck(self.inplanes, planes))
        return nn.Sequential(*layers)

    def init_weights(self, use_pretrained_weights=False):
        xavier_init(self)
        if use_pretrained_weights:
            raise RuntimeError('pretrained ResNet weights must be supplied with load_res_params(state_dict); this '
                               'environment has no network access')

    def load_res_params(self, params):
        model_dict = self.state_dict()
        model_dict.update({k: v for k, v in params.items() if k in model_dict})
        self.load_state_dict(model_dict)

    def _build_plan(self, plan, shapes):
        N, Cin, H, W = shapes[0]
        assert Cin == 3 and H % 32 == 0 and W % 32 == 0
        nb = plan.new_buf
        inp = plan.add(E.InputOp(plan, shapes))
        up = [512, 256, 128, 64, 32, 16]
        cat5 = nb(N, H // 16, W // 16, up[1] + 1024)
        cat4 = nb(N, H // 8, W // 8, up[2] + 512)
        cat3 = nb(N, H // 4, W // 4, up[3] + 256 + 1)
        cat2 = nb(N, H // 2, W // 2, up[4] + 64 + 1)
        cat1 = nb(N, H, W, up[5] + 1)
        y1 = nb(N, H // 2, W // 2, 64).view()
        plan.add(E.ConvOp(plan, 'conv1', inp.out, y1, 7, stride=2, pad=3, bias=False, needs_dx=False))
        plan.add(E.BNOp(plan, 'bn1', y1, None))
        relu1 = cat2.view().channels(up[4], 64)
        plan.add(E.ActOp(plan, y1, relu1, ACT_RELU, needs_dx=False))
        x = nb(N, H // 4, W // 4, 64).view()
        plan.add(E.MaxPoolOp(plan, relu1, x, 3, 2, 1))
        layer_dst = [cat3.view().channels(up[3], 256), cat4.view().channels(up[2], 512), cat5.view().channels(up[1], 1024),
                     None]
        h, w = H // 4, W // 4
        for li, (pl, nblk, stride) in enumerate(_BLOCKS):
            for b in range(nblk):
                p = 'layer%d.%d.' % (li + 1, b)
                s = stride if b == 0 else 1
                ho, wo = h // s, w // s
                a1 = nb(N, h, w, pl).view()
                E.conv_bn(plan, p + 'conv1', p + 'bn1', x, (N, h, w, pl), a1, 1, ACT_RELU, pad=0, bias=False)
                a2 = nb(N, ho, wo, pl).view()
                E.conv_bn(plan, p + 'conv2', p + 'bn2', a1, (N, ho, wo, pl), a2, 3, ACT_RELU, stride=s, pad=1, bias=False)
                yc = nb(N, ho, wo, pl * 4).view()
                conv3 = plan.add(E.ConvOp(plan, p + 'conv3', a2, yc, 1, pad=0, bias=False, bn_follows=True))
                if b == 0:
                    idn = nb(N, ho, wo, pl * 4).view()
                    E.conv_bn(plan, p + 'downsample.0', p + 'downsample.1', x, (N, ho, wo, pl * 4), idn, 1, ACT_NONE, stride=s,
                              pad=0, bias=False)
                else:
                    idn = x
                last = b == nblk - 1
                out = layer_dst[li] if (last and layer_dst[li] is not None) else nb(N, ho, wo, pl * 4).view()
                plan.add(E.BNOp(plan, p + 'bn3', yc, out, ACT_RELU, residual=idn, conv=conv3))
                x = out
                h, w = ho, wo
        c5 = x
        plan.encoder_end = len(plan.ops)          # `only_train_dec` (reference :154-160) detaches relu1 .. conv5 here

        def upc(name, src, dst):
            plan.add(E.ConvOp(plan, name + '.0', src, dst, 3, stride=2, pad=1, transposed=True, act=ACT_LRELU))

        def iconv(name, cat, cout):
            o = nb(N, cat.H, cat.W, cout).view()
            plan.add(E.ConvOp(plan, name + '.0', cat.view(), o, 3, act=ACT_LRELU))
            return o

        def head(name, src, up_view):
            z = nb(N, src.H, src.W, 1, torch.float32).view()
            plan.add(E.HeadConvOp(plan, name + '.0', src, z))
            return plan.add(E.HeadOp(plan, z, self.alpha, self.beta, up_view, 0))

        upc('upconv5', c5, cat5.view().channels(0, up[1]))
        i5 = iconv('iconv5', cat5, up[1])
        upc('upconv4', i5, cat4.view().channels(0, up[2]))
        i4 = iconv('iconv4', cat4, up[2])
        d4 = head('predict_disp4', i4, cat3.view().channels(up[3] + 256, 1))
        upc('upconv3', i4, cat3.view().channels(0, up[3]))
        i3 = iconv('iconv3', cat3, up[3])
        d3 = head('predict_disp3', i3, cat2.view().channels(up[4] + 64, 1))
        upc('upconv2', i3, cat2.view().channels(0, up[4]))
        i2 = iconv('iconv2', cat2, up[4])
        d2 = head('predict_disp2', i2, cat1.view().channels(up[5], 1))
        upc('upconv1', i2, cat1.view().channels(0, up[5]))
        i1 = iconv('iconv1', cat1, up[5])
        d1 = head('predict_disp1', i1, None)
        plan.out_order = [d1.idx, d2.idx, d3.idx, d4.idx]

    def forward(self, x):
        outs = self._run([x])
        plan = self._plan_for([x])
        d = [outs[i] for i in plan.out_order]
        if self.training:
            return d[0], d[1], d[2], d[3]
        return d[0]
