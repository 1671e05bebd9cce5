"""Oracle (TEST INFRASTRUCTURE ONLY): functional fp32 restatement of the four hot-path networks.

Each network is a pure function ``f(sd, x, training)`` of a ``state_dict``-shaped dict ``sd`` whose keys
and shapes are exactly the reference's, so reference checkpoints drive it directly
and torch autograd through it is the gradient oracle.

Reference call sites restated here:
  * Disp_vgg_BN.forward      models/Disp_vgg_BN.py:136-191   (ctor :72-110, init :112-125)
  * DispNetS.forward         models/DispNetS.py:93-140       (helpers :7-39, init :86-91)
  * PoseExpNet.forward       models/PoseExpNet.py:58-95      (init :51-56)
  * Disp_res_50.forward      models/Disp_res_50.py:139-198   (Bottleneck :212-247, resblock :98-113)
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# parameter tables (name, shape) in the order the reference's ``self.modules()`` walk meets them, which
# is the order ``init_weights`` consumes the RNG in.
# --------------------------------------------------------------------------------------------------

VGG_CONV_IDX = [0, 3, 7, 10, 14, 17, 20, 24, 27, 30, 34, 37, 40]          # torchvision vgg16_bn.features
VGG_BN_IDX = [i + 1 for i in VGG_CONV_IDX]
VGG_PLANES = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256),
              (256, 512), (512, 512), (512, 512), (512, 512), (512, 512), (512, 512)]
# blocks as sliced at models/Disp_vgg_BN.py:137-141 -> number of convs per block, each block ends in a 2x2 pool
VGG_BLOCKS = [2, 2, 3, 3, 3]


def _conv_entry(name, cout, cin, k, bias=True, transposed=False):
    shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    return dict(name=name, shape=shape, bias=cout if bias else None, kind='conv')


def disp_vgg_bn_layout():
    """[(entry)] in reference module order: features convs/BNs, classifier linears, decoder."""
    ents = []
    for idx, (ci, co) in zip(VGG_CONV_IDX, VGG_PLANES):
        ents.append(_conv_entry('features.features.%d' % idx, co, ci, 3))
        ents.append(dict(name='features.features.%d' % (idx + 1), kind='bn', c=co))
    for i, (fin, fout) in zip((0, 3, 6), ((512 * 7 * 7, 4096), (4096, 4096), (4096, 1000))):
        ents.append(dict(name='features.classifier.%d' % i, kind='linear', shape=(fout, fin), bias=fout))
    dec = [('upconv4', 512, 256, True), ('iconv4', 768, 256, False), ('upconv3', 256, 128, True),
           ('iconv3', 384, 128, False), ('upconv2', 128, 64, True), ('iconv2', 193, 64, False),
           ('upconv1', 64, 32, True), ('iconv1', 97, 32, False), ('upconv0', 32, 16, True),
           ('iconv0', 17, 16, False)]
    for name, ci, co, tr in dec:
        ents.append(_conv_entry(name + '.0', co, ci, 4 if tr else 3, transposed=tr))
    for name, ci in (('disp3', 128), ('disp2', 64), ('disp1', 32), ('disp0', 16)):
        ents.append(_conv_entry(name + '.0', 1, ci, 3))
    return ents


def dispnets_layout():
    ents = []
    planes = [32, 64, 128, 256, 512, 512, 512]
    ks = [7, 5, 3, 3, 3, 3, 3]
    cin = 3
    for i, (co, k) in enumerate(zip(planes, ks)):
        ents.append(_conv_entry('conv%d.0' % (i + 1), co, cin, k))
        ents.append(_conv_entry('conv%d.2' % (i + 1), co, co, k))
        cin = co
    up = [512, 512, 256, 128, 64, 32, 16]
    cin = 512
    for i, co in zip(range(7, 0, -1), up):
        ents.append(_conv_entry('upconv%d.0' % i, co, cin, 3, transposed=True))
        cin = co
    iin = [512 + 512, 512 + 512, 256 + 256, 128 + 128, 1 + 64 + 64, 1 + 32 + 32, 1 + 16]
    for i, ci, co in zip(range(7, 0, -1), iin, up):
        ents.append(_conv_entry('iconv%d.0' % i, co, ci, 3))
    for i, ci in zip((4, 3, 2, 1), (128, 64, 32, 16)):
        ents.append(_conv_entry('predict_disp%d.0' % i, 1, ci, 3))
    return ents


def poseexpnet_layout(nb_ref_imgs=2, output_exp=False):
    ents = []
    planes = [16, 32, 64, 128, 256, 256, 256]
    ks = [7, 5, 3, 3, 3, 3, 3]
    cin = 3 * (1 + nb_ref_imgs)
    for i, (co, k) in enumerate(zip(planes, ks)):
        ents.append(_conv_entry('conv%d.0' % (i + 1), co, cin, k))
        cin = co
    ents.append(_conv_entry('pose_pred', 6 * nb_ref_imgs, 256, 1))
    if output_exp:
        up = [256, 128, 64, 32, 16]
        cin = 256
        for i, co in zip(range(5, 0, -1), up):
            ents.append(_conv_entry('upconv%d.0' % i, co, cin, 4, transposed=True))
            cin = co
        for i, ci in zip((4, 3, 2, 1), (128, 64, 32, 16)):
            ents.append(_conv_entry('predict_mask%d' % i, nb_ref_imgs, ci, 3))
    return ents


RES50_BLOCKS = [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]


def disp_res_50_layout():
    ents = [_conv_entry('conv1', 64, 3, 7, bias=False), dict(name='bn1', kind='bn', c=64)]
    inpl = 64
    for li, (pl, nb, stride) in enumerate(RES50_BLOCKS):
        for b in range(nb):
            p = 'layer%d.%d.' % (li + 1, b)
            ents.append(_conv_entry(p + 'conv1', pl, inpl, 1, bias=False))
            ents.append(dict(name=p + 'bn1', kind='bn', c=pl))
            ents.append(_conv_entry(p + 'conv2', pl, pl, 3, bias=False))
            ents.append(dict(name=p + 'bn2', kind='bn', c=pl))
            ents.append(_conv_entry(p + 'conv3', pl * 4, pl, 1, bias=False))
            ents.append(dict(name=p + 'bn3', kind='bn', c=pl * 4))
            if b == 0:   # stride != 1 or inplanes != planes*4  -> always true for block 0 (resblock :98-113)
                ents.append(_conv_entry(p + 'downsample.0', pl * 4, inpl, 1, bias=False))
                ents.append(dict(name=p + 'downsample.1', kind='bn', c=pl * 4))
            inpl = pl * 4
    up = [256, 128, 64, 32, 16]
    cin = 2048
    for i, co in zip(range(5, 0, -1), up):
        ents.append(_conv_entry('upconv%d.0' % i, co, cin, 3, transposed=True))
        cin = co
    iin = [256 + 1024, 128 + 512, 1 + 64 + 256, 1 + 32 + 64, 1 + 16]
    for i, ci, co in zip(range(5, 0, -1), iin, up):
        ents.append(_conv_entry('iconv%d.0' % i, co, ci, 3))
    for i, ci in zip((4, 3, 2, 1), (128, 64, 32, 16)):
        ents.append(_conv_entry('predict_disp%d.0' % i, 1, ci, 3))
    return ents


LAYOUTS = {
    'Disp_vgg_BN': disp_vgg_bn_layout,
    'DispNetS': dispnets_layout,
    'PoseExpNet': poseexpnet_layout,
    'Disp_res_50': disp_res_50_layout,
}


def init_state_dict(model, seed=0, skip_dead=False, **kw):
    """Reproduce ``torch.manual_seed(seed); net.init_weights()`` of the reference without the reference.

    ``init_weights`` (models/Disp_vgg_BN.py:112-120, DispNetS.py:86-91, PoseExpNet.py:51-56,
    Disp_res_50.py:115-123) xavier-uniforms every conv / convT / linear weight in ``self.modules()`` order
    and zeroes the biases; BatchNorm keeps its constructor default (gamma=1, beta=0, running 0/1).
    ``skip_dead=True`` still *consumes* the RNG for the dead VGG classifier but stores zeros-free tiny
    placeholders so small tests do not hold 0.5 GB.
    """
    torch.manual_seed(seed)
    sd = OrderedDict()
    for e in LAYOUTS[model](**kw):
        n = e['name']
        if e['kind'] == 'bn':
            c = e['c']
            sd[n + '.weight'] = torch.ones(c)
            sd[n + '.bias'] = torch.zeros(c)
            sd[n + '.running_mean'] = torch.zeros(c)
            sd[n + '.running_var'] = torch.ones(c)
            sd[n + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
            continue
        w = torch.empty(e['shape'])
        torch.nn.init.xavier_uniform_(w)
        if e['kind'] == 'linear' and skip_dead:
            continue
        sd[n + '.weight'] = w
        if e['bias'] is not None:
            sd[n + '.bias'] = torch.zeros(e['bias'])
    return sd


# --------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------

def _bn(sd, name, x, training):
    """nn.BatchNorm2d: batch statistics + running update (momentum .1, eps 1e-5) in training, running
    statistics in eval.  Mutates the running buffers in ``sd`` exactly as the module does."""
    if training and (name + '.num_batches_tracked') in sd:
        sd[name + '.num_batches_tracked'] += 1
    return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'],
                        sd[name + '.weight'], sd[name + '.bias'], training, 0.1, 1e-5)


def _conv(sd, name, x, stride=1, pad=None):
    w = sd[name + '.weight']
    if pad is None:
        pad = (w.shape[2] - 1) // 2
    return F.conv2d(x, w, sd.get(name + '.bias'), stride, pad)


def _convT(sd, name, x, pad=1, out_pad=0):
    return F.conv_transpose2d(x, sd[name + '.weight'], sd.get(name + '.bias'), 2, pad, out_pad)


def _up_nearest(x):
    """F.upsample(scale_factor=2, mode='nearest') (models/Disp_vgg_BN.py:10-11): out[2i+a,2j+b]=in[i,j]."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def _up_bilinear(x):
    """F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) (models/DispNetS.py:120)."""
    return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)


def _crop_like(x, ref):
    return x[:, :, :ref.size(2), :ref.size(3)]


def _head(sd, name, x, alpha, beta):
    return alpha * torch.sigmoid(_conv(sd, name, x)) + beta


def alpha_beta(datasets):
    return (10, 0.01) if datasets == 'kitti' else (10, 0.1)


# --------------------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------------------

def disp_vgg_bn(sd, x, training=True, datasets='kitti'):
    alpha, beta = alpha_beta(datasets)
    skips = []
    k = 0
    for nconv in VGG_BLOCKS:
        for _ in range(nconv):
            x = _conv(sd, 'features.features.%d' % VGG_CONV_IDX[k], x)
            x = F.relu(_bn(sd, 'features.features.%d' % VGG_BN_IDX[k], x, training))
            k += 1
        x = F.max_pool2d(x, 2, 2)
        skips.append(x)
    c1, c2, c3, c4, c5 = skips
    lrelu = lambda t: F.leaky_relu(t, 0.1)
    up4 = lrelu(_convT(sd, 'upconv4.0', c5))
    i4 = lrelu(_conv(sd, 'iconv4.0', torch.cat((up4, c4), 1)))
    up3 = lrelu(_convT(sd, 'upconv3.0', i4))
    i3 = lrelu(_conv(sd, 'iconv3.0', torch.cat((up3, c3), 1)))
    d3 = _head(sd, 'disp3.0', i3, alpha, beta)
    up2 = lrelu(_convT(sd, 'upconv2.0', i3))
    i2 = lrelu(_conv(sd, 'iconv2.0', torch.cat((up2, c2, _up_nearest(d3)), 1)))
    d2 = _head(sd, 'disp2.0', i2, alpha, beta)
    up1 = lrelu(_convT(sd, 'upconv1.0', i2))
    i1 = lrelu(_conv(sd, 'iconv1.0', torch.cat((up1, c1, _up_nearest(d2)), 1)))
    d1 = _head(sd, 'disp1.0', i1, alpha, beta)
    up0 = lrelu(_convT(sd, 'upconv0.0', i1))
    i0 = lrelu(_conv(sd, 'iconv0.0', torch.cat((up0, _up_nearest(d1)), 1)))
    d0 = _head(sd, 'disp0.0', i0, alpha, beta)
    return (d0, d1, d2, d3) if training else d0


def dispnets(sd, x, training=True, datasets='kitti'):
    alpha, beta = alpha_beta(datasets)
    outs = []
    t = x
    for i in range(1, 8):
        t = F.relu(_conv(sd, 'conv%d.0' % i, t, stride=2))
        t = F.relu(_conv(sd, 'conv%d.2' % i, t))
        outs.append(t)
    c1, c2, c3, c4, c5, c6, c7 = outs
    up = lambda n, t, ref: _crop_like(F.relu(_convT(sd, n, t, 1, 1)), ref)
    ic = lambda n, t: F.relu(_conv(sd, n, t))
    i7 = ic('iconv7.0', torch.cat((up('upconv7.0', c7, c6), c6), 1))
    i6 = ic('iconv6.0', torch.cat((up('upconv6.0', i7, c5), c5), 1))
    i5 = ic('iconv5.0', torch.cat((up('upconv5.0', i6, c4), c4), 1))
    i4 = ic('iconv4.0', torch.cat((up('upconv4.0', i5, c3), c3), 1))
    d4 = _head(sd, 'predict_disp4.0', i4, alpha, beta)
    i3 = ic('iconv3.0', torch.cat((up('upconv3.0', i4, c2), c2, _crop_like(_up_bilinear(d4), c2)), 1))
    d3 = _head(sd, 'predict_disp3.0', i3, alpha, beta)
    i2 = ic('iconv2.0', torch.cat((up('upconv2.0', i3, c1), c1, _crop_like(_up_bilinear(d3), c1)), 1))
    d2 = _head(sd, 'predict_disp2.0', i2, alpha, beta)
    i1 = ic('iconv1.0', torch.cat((up('upconv1.0', i2, x), _crop_like(_up_bilinear(d2), x)), 1))
    d1 = _head(sd, 'predict_disp1.0', i1, alpha, beta)
    return (d1, d2, d3, d4) if training else d1


def poseexpnet(sd, target_image, ref_imgs, training=True, output_exp=False):
    nb = len(ref_imgs)
    inp = torch.cat([target_image] + list(ref_imgs), 1)
    t = inp
    outs = []
    for i in range(1, 8):
        t = F.relu(_conv(sd, 'conv%d.0' % i, t, stride=2))
        outs.append(t)
    c1, c2, c3, c4, c5, c6, c7 = outs
    pose = _conv(sd, 'pose_pred', c7, pad=0)
    pose = pose.mean(3).mean(2)
    pose = 0.01 * pose.view(pose.size(0), nb, 6)
    if output_exp:
        upc = lambda n, t, ref: F.relu(_convT(sd, n, t))[:, :, :ref.size(2), :ref.size(3)]
        u5 = upc('upconv5.0', c5, c4)
        u4 = upc('upconv4.0', u5, c3)
        u3 = upc('upconv3.0', u4, c2)
        u2 = upc('upconv2.0', u3, c1)
        u1 = upc('upconv1.0', u2, inp)
        m4 = torch.sigmoid(_conv(sd, 'predict_mask4', u4))
        m3 = torch.sigmoid(_conv(sd, 'predict_mask3', u3))
        m2 = torch.sigmoid(_conv(sd, 'predict_mask2', u2))
        m1 = torch.sigmoid(_conv(sd, 'predict_mask1', u1))
    else:
        m1 = m2 = m3 = m4 = None
    if training:
        return [m1, m2, m3, m4], pose
    return m1, pose


def _bottleneck(sd, p, x, stride, has_down, training):
    out = F.relu(_bn(sd, p + 'bn1', _conv(sd, p + 'conv1', x, pad=0), training))
    out = F.relu(_bn(sd, p + 'bn2', _conv(sd, p + 'conv2', out, stride=stride, pad=1), training))
    out = _bn(sd, p + 'bn3', _conv(sd, p + 'conv3', out, pad=0), training)
    idn = x
    if has_down:
        idn = _bn(sd, p + 'downsample.1', _conv(sd, p + 'downsample.0', x, stride=stride, pad=0), training)
    return F.relu(out + idn)


def disp_res_50(sd, x, training=True, datasets='kitti'):
    alpha, beta = alpha_beta(datasets)
    conv1 = _conv(sd, 'conv1', x, stride=2, pad=3)
    _bn(sd, 'bn1', conv1, training)          # evaluated and discarded (models/Disp_res_50.py:143-145)
    relu1 = F.relu(conv1)
    t = F.max_pool2d(relu1, 3, 2, 1)
    feats = []
    for li, (pl, nb, stride) in enumerate(RES50_BLOCKS):
        for b in range(nb):
            t = _bottleneck(sd, 'layer%d.%d.' % (li + 1, b), t, stride if b == 0 else 1, b == 0, training)
        feats.append(t)
    c2, c3, c4, c5 = feats
    lrelu = lambda t: F.leaky_relu(t, 0.1)
    up = lambda n, t: lrelu(_convT(sd, n, t, 1, 1))
    ic = lambda n, t: lrelu(_conv(sd, n, t))
    i5 = ic('iconv5.0', torch.cat((up('upconv5.0', c5), c4), 1))
    i4 = ic('iconv4.0', torch.cat((up('upconv4.0', i5), c3), 1))
    d4 = _head(sd, 'predict_disp4.0', i4, alpha, beta)
    i3 = ic('iconv3.0', torch.cat((up('upconv3.0', i4), c2, _up_nearest(d4)), 1))
    d3 = _head(sd, 'predict_disp3.0', i3, alpha, beta)
    i2 = ic('iconv2.0', torch.cat((up('upconv2.0', i3), relu1, _up_nearest(d3)), 1))
    d2 = _head(sd, 'predict_disp2.0', i2, alpha, beta)
    i1 = ic('iconv1.0', torch.cat((up('upconv1.0', i2), _up_nearest(d2)), 1))
    d1 = _head(sd, 'predict_disp1.0', i1, alpha, beta)
    return (d1, d2, d3, d4) if training else d1


FORWARDS = {'Disp_vgg_BN': disp_vgg_bn, 'DispNetS': dispnets, 'Disp_res_50': disp_res_50}
