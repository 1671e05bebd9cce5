"""Parameter containers shared by the model classes.

The nn.Conv2d / nn.ConvTranspose2d / nn.BatchNorm2d objects below are *containers only*: they give the
modules the reference's state_dict keys, make `init_weights`, `.apply(fn)`, `.to(device)` and DDP behave as
they do on the reference, and are never called -- the arithmetic runs in libdispnet_b200.so via engine.Plan.
"""
import torch
import torch.nn as nn

from .. import _lib as L


def alpha_beta(datasets):
    if datasets == 'kitti':
        return 10, 0.01
    if datasets == 'nyu':
        return 10, 0.1
    return None, None


def conv_block(c_in, c_out, k, stride, act):
    """nn.Sequential(conv, act) like Conv2dBlock1 / conv() / downsample_conv halves of the reference."""
    return nn.Sequential(nn.Conv2d(c_in, c_out, k, stride, (k - 1) // 2), act())


def upconv_block(c_in, c_out, k, pad, out_pad, act):
    return nn.Sequential(nn.ConvTranspose2d(c_in, c_out, k, 2, pad, out_pad), act())


def predict_disp(c_in):
    return nn.Sequential(nn.Conv2d(c_in, 1, kernel_size=3, padding=1), nn.Sigmoid())


class LeakyReLU01(nn.LeakyReLU):
    def __init__(self):
        super().__init__(0.1)


def xavier_init(module, with_linear=True):
    """`init_weights` of the reference (e.g. models/Disp_vgg_BN.py:112-120): xavier_uniform on conv / convT
    (/ linear) weights in modules() order, zero biases; BatchNorm keeps its defaults."""
    kinds = (nn.Conv2d, nn.ConvTranspose2d, nn.Linear) if with_linear else (nn.Conv2d, nn.ConvTranspose2d)
    for m in module.modules():
        if isinstance(m, kinds):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                torch.nn.init.constant_(m.bias, 0)


ACT_NONE, ACT_RELU, ACT_LRELU = L.ACT_NONE, L.ACT_RELU, L.ACT_LRELU
