"""Drop-in `models` package: the four hot-path networks of zenithfang/supervised_dispnet
(`models/__init__.py:1,2,6,13` of the reference) as nn.Modules with identical constructor arguments,
attributes and state_dict keys, whose forward/backward run on libdispnet_b200.so."""
from .DispNetS import DispNetS
from .PoseExpNet import PoseExpNet
from .Disp_vgg_BN import Disp_vgg_BN
from .Disp_res_50 import Disp_res_50

__all__ = ['DispNetS', 'PoseExpNet', 'Disp_vgg_BN', 'Disp_res_50']
