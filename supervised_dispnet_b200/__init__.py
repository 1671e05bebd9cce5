"""supervised_dispnet_b200 -- B200-native (sm_100a) implementation of the DispNet-family training hot path of
zenithfang/supervised_dispnet behind the reference's own Python surface:

    from supervised_dispnet_b200 import models, loss_functions
    from supervised_dispnet_b200.inverse_warp import inverse_warp

All arithmetic runs in libdispnet_b200.so (hand-written CUDA; see include/dispnet_b200.h); PyTorch provides
device memory, streams, autograd glue and torch.distributed only.
"""
from . import _lib  # noqa: F401
from . import engine  # noqa: F401
from . import models  # noqa: F401
from . import loss_functions  # noqa: F401
from . import inverse_warp  # noqa: F401
from . import layers  # noqa: F401
from . import custom_transforms  # noqa: F401
from . import utils  # noqa: F401

__version__ = '0.1.0'
