"""Import shims for the UNMODIFIED reference checkout (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference (zenithfang/supervised_dispnet @ f81dfcc) is plain Python on top of torch; it has no setup.py, so it cannot be
pip-installed (`pip install /root/reference` fails: "neither setup.py nor pyproject.toml").  What the GPU box needs of it
(train.py, loss_functions.py, inverse_warp.py, layers.py, utils.py, logger.py, custom_transforms.py, models/, networks/,
datasets/) is staged by `stage()` into the git-ignored `baseline/_ref/` -- never into the
repository history.  `import_reference(root)` makes `import train` work under the container's torch 2.11 with the shims of
SURVEY.md 8(c): `scipy.misc.imresize`, stub modules for the CLI-only dependencies (`path`, `blessings`, `progressbar`,
`tensorboardX`, `imageio`, `skimage.transform`), and the reference's `__init__`-less `datasets/` directory bound over the
HuggingFace `datasets` package that shadows it in this image.

Users: tests/ (the unchanged-`train.train` drop-in test, fixture generation) and bench.py's `gpu_baseline` leg (the
reference's own modules on the B200 through PyTorch/cuDNN -- the kernel to beat).  Nothing under supervised_dispnet_b200/
imports this file.
"""
import importlib
import os
import shutil
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference'
STAGED = os.path.join(REPO, 'baseline', '_ref')
FILES = ['train.py', 'loss_functions.py', 'inverse_warp.py', 'layers.py', 'utils.py', 'logger.py', 'custom_transforms.py']
DIRS = ['models', 'networks', 'datasets']
_MODS = ['train', 'models', 'loss_functions', 'inverse_warp', 'layers', 'utils', 'logger', 'custom_transforms', 'networks',
         'datasets']


def stage(dst=STAGED, src=SRC):
    """Copy the files of the reference that the training path imports into baseline/_ref/ (git-ignored, travels with gpurun)."""
    if not os.path.isdir(src):
        return False
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copy2(os.path.join(src, f), os.path.join(dst, f))
    for d in DIRS:
        shutil.copytree(os.path.join(src, d), os.path.join(dst, d), dirs_exist_ok=True,
                        ignore=shutil.ignore_patterns('__pycache__', '*.pyc', '*.pth', '*.npy', '*.png', '*.jpg'))
    return True


def find_root():
    """The reference checkout if present (authoring container), else the staged copy (GPU box), else None."""
    for r in (SRC, STAGED):
        if os.path.exists(os.path.join(r, 'train.py')):
            return r
    return None


class _Anything(types.ModuleType):
    """Stub module: any attribute that was not given explicitly resolves to a do-nothing callable / base class."""

    def __getattr__(self, item):
        if item.startswith('__'):
            raise AttributeError(item)
        return type(item, (), {'__init__': lambda self, *a, **k: None, '__call__': lambda self, *a, **k: None})


def _stub(name, **attrs):
    m = _Anything(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    import pathlib
    import scipy.misc
    if not hasattr(scipy.misc, 'imresize'):
        scipy.misc.imresize = lambda *a, **k: None             # imported at loss_functions.py:8, never called on the hot path
    if 'path' not in sys.modules:
        _stub('path', Path=pathlib.Path)
    if 'blessings' not in sys.modules:
        _stub('blessings', Terminal=type('Terminal', (), {'height': 24, 'width': 80}))
    if 'progressbar' not in sys.modules:
        _stub('progressbar', ProgressBar=object, Bar=object, ETA=object, SimpleProgress=object, Timer=object)
    if 'tensorboardX' not in sys.modules:
        _stub('tensorboardX', SummaryWriter=type('SummaryWriter', (), {'__init__': lambda self, *a, **k: None,
                                                                          'add_scalar': lambda self, *a, **k: None,
                                                                          'add_image': lambda self, *a, **k: None}))
    for name in ('imageio', 'h5py'):
        try:
            importlib.import_module(name)
        except Exception:  # noqa: BLE001
            _stub(name, imread=lambda *a, **k: None, imsave=lambda *a, **k: None)
    try:
        importlib.import_module('skimage.transform')
    except Exception:  # noqa: BLE001
        sk = sys.modules.get('skimage') or _stub('skimage')
        sk.transform = _stub('skimage.transform', resize=lambda *a, **k: None)


def import_reference(root=None, with_train=True):
    """Returns a namespace of the reference's modules imported from `root` (unmodified source files)."""
    root = root or find_root()
    if root is None:
        raise RuntimeError('reference not available: neither %s nor %s exists' % (SRC, STAGED))
    install_stubs()
    saved = {m: sys.modules.pop(m) for m in list(sys.modules) if m.split('.')[0] in _MODS}
    sys.path.insert(0, root)
    try:
        ds = types.ModuleType('datasets')                       # the reference's datasets/ has no __init__.py
        ds.__path__ = [os.path.join(root, 'datasets')]
        sys.modules['datasets'] = ds
        ns = types.SimpleNamespace(root=root)
        for m in ('models', 'loss_functions', 'inverse_warp', 'layers', 'utils', 'logger', 'custom_transforms'):
            setattr(ns, m, importlib.import_module(m))
        if with_train:
            ns.train = importlib.import_module('train')
    finally:
        sys.path.remove(root)
        for m in list(sys.modules):                              # leave no `models` / `datasets` ... behind for other importers
            if m.split('.')[0] in _MODS:
                sys.modules.pop(m)
        sys.modules.update(saved)
    return ns


class NullLogger(object):
    """What train.train needs of logger.TermLogger (logger.py:7-41): a progress bar and a line writer."""

    class _Bar(object):
        def update(self, *_):
            pass

    class _Writer(object):
        def __init__(self):
            self.lines = []

        def write(self, s):
            self.lines.append(s)

    def __init__(self):
        self.train_bar, self.train_writer = self._Bar(), self._Writer()
        self.valid_bar, self.valid_writer = self._Bar(), self._Writer()


class NullWriter(object):
    """tensorboardX.SummaryWriter surface that train.train touches (train.py:490-515)."""

    def __init__(self):
        self.scalars = []

    def add_scalar(self, tag, value, step):
        self.scalars.append((tag, value, step))

    def add_image(self, *a, **k):
        pass


def reference_args(save_path, **kw):
    """argparse defaults that train.train reads (train.py:28-91)."""
    import pathlib
    a = dict(photo_loss_weight=1.0, mask_loss_weight=0.0, smooth_loss_weight=0.0, unsupervised=False, dataset='kitti',
             loss='L1', monodepth2=False, diff_lr=False, rotation_mode='euler', padding_mode='zeros', print_freq=10,
             training_output_freq=0, batch_size=4, network='disp_vgg_BN', save_path=pathlib.Path(str(save_path)),
             log_full='progress_log_full.csv', ordinal_c=71)
    a.update(kw)
    return types.SimpleNamespace(**a)


if __name__ == '__main__':
    print('staged' if stage() else 'reference checkout not present', '->', STAGED)
