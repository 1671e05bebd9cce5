"""CPU: the C-ABI shared library builds, loads and exports every symbol include/dispnet_b200.h declares; the ctypes
mirror of the structs matches the C layout; the product package has no dependency on the oracle and no CPU path."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'dispnet_b200.h')


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    from supervised_dispnet_b200 import _lib as L
    return L


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dn_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 40
    L = lib.lib()
    for n in names:
        assert hasattr(L, n), 'missing export ' + n
    # and the binding declares a signature for each of them
    for n in names:
        assert n in lib.EXPORTS, 'no ctypes signature for ' + n


def test_struct_layout_matches_c(lib, tmp_path):
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dispnet_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(dn_view),sizeof(dn_tap),sizeof(dn_igemm),sizeof(dn_wgrad),offsetof(dn_igemm,taps),'
                   'offsetof(dn_igemm,out_pad_ok),offsetof(dn_wgrad,scale),sizeof(dn_pack_job),offsetof(dn_pack_job,row_scale),sizeof(dn_photo_scale),sizeof(dn_photo_batch),'
                   'offsetof(dn_photo_batch,K),sizeof(dn_pyr_job));'
                   'return 0;}\n')
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    c = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    py = [ctypes.sizeof(lib.DnView), ctypes.sizeof(lib.DnTap), ctypes.sizeof(lib.DnIgemm), ctypes.sizeof(lib.DnWgrad),
          lib.DnIgemm.taps.offset, lib.DnIgemm.out_pad_ok.offset, lib.DnWgrad.scale.offset, ctypes.sizeof(lib.DnPackJob),
          lib.DnPackJob.row_scale.offset, ctypes.sizeof(lib.DnPhotoScale), ctypes.sizeof(lib.DnPhotoBatch), lib.DnPhotoBatch.K.offset,
          ctypes.sizeof(lib.DnPyrJob)]
    assert c == py


def test_error_strings_and_version(lib):
    L = lib.lib()
    assert L.dn_version() >= 100
    assert b'invalid argument' in L.dn_error_string(-1)
    assert L.dn_reduce_ws_floats(64) > 0


def test_sass_has_tcgen05_and_tma():
    """The shipped binary really contains 5th-gen tensor-core and TMA instructions for sm_100a."""
    so = os.path.join(ROOT, 'supervised_dispnet_b200', 'libdispnet_b200.so')
    sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass or 'SM100' in sass.upper()
    for mnem in ('UTCHMMA', 'UTMALDG', 'LDTM', 'UTCBAR', 'UTMASTG.4D', 'UTMALDG.5D'):      # (TMA-store epilogue, 5-D operand maps)
        assert mnem in sass, mnem


def test_product_does_not_import_oracle_or_reference():
    pkg = os.path.join(ROOT, 'supervised_dispnet_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', txt, flags=re.M), f
                assert '/root/reference' not in txt, f


def test_no_cpu_fallback():
    import supervised_dispnet_b200 as S
    with pytest.raises(RuntimeError):
        S.models.DispNetS()(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        S.loss_functions.l1_loss(torch.rand(1, 8, 8), [torch.rand(1, 1, 8, 8)], 'kitti')
    with pytest.raises(RuntimeError):
        S.inverse_warp.inverse_warp(torch.rand(1, 3, 8, 8), torch.rand(1, 8, 8), torch.zeros(1, 6), torch.eye(3)[None],
                                    torch.eye(3)[None])


@pytest.mark.parametrize('cls,kw,sd_name', [('Disp_vgg_BN', {}, 'Disp_vgg_BN'), ('DispNetS', {}, 'DispNetS'),
                                            ('Disp_res_50', {}, 'Disp_res_50'),
                                            ('PoseExpNet', dict(nb_ref_imgs=4, output_exp=True), 'PoseExpNet')])
def test_state_dict_keys_match_reference_layout(cls, kw, sd_name):
    """Checkpoint compatibility (train.py:281,378): same keys and shapes as the reference modules (oracle layout tables are
    pinned bit-wise to the reference by tests/test_oracle_golden.py::test_g0)."""
    import supervised_dispnet_b200 as S
    from oracle import nets as ON
    m = getattr(S.models, cls)(**kw)
    ref = ON.init_state_dict(sd_name, 0, **kw)
    mine = m.state_dict()
    assert set(mine) == set(ref)
    for k in ref:
        assert tuple(mine[k].shape) == tuple(ref[k].shape), k


def test_init_weights_matches_reference_rng_stream(golden):
    """torch.manual_seed(0); net.init_weights() reproduces the reference's initial weights bit for bit."""
    import supervised_dispnet_b200 as S
    fp = golden('g0_init_fingerprints')['DispNetS']
    m = S.models.DispNetS()
    torch.manual_seed(0)
    m.init_weights()
    for k, v in m.state_dict().items():
        assert (float(v.double().sum()), float(v.double().abs().sum())) == fp[k], k


def test_dead_classifier_is_frozen_for_ddp():
    import supervised_dispnet_b200 as S
    m = S.models.Disp_vgg_BN()
    dead = [n for n, p in m.named_parameters() if not p.requires_grad]
    assert dead and all(n.startswith('features.classifier') for n in dead)
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 19_873_156       # 143 516 012 minus the 123 642 856 classifier
