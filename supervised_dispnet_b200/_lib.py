"""ctypes binding of libdispnet_b200.so (the C ABI declared in include/dispnet_b200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or a call fails, this
module raises.  PyTorch is used by the callers only for device memory, streams and autograd glue.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DISPNET_B200_LIB') or os.path.join(_HERE, 'libdispnet_b200.so')      # (override: A/B runs of two builds)

DN_F32, DN_F16, DN_BF16, DN_BF16_LO = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
MAX_TAPS, MAX_SRC = 160, 8


class DnView(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('dtype', C.c_int32), ('N', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('C', C.c_int32), ('c_ext', C.c_int32), ('sN', C.c_int64), ('sH', C.c_int64), ('sW', C.c_int64)]


class DnTap(C.Structure):
    _fields_ = [('src', C.c_int32), ('dh', C.c_int32), ('dw', C.c_int32), ('wt', C.c_int32)]


class DnIgemm(C.Structure):
    _fields_ = [('inp', DnView * MAX_SRC), ('nsrc', C.c_int32), ('out', DnView), ('w', C.c_void_p),
                ('w_dtype', C.c_int32), ('cin_pad', C.c_int32), ('cout_pad', C.c_int32), ('bias', C.c_void_p),
                ('act', C.c_int32), ('accumulate', C.c_int32), ('stride', C.c_int32), ('ntaps', C.c_int32),
                ('taps', DnTap * MAX_TAPS), ('out_scale', C.c_float), ('out_pad_ok', C.c_int32), ('out2', C.c_void_p),
                ('out2_dtype', C.c_int32), ('nphase', C.c_int32), ('phase_off', C.c_int64 * 4), ('phase_cout', C.c_int32),
                ('pad2_', C.c_int32)]


class DnWgrad(C.Structure):
    _fields_ = [('p', DnView * MAX_SRC), ('nsrc', C.c_int32), ('q', DnView), ('dw', C.c_void_p),
                ('cp_pad', C.c_int32), ('cq_pad', C.c_int32), ('stride', C.c_int32), ('ntaps', C.c_int32),
                ('taps', DnTap * MAX_TAPS), ('scale', C.c_float)]


class DnPackJob(C.Structure):
    _fields_ = [('src', C.c_void_p), ('dst', C.c_void_p), ('dst_dtype', C.c_int32), ('unpack', C.c_int32), ('T', C.c_int32),
                ('R', C.c_int32), ('Cc', C.c_int32), ('R_pad', C.c_int32), ('C_pad', C.c_int32), ('k', C.c_int32),
                ('s_r', C.c_int64), ('s_c', C.c_int64), ('s_kh', C.c_int64), ('s_kw', C.c_int64), ('scale', C.c_float),
                ('pad_', C.c_int32), ('row_scale', C.c_void_p)]


PHOTO_MAX_SCALES, PHOTO_MAX_REFS = 4, 4


class DnPhotoScale(C.Structure):
    _fields_ = [('tgt', C.c_void_p), ('ref', C.c_void_p * PHOTO_MAX_REFS), ('depth', C.c_void_p), ('mask', C.c_void_p),
                ('gdepth', C.c_void_p), ('gmask', C.c_void_p), ('h', C.c_int32), ('w', C.c_int32), ('downscale', C.c_float),
                ('pad_', C.c_int32)]


class DnPhotoBatch(C.Structure):
    _fields_ = [('sc', DnPhotoScale * PHOTO_MAX_SCALES), ('nscales', C.c_int32), ('nrefs', C.c_int32), ('B', C.c_int32),
                ('rot_mode', C.c_int32), ('pad_mode', C.c_int32), ('align_corners', C.c_int32), ('K', C.c_void_p),
                ('Kinv', C.c_void_p), ('pose', C.c_void_p)]


class DnPyrJob(C.Structure):
    _fields_ = [('src', C.c_void_p), ('l1', C.c_void_p), ('l2', C.c_void_p), ('l3', C.c_void_p)]


_P = C.c_void_p
_I = C.c_int
_I64 = C.c_int64
_F = C.c_float
_D = C.c_double
_V = C.POINTER(DnView)
_IP = C.POINTER(C.c_int32)

_SIGS = {
    'dn_version': ([], _I),
    'dn_tc_available': ([], _I),
    'dn_tc_set_debug': ([_P], _I),
    'dn_tc_set_halo': ([_I], _I),
    'dn_pack_input': ([_P, _I, _I, _I, _I, _V, _I, _P], _I),
    'dn_input_transform': ([_P, _I, _I, _I, _I, _P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P], _I),
    'dn_flip_rows': ([_P, _I, _I64, _I, _P, _P, _P], _I),
    'dn_pack_weight': ([_P, _P, _I, _I, _I, _I, _I, _I, _IP, _IP, _I64, _I64, _I64, _I64, _P], _I),
    'dn_unpack_wgrad': ([_P, _P, _I, _I, _I, _I, _I, _IP, _IP, _I64, _I64, _I64, _I64, _F, _P], _I),
    'dn_pack_jobs': ([_P, _I, _P], _I),
    'dn_igemm_run': ([C.POINTER(DnIgemm), _I, _P], _I),
    'dn_wgrad_run': ([C.POINTER(DnWgrad), _I, _P], _I),
    'dn_igemm_tc_supported': ([C.POINTER(DnIgemm)], _I),
    'dn_wgrad_tc_supported': ([C.POINTER(DnWgrad)], _I),
    'dn_reduce_ws_floats': ([_I], _I64),
    'dn_rowx_expand': ([_V, _I, _I, _I, _V, _V, _P], _I),
    'dn_rowx_pack_weight': ([_P, _I, _I, _I, _P, _I, _I, _I, _P, _P], _I),
    'dn_rowx_unpack_wgrad': ([_P, _P, _I, _I, _I, _I, _I, _F, _P], _I),
    'dn_bn_stats': ([_V, _P, _P, _P], _I),
    'dn_bn_fold_bias': ([_P, _P, _I, _P, _P], _I),
    'dn_resize_bilinear_ac': ([_P, _I, _I, _I, _I, _I, _P, _P], _I),
    'dn_bn_finalize': ([_P, _D, _P, _P, _P, _P, _F, _F, _I, _I, _P, _P, _I, _P], _I),
    'dn_bn_train_stats': ([_V, _P, _P, _P, _P, _P, _F, _F, _I, _P, _P, _P, _P, _P], _I),
    'dn_bn_apply': ([_V, _P, _V, _I, _I, _V, _V, _P], _I),
    'dn_bn_bwd_reduce': ([_V, _V, _V, _P, _P, _P, _I, _I, _P, _P, _P], _I),
    'dn_bn_bwd_apply': ([_V, _V, _V, _P, _P, _P, _I, _I, _P, _D, _F, _P, _P, _V, _V, _I, _P], _I),
    'dn_act_bwd': ([_V, _V, _I, _P, _F, _P, _P], _I),
    'dn_maxpool_fwd': ([_V, _V, _I, _I, _I, _P], _I),
    'dn_maxpool_bwd': ([_V, _V, _V, _I, _I, _I, _I, _P], _I),
    'dn_add_act_fwd': ([_V, _V, _I, _V, _P], _I),
    'dn_add_act_bwd': ([_V, _V, _I, _V, _I, _V, _I, _P], _I),
    'dn_act_fwd': ([_V, _I, _V, _P], _I),
    'dn_copy_view': ([_V, _V, _I, _P], _I),
    'dn_split_bf16': ([_V, _V, _V, _P], _I),
    'dn_head_conv_fwd': ([_V, _P, _P, _V, _P], _I),
    'dn_head_conv_bwd': ([_V, _P, _V, _V, _I, _P, _P, _F, _P, _P], _I),
    'dn_head_fwd': ([_V, _F, _F, _P, _V, _I, _P], _I),
    'dn_head_fwd2': ([_V, _F, _F, _P, _V, _V, _I, _P], _I),
    'dn_head_bwd': ([_P, _V, _I, _V, _F, _F, _V, _P], _I),
    'dn_sigmoid_nchw_fwd': ([_V, _P, _P], _I),
    'dn_sigmoid_nchw_bwd': ([_P, _P, _F, _V, _P], _I),
    'dn_spatial_mean_fwd': ([_V, _F, _P, _P], _I),
    'dn_spatial_mean_bwd': ([_P, _F, _V, _P], _I),
    'dn_l1_fwd': ([_P, _P, _I, _I, _F, _P, _P, _P], _I),
    'dn_l1_bwd': ([_P, _P, _I, _I, _F, _P, _P, _P, _P], _I),
    'dn_depth_loss_ws_floats': ([_I], _I64),
    'dn_depth_loss_fwd': ([_P, _P, _I, _I, _I, _I, _I, _F, _I, _I, _F, _I, _P, _P, _P], _I),
    'dn_depth_loss_bwd': ([_P, _P, _I, _I, _I, _I, _I, _F, _I, _I, _F, _P, _P, _P, _P], _I),
    'dn_pool2': ([_P, _I64, _I, _I, _I, _P, _P], _I),
    'dn_smooth_fwd': ([_P, _I, _I, _I, _F, _P, _P], _I),
    'dn_smooth_bwd': ([_P, _I, _I, _I, _F, _P, _P, _P], _I),
    'dn_depth_errors': ([_P, _P, _I, _I, _I, _F, _I, _I, _I, _I, _I, _P, _P, _P, _P], _I),
    'dn_area_down': ([_P, _I, _I, _I, _I, _P, _P], _I),
    'dn_area_pyramid': ([C.POINTER(DnPyrJob), _I, _I64, _I, _I, _P], _I),
    'dn_photo_ws_floats': ([C.POINTER(DnPhotoBatch)], _I64),
    'dn_photo_batch_fwd': ([C.POINTER(DnPhotoBatch), _P, _P, _P, _P], _I),
    'dn_photo_batch_bwd': ([C.POINTER(DnPhotoBatch), _P, _P, _P, _P], _I),
    'dn_warp_photo_fwd': ([_P, _P, _P, _P, _I, _P, _P, _P, _I64, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P], _I),
    'dn_warp_photo_bwd': ([_P, _P, _P, _P, _I, _P, _P, _P, _I64, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I64, _P,
                           _P], _I),
    'dn_inverse_warp_fwd': ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P], _I),
    'dn_inverse_warp_bwd': ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P], _I),
    'dn_explain_fwd': ([_P, _I64, _P, _P], _I),
    'dn_explain_bwd': ([_P, _I64, _P, _P, _P], _I),
    'dn_ssim_fwd': ([_P, _P, _I, _I, _I, _P, _P], _I),
    'dn_ssim_bwd': ([_P, _P, _P, _I, _I, _I, _P, _P, _P], _I),
    'dn_edge_smooth_fwd': ([_P, _P, _I, _I, _I, _I, _P, _P], _I),
    'dn_edge_smooth_bwd': ([_P, _P, _I, _I, _I, _I, _P, _P, _P], _I),
    'dn_depth_errors_raw': ([_P, _P, _I64, _P, _P, _P], _I),
    'dn_fill_f32': ([_P, _I64, _F, _P], _I),
    'dn_axpy_f32': ([_P, _F, _P, _I64, _P], _I),
}

EXPORTS = tuple(sorted(list(_SIGS) + ['dn_error_string']))

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'libdispnet_b200.so is missing (%s): build it with `python -c "import __graft_entry__ as g; '
                'g.build()"` or `make -C supervised_dispnet_b200/csrc`. There is no CPU / PyTorch fallback.' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = res
        L.dn_error_string.argtypes = [_I]
        L.dn_error_string.restype = C.c_char_p
        _lib = L
    return _lib


def check(code, what=''):
    if code != 0:
        msg = lib().dn_error_string(code).decode()
        raise RuntimeError('dispnet_b200 %s failed: %s (code %d)' % (what, msg, code))


# ---- optional per-call device timing (bench.py roofline pass): PROFILE = [] switches it on
PROFILE = None
CALLS = 0


def call(name, *args, tag=None):
    global CALLS
    CALLS += 1
    if PROFILE is None:
        check(getattr(lib(), name)(*args), name)
        return
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib(), name)(*args), name)
    e1.record()
    PROFILE.append((name, tag, e0, e1))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    """Every tensor handed to a kernel must live on the CURRENT CUDA device: launches go to that device's current stream
    (stream_ptr) and the library never switches devices, so a tensor of another GPU would be an illegal access / a
    wrong-stream race.  One process per GPU with torch.cuda.set_device(LOCAL_RANK) (bench.py, DDP) satisfies this; use
    `with torch.cuda.device(t.device):` otherwise."""
    import torch
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('dispnet_b200: tensors must live on a CUDA device (got %s); the CUDA extension is the '
                               'only implementation of this path' % t.device)
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError('dispnet_b200: tensor on cuda:%d but the current device is cuda:%d -- wrap the call in '
                               '`with torch.cuda.device(tensor.device):`' % (t.device.index, cur))
