"""ctypes binding of ``libs4former_b200.so`` (the C ABI declared in ``include/s4former.h``).

The product path has no CPU or PyTorch fallback: if the library is missing or a call returns
a non-zero code an exception is raised (``S4Error``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libs4former_b200.so')

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU = 0, 1
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TC = 0, 1, 2


class S4Error(RuntimeError):
    pass


class GemmParams(C.Structure):
    _fields_ = [
        ('a', C.c_void_p), ('b', C.c_void_p), ('c', C.c_void_p), ('bias', C.c_void_p),
        ('aux', C.c_void_p), ('res', C.c_void_p), ('pre', C.c_void_p),
        ('M', C.c_int), ('N', C.c_int), ('K', C.c_int), ('nb1', C.c_int), ('nb2', C.c_int),
        ('a_sm', C.c_longlong), ('a_sk', C.c_longlong), ('a_b1', C.c_longlong), ('a_b2', C.c_longlong),
        ('b_sk', C.c_longlong), ('b_sn', C.c_longlong), ('b_b1', C.c_longlong), ('b_b2', C.c_longlong),
        ('c_sm', C.c_longlong), ('c_b1', C.c_longlong), ('c_b2', C.c_longlong),
        ('alpha', C.c_float), ('act', C.c_int), ('accumulate', C.c_int), ('dtype', C.c_int),
        ('c_dtype', C.c_int), ('backend', C.c_int), ('split_k', C.c_int),
        ('colsum', C.c_void_p),
    ]


_P, _I, _L, _F, _D, _Z = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); the stream is always the last argument of compute calls
_SPEC = {
    's4_version': (_I, []),
    's4_built_arch': (_I, []),
    's4_last_error': (C.c_char_p, []),
    's4_launch_count': (C.c_longlong, []),
    's4_prof_enable': (_I, [_I]),
    's4_prof_num_kinds': (_I, []),
    's4_prof_get': (_I, [_I, C.c_char_p, _I, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                         C.POINTER(C.c_longlong)]),
    's4_gemm': (_I, [C.POINTER(GemmParams), _P]),
    's4_gemm_uses_tc': (_I, [C.POINTER(GemmParams)]),
    's4_set_tc_pair_mode': (_I, [_I]),
    's4_set_tc_sched': (_I, [_I]),
    's4_layernorm_fwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _I, _P]),
    's4_layernorm_bwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    's4_attention_workspace': (_Z, [_I, _I, _I, _I, _I, _I]),
    's4_attention_fwd': (_I, [_P, _P, _P, _F, _P, _P, _P, _Z, _I, _I, _I, _I, _I, _I, _P]),
    's4_attention_bwd': (_I, [_P, _P, _P, _P, _P, _P, _F, _P, _P, _Z, _I, _I, _I, _I, _I, _I, _P]),
    's4_attention_set_trace': (_I, [_P, _I]),
    's4_patchify': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    's4_assemble_tokens': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    's4_assemble_tokens_bwd': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    's4_colsum': (_I, [_P, _P, _P, _L, _I, _I, _P]),
    's4_cast': (_I, [_P, _P, _L, _I, _I, _P]),
    's4_transpose': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    's4_conv3x3_fwd': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    's4_conv3x3_fwd_stats': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    's4_conv3x3_dgrad': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    's4_conv3x3_wgrad': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    's4_pack_conv3x3_weight': (_I, [_P, _P, _P, _I, _I, _I, _P]),
    's4_bn_finalize': (_I, [_P, _P, _D, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    's4_bn_eval_affine': (_I, [_P, _P, _P, _P, _F, _P, _P, _I, _P]),
    's4_bn_relu_upsample_fwd': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    's4_bn_relu_upsample_bwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    's4_bn_bwd_apply': (_I, [_P, _P, _P, _P, _P, _P, _P, _D, _P, _L, _I, _I, _P]),
    's4_bn_relu_conv1x1_fwd': (_I, [_P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P]),
    's4_bn_relu_conv1x1_bwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P]),
    's4_upsample_logits_fwd': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    's4_upsample_logits_bwd': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    's4_cls_supported': (_I, [_I, _I, _I]),
    's4_cls_upsample_bwd_padded': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    's4_cls_bwd_reduce': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P]),
    's4_cls_bwd_apply': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _D, _P, _L, _I, _I, _P]),
    's4_pseudo_label': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
    's4_ce_ncr_workspace': (_Z, [_I, _I, _I]),
    's4_ce_ncr': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _I, _P, _Z, _P]),
    's4_scale_by_scalar': (_I, [_P, _P, _Z, _P]),
    's4_ce_ncr_grad_fixup': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _I, _P]),
    's4_cutmix': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    's4_patchshuffle': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    's4_chunk_elems': (_I, []),
    's4_ema_multi_tensor': (_I, [_P, _P, _P, _P, _P, _P, _I, _F, _F, _P]),
    's4_sgd_multi_tensor': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _F, _F, _I, _P]),
    's4_resize_bilinear_nchw': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    's4_softmax_argmax_nchw': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    's4_accumulate_crop': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    's4_intersect_union': (_I, [_P, _P, _L, _I, _L, _P, _P]),
    's4_branch_pipeline': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _I, _I, _I, _P]),
    's4_pmd_params_size': (_I, []),
    's4_peer_allreduce_buffer_bytes': (_L, []),
    's4_peer_allreduce_max_elems': (_I, []),
    's4_peer_allreduce_f32': (_I, [_P, _I, _P, _I, _I, _P, _P]),
    's4_bn_finalize_peer': (_I, [_P, _D, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _I, _I, _P, _P]),
    's4_sgd_ema_multi_tensor': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _F, _F, _I, _P]),
}

EXPORTED_SYMBOLS = tuple(_SPEC)
_lib = None


def load():
    """Load the shared library (raises S4Error when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise S4Error(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'or `make -C s4former_b200/csrc`. There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SPEC.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().s4_last_error().decode('utf-8', 'replace')
        raise S4Error(f'{what} failed with code {rc}: {msg}')


def call(name, *args):
    """Call a compute entry point and raise on a non-zero return code."""
    check(getattr(load(), name)(*args), name)
