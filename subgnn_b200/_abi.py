"""ctypes binding of the C ABI declared in include/subgnn_b200.h.

There is NO fallback: importing this module raises if libsubgnn_b200.so has not been built
(``python -c 'import __graft_entry__ as g; g.build()'`` or ``./build.sh``), and every call
raises ``SubgnnError`` on a non-zero status.
"""
import ctypes as C
import os
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / 'libsubgnn_b200.so'
if os.environ.get('SUBGNN_B200_LIB'):           # tuning aid: A/B another BUILD of the same library (tools/ab_bench.sh)
    _LIB_PATH = Path(os.environ['SUBGNN_B200_LIB']).resolve()


class SubgnnError(RuntimeError):
    pass


if not _LIB_PATH.exists():
    raise ImportError('libsubgnn_b200.so is missing at %s — build it with ./build.sh (no CPU fallback exists)' % _LIB_PATH)

lib = C.CDLL(str(_LIB_PATH))

P, I, F, U64, U32, LL = C.c_void_p, C.c_int, C.c_float, C.c_ulonglong, C.c_uint, C.c_longlong

# name -> argtypes (every function returns int status unless listed in _OTHER)
SIGNATURES = {
    'subgnn_walk_full': [P, P, I, I, I, F, U64, P, P],
    'subgnn_walk_patch': [P, P, I, P, P, I, I, I, I, F, I, U64, P, P],
    'subgnn_sample_rows': [P, P, I, I, I, I, U64, U32, I, P, P],
    'subgnn_border_khop_bitmap': [P, P, I, P, P, I, I, P, P, P],
    'subgnn_border_khop_expand': [P, I, I, P, P, P],
    'subgnn_sp_min_dense': [P, I, LL, P, P, I, P, P],
    'subgnn_sp_min_gather': [P, LL, P, P, I, P, P, I, P, P],
    'subgnn_degree_seq': [P, P, P, I, I, I, P, P, P],
    'subgnn_dtw_batch': [P, P, I, I, P, P, I, I, I, I, I, P, P],
    'subgnn_dtw_batch_rows': [P, P, P, I, I, P, P, I, I, I, I, I, P, P],
    'subgnn_hop_table': [P, P, I, I, I, P, LL, P],
    'subgnn_linear_fwd': [P, I, P, P, I, P, P, I, I, I, I, I, P],
    'subgnn_tc_linear_fwd': [P, I, P, P, I, P, P, I, I, I, I, I, P],
    'subgnn_tc_linear_bwd_input': [P, I, P, I, P, I, P, I, I, I, I, P],
    'subgnn_tc_linear_bwd_weight': [P, I, P, I, P, P, I, P, I, I, I, P],
    'subgnn_linear_bwd_input': [P, I, P, I, P, I, P, I, I, I, I, P],
    'subgnn_linear_bwd_weight': [P, I, P, I, P, P, I, P, I, I, I, P, P],
    'subgnn_tc_gemm_group': [P, I, I, P],
    'subgnn_gather_rows': [P, P, P, I, I, P],
    'subgnn_colsum': [P, I, P, I, I, P, P],
    'subgnn_lstm_prep': [P, P, P, P, P, I, P],
    'subgnn_lstm_prep_layers': [P, P, P, P, P, I, I, P],
    'subgnn_lstm_recur_fwd': [P, P, P, P, I, I, I, I, I, P],
    'subgnn_lstm_recur_bwd': [P, P, P, P, P, I, I, I, I, I, I, P, P, P],
    'subgnn_lstm_recur_fwd_drop': [P, P, P, P, I, I, I, I, I, P, F, U64, U32, P, P],
    'subgnn_lstm_recur_fwd_tc': [P, P, P, P, I, I, I, I, I, P, F, U64, U32, P, P],
    'subgnn_lstm_recur_fwd_tc_supported': [I],
    'subgnn_lstm_recur_bwd_drop': [P, P, P, P, P, I, I, I, I, I, I, P, P, F, U64, U32, P, P],
    'subgnn_lstm_recur_bwd_add': [P, P, P, P, P, P, I, I, I, I, I, I, P, P, F, U64, U32, P, P],
    'subgnn_lstm_head_fwd': [P, P, P, P, P, I, I, I, I, I, I, P],
    'subgnn_lstm_head_bwd': [P, P, P, P, I, I, I, I, I, I, P],
    'subgnn_dropout': [P, P, LL, F, U64, U32, P, P],
    'subgnn_model_prep_batch': [P, P],
    'subgnn_model_prep_weights': [P, P],
    'subgnn_model_q_fwd_part': [P, I, P],
    'subgnn_model_rows_fwd': [P, I, P],
    'subgnn_model_mlp_fwd': [P, P],
    'subgnn_model_mlp_stage': [P, I, I, P],
    'subgnn_model_readout': [P, P],
    'subgnn_model_rows_bwd': [P, I, P],
    'subgnn_model_mlp_bwd': [P, P],
    'subgnn_model_q_bwd': [P, P],
    'subgnn_model_q_bwd_part': [P, I, P],
    'subgnn_model_wgrad': [P, P],
    'subgnn_model_mlp_wgrad': [P, P],
    'subgnn_mpn_fwd': [P, P, P, I, P, P, P, P, P, P, P, P, I, I, I, P],
    'subgnn_mpn_bwd': [P, P, P, P, P, P, P, P, I, I, I, P],
    'subgnn_fill_zero': [P, LL, P],
    'subgnn_grad_sumsq': [P, LL, P, P],
    'subgnn_adam_step': [P, P, P, P, LL, F, F, F, F, P, P, F, F, P],
    'subgnn_clip_adam_step': [P, P, P, P, LL, F, F, F, F, P, P, F, F, P],
    'subgnn_sum_to_scalar': [P, I, P, P],
    'subgnn_dp_reduce_scatter': [P, P, P, P, P, I, I, LL, LL, P, P],
    'subgnn_dp_adam_allgather': [P, P, P, P, I, I, LL, LL, P, P, P, F, F, F, F, P, P, F, F, P],
    'subgnn_inc_step': [P, P],
}
_OTHER = {
    'subgnn_last_error': ([], C.c_char_p),
    'subgnn_abi_version': ([], I),
    'subgnn_device_sm_count': ([], I),
    'subgnn_model_desc_size': ([], I),
    'subgnn_gemm_desc_size': ([], I),
    'subgnn_model_readout_supported': ([P], I),
    'subgnn_tc_ws_available': ([], I),
    'subgnn_dp_flag_words': ([], I),
    'subgnn_launch_count': ([], U64),
    'subgnn_lstm_fused_dropout_supported': ([I], I),
    'subgnn_variant_log': ([C.c_char_p, I], I),
    'subgnn_variant_log_reset': ([], None),
}


def _bind():
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch: fail loudly
        fn.argtypes = args
        fn.restype = I
    for name, (args, res) in _OTHER.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res


def register(name, args):
    """Late registration used by modules that add entry points (keeps one table for the export test)."""
    SIGNATURES[name] = args
    fn = getattr(lib, name)
    fn.argtypes = args
    fn.restype = I


_bind()


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), 'C ABI takes contiguous device tensors'
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


_profile_hook = None   # set by bench.py's instrumented pass: hook(name) -> context manager timing one entry point


def call(name, *args):
    if _profile_hook is not None:
        with _profile_hook(name):
            rc = getattr(lib, name)(*args)
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise SubgnnError('%s failed (%d): %s' % (name, rc, lib.subgnn_last_error().decode()))


def _parse_desc_fields(struct='subgnn_model_desc'):
    """Builds the ctypes mirror of a descriptor struct from include/subgnn_b200.h so the two cannot drift."""
    import re
    hdr = (Path(__file__).resolve().parent.parent / 'include' / 'subgnn_b200.h').read_text()
    body = hdr.split('typedef struct %s {' % struct)[1].split('} %s;' % struct)[0]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for stmt in body.split(';'):
        stmt = ' '.join(stmt.split())
        if not stmt:
            continue
        m = re.match(r'^(const )?(unsigned long long|unsigned|int|float)( ?\*)? ?(.*)$', stmt)
        assert m, stmt
        base, is_ptr, names = m.group(2), bool(m.group(3)), m.group(4)
        ctype = P if is_ptr else {'unsigned long long': U64, 'unsigned': U32, 'int': I, 'float': F}[base]
        for name in names.split(','):
            name = name.strip()
            arr = re.match(r'^(\w+)\[(\d+)\]$', name)
            if arr:
                fields.append((arr.group(1), ctype * int(arr.group(2))))
            else:
                fields.append((name, ctype))
    return fields


class ModelDesc(C.Structure):
    _fields_ = _parse_desc_fields()


class GemmDesc(C.Structure):
    _fields_ = _parse_desc_fields('subgnn_gemm_desc')


GEMM_FWD, GEMM_BWD_INPUT, GEMM_BWD_WEIGHT, GEMM_BWD_WEIGHT_SHIFT = 0, 1, 2, 3      # include/subgnn_b200.h
assert C.sizeof(GemmDesc) == lib.subgnn_gemm_desc_size(), 'subgnn_gemm_desc layout mismatch'


def gemm_desc(op, a, lda, b, ldb, out, ldo, M, N, K, bias=None, scatter_ids=None, relu=0, accumulate=0, shift=0, period=0):
    """one problem of subgnn_tc_gemm_group; a / b / out / bias / scatter_ids are raw device addresses (int) or None."""
    d = GemmDesc()
    d.a, d.b, d.out, d.bias, d.scatter_ids = a, b, out, bias, scatter_ids
    d.op, d.lda, d.ldb, d.ldo, d.M, d.N, d.K = op, lda, ldb, ldo, M, N, K
    d.relu, d.accumulate, d.shift, d.period = relu, accumulate, shift, period
    return d


def gemm_group(descs, stream, max_ctas=0):
    arr = (GemmDesc * len(descs))(*descs)
    call('subgnn_tc_gemm_group', C.addressof(arr), len(descs), max_ctas, stream)


assert C.sizeof(ModelDesc) == lib.subgnn_model_desc_size(), \
    'subgnn_model_desc layout mismatch: ctypes %d vs C %d' % (C.sizeof(ModelDesc), lib.subgnn_model_desc_size())


def variant_log(reset=False):
    """set of kernel template instantiations launched so far (test aid, see include/subgnn_b200.h)."""
    buf = C.create_string_buffer(4096)
    lib.subgnn_variant_log(buf, 4096)
    out = set(x for x in buf.value.decode().split(';') if x)
    if reset:
        lib.subgnn_variant_log_reset()
    return out


def exported_symbols():
    return list(SIGNATURES) + list(_OTHER)
