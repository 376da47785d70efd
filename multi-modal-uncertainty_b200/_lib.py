"""ctypes binding of ``libmmu_b200.so`` (the C ABI declared in ``include/mmu_b200.h``).

There is no fallback of any kind: if the shared library is missing, importing the package
raises; if a compute entry point is called without a CUDA device it returns an error code and
``check`` raises ``MMUError``.  Pointers are passed as ``tensor.data_ptr()`` and the stream as
``torch.cuda.current_stream().cuda_stream`` -- no torch types cross the boundary.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmu_b200.so")

F32, BF16 = 0, 1
EPI_STORE, EPI_QUICKGELU, EPI_RESIDUAL, EPI_DGELU, EPI_ATOMIC = 0, 1, 2, 3, 4
EPI_RESID_LN = 6

EXPORTS = [
    "mmu_version", "mmu_error_string", "mmu_launch_count", "mmu_gemm", "mmu_mask_gather_tokens", "mmu_layernorm_fwd",
    "mmu_layernorm_bwd", "mmu_batchaxis_attention_fwd", "mmu_batchaxis_attention_bwd",
    "mmu_heads_uncertainty_epilogue", "mmu_adamw_flat_step", "mmu_flava_param_count",
    "mmu_flava_param_table", "mmu_flava_workspace_bytes", "mmu_flava_num_stages",
    "mmu_flava_forward", "mmu_flava_backward", "mmu_cast_f32_to_bf16", "mmu_struct_size",
    "mmu_posthoc_scoring", "mmu_ragged_pad", "mmu_pair_concordance", "mmu_top_truncate",
    "mmu_resnet_param_count", "mmu_resnet_stat_count", "mmu_resnet_param_table",
    "mmu_resnet_stat_table", "mmu_resnet_workspace_bytes", "mmu_resnet_forward",
    "mmu_resnet_backward",
    "mmu_mmbt_param_count", "mmu_mmbt_param_table", "mmu_mmbt_workspace_bytes", "mmu_mmbt_forward",
    "mmu_mmbt_backward", "mmu_bertadam_flat_step",
    "mmu_imgenc_param_count", "mmu_imgenc_stat_count", "mmu_imgenc_param_table", "mmu_imgenc_stat_table",
    "mmu_imgenc_workspace_bytes", "mmu_imgenc_forward", "mmu_imgenc_backward",
    "mmu_seq_attention_fwd", "mmu_seq_attention_bwd", "mmu_modality_keep_mask", "mmu_set_gemm_sm_limit",
    "mmu_ln_fold_weights", "mmu_layernorm_raw_stats",
]


class MMUError(RuntimeError):
    pass


class GemmEpilogue(C.Structure):
    _fields_ = [("mode", C.c_int), ("out_lp", C.c_int), ("out", C.c_void_p), ("out2", C.c_void_p),
                ("bias", C.c_void_p), ("aux", C.c_void_p), ("ld_out", C.c_longlong),
                ("ld_out2", C.c_longlong), ("ld_aux", C.c_longlong), ("seg_len", C.c_int),
                ("seg_stride", C.c_int), ("seg_off", C.c_int), ("alpha", C.c_float),
                ("drop_p", C.c_float), ("drop_site", C.c_int), ("drop_seed", C.c_ulonglong),
                ("ln_stats", C.c_void_p), ("ln_cw", C.c_void_p), ("ln_nt", C.c_int),
                ("ln_inv_d", C.c_float), ("ln_eps", C.c_float), ("stats_nt", C.c_int),
                ("stats_out", C.c_void_p)]


class MetricAccum(C.Structure):
    _fields_ = [("conf_count", C.c_ulonglong * 15), ("conf_correct", C.c_ulonglong * 15),
                ("hpred_count", C.c_ulonglong * 32), ("mi_count", C.c_ulonglong * 32),
                ("n_samples", C.c_ulonglong), ("n_rows", C.c_ulonglong),
                ("n_correct_rows", C.c_ulonglong), ("n_correct_prob", C.c_ulonglong),
                ("conf_sum", C.c_double * 15), ("loss_sum", C.c_double),
                ("sum_h_pred", C.c_double), ("sum_h_exp", C.c_double), ("sum_mi", C.c_double)]


# Offsets (in 8-byte words) of the accumulator fields: the struct is 117 words of 8 bytes.
ACC_WORDS = C.sizeof(MetricAccum) // 8
ACC_OFF = {name: getattr(MetricAccum, name).offset // 8 for name, _ in MetricAccum._fields_}
ACC_INT_WORDS = ACC_OFF["conf_sum"]  # words [0, ACC_INT_WORDS) are uint64, the rest are doubles


class PosthocAccum(C.Structure):
    _fields_ = [("sx", C.c_double * 2), ("sy", C.c_double * 2), ("sxx", C.c_double * 2),
                ("syy", C.c_double * 2), ("sxy", C.c_double * 2), ("n_samples", C.c_ulonglong),
                ("correct", C.c_ulonglong * 128)]


class FlavaConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "l_img", "l_txt", "d_img", "d_txt", "D", "n_head",
                                       "n_layers", "E", "C", "avg_pool", "cls_token", "precision",
                                       "max_variants", "group_pool")]


class ParamEntry(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("offset", C.c_longlong), ("numel", C.c_longlong),
                ("rows", C.c_int), ("cols", C.c_int), ("stage", C.c_int)]


class ResNetConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "cin", "H", "W", "E", "C")]


class ImgEncConfig(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("layers", C.c_int * 4), ("width_per_group", C.c_int),
                ("pool_h", C.c_int), ("pool_w", C.c_int), ("pool_max", C.c_int)]


class MmbtConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "S_txt", "n_img", "d_img", "D", "n_head", "n_layers",
                                       "d_ff", "vocab", "max_pos", "n_types", "C", "cls_id",
                                       "sep_id", "precision", "max_seq")] + \
               [("drop_hidden", C.c_float), ("drop_attn", C.c_float), ("drop_img", C.c_float),
                ("drop_reserved", C.c_int)]


class MmbtInputs(C.Structure):
    _fields_ = [("txt", C.c_void_p), ("mask", C.c_void_p), ("segment", C.c_void_p),
                ("img", C.c_void_p), ("indices", C.c_void_p), ("n_sel", C.c_int),
                ("indices_per_sample", C.c_int), ("params_bf16", C.c_void_p), ("dimg", C.c_void_p),
                ("drop_seed", C.c_ulonglong)]


class FlavaInputs(C.Structure):
    _fields_ = [("img", C.c_void_p), ("txt", C.c_void_p), ("idx_img", C.c_void_p),
                ("idx_txt", C.c_void_p), ("n_img", C.c_int), ("n_txt", C.c_int),
                ("keep", C.c_void_p), ("params_bf16", C.c_void_p), ("src_l_img", C.c_int),
                ("src_l_txt", C.c_int), ("n_variants", C.c_int), ("var_segments", C.c_void_p),
                ("drop_p", C.c_float), ("src_bf16", C.c_int), ("drop_seed", C.c_ulonglong)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C multi-modal-uncertainty_b200/csrc` (needs nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for this package.")
    lib = C.CDLL(LIB_PATH)
    vp, ll, i, f = C.c_void_p, C.c_longlong, C.c_int, C.c_float
    lib.mmu_version.restype = C.c_char_p
    lib.mmu_error_string.restype = C.c_char_p
    lib.mmu_error_string.argtypes = [i]
    lib.mmu_launch_count.restype = ll
    lib.mmu_gemm.argtypes = [i, vp, ll, i, vp, ll, i, i, i, i, i, C.POINTER(GemmEpilogue), vp]
    lib.mmu_mask_gather_tokens.argtypes = [vp, vp, i, i, i, i, vp, i, vp, i, i, vp]
    lib.mmu_ln_fold_weights.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, vp]
    lib.mmu_layernorm_raw_stats.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, vp]
    lib.mmu_struct_size.argtypes = [i]
    lib.mmu_set_gemm_sm_limit.argtypes = [i]
    rcfgp = C.POINTER(ResNetConfig)
    for fn in (lib.mmu_resnet_param_count, lib.mmu_resnet_stat_count):
        fn.restype, fn.argtypes = ll, [rcfgp]
    for fn in (lib.mmu_resnet_param_table, lib.mmu_resnet_stat_table):
        fn.argtypes = [rcfgp, C.POINTER(ParamEntry), i]
    lib.mmu_resnet_workspace_bytes.restype = ll
    lib.mmu_resnet_workspace_bytes.argtypes = [rcfgp, i]
    lib.mmu_resnet_forward.argtypes = [rcfgp, vp, vp, vp, vp, vp, ll, i, vp, vp]
    lib.mmu_resnet_backward.argtypes = [rcfgp, vp, vp, vp, vp, vp, ll, vp, vp, vp]
    lib.mmu_ragged_pad.argtypes = [vp, vp, vp, i, i, i, vp]
    lib.mmu_modality_keep_mask.argtypes = [vp, vp, vp, vp, i, i, f, i, vp, vp]
    lib.mmu_posthoc_scoring.argtypes = [vp, vp, i, i, i, i, i, vp, vp, vp]
    lib.mmu_pair_concordance.argtypes = [vp, vp, ll, i, ll, ll, vp, vp]
    lib.mmu_top_truncate.argtypes = [vp, vp, i, i, i, i, vp, vp]
    lib.mmu_cast_f32_to_bf16.argtypes = [vp, vp, C.c_size_t, vp]
    lib.mmu_layernorm_fwd.argtypes = [vp, vp, vp, vp, i, vp, vp, i, i, vp]
    lib.mmu_layernorm_bwd.argtypes = [vp, i, vp, vp, vp, vp, vp, i, vp, i, vp, vp, vp, i, i, vp]
    lib.mmu_batchaxis_attention_fwd.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
    lib.mmu_batchaxis_attention_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
    lib.mmu_heads_uncertainty_epilogue.argtypes = [vp, vp, i, i, i, i, i, i, f, vp, vp, vp, vp, vp]
    lib.mmu_adamw_flat_step.argtypes = [vp, vp, vp, vp, vp, C.c_size_t, f, f, f, f, f, i, f, vp]
    cfgp = C.POINTER(FlavaConfig)
    lib.mmu_flava_param_count.restype = ll
    lib.mmu_flava_param_count.argtypes = [cfgp]
    lib.mmu_flava_param_table.argtypes = [cfgp, C.POINTER(ParamEntry), i]
    lib.mmu_flava_workspace_bytes.restype = ll
    lib.mmu_flava_workspace_bytes.argtypes = [cfgp, i]
    lib.mmu_flava_num_stages.argtypes = [cfgp]
    lib.mmu_flava_forward.argtypes = [cfgp, vp, C.POINTER(FlavaInputs), vp, ll, i, vp, vp]
    lib.mmu_flava_backward.argtypes = [cfgp, vp, C.POINTER(FlavaInputs), vp, ll, vp, vp, i, i, vp]
    lib.mmu_seq_attention_fwd.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
    lib.mmu_seq_attention_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp]
    icfgp = C.POINTER(ImgEncConfig)
    for fn in (lib.mmu_imgenc_param_count, lib.mmu_imgenc_stat_count):
        fn.restype, fn.argtypes = ll, [icfgp]
    for fn in (lib.mmu_imgenc_param_table, lib.mmu_imgenc_stat_table):
        fn.argtypes = [icfgp, C.POINTER(ParamEntry), i]
    lib.mmu_imgenc_workspace_bytes.restype, lib.mmu_imgenc_workspace_bytes.argtypes = ll, [icfgp, i]
    lib.mmu_imgenc_forward.argtypes = [icfgp, vp, vp, vp, vp, vp, ll, i, vp, vp]
    lib.mmu_imgenc_backward.argtypes = [icfgp, vp, vp, vp, vp, vp, ll, vp, vp, vp]
    mcfgp, minp = C.POINTER(MmbtConfig), C.POINTER(MmbtInputs)
    lib.mmu_mmbt_param_count.restype, lib.mmu_mmbt_param_count.argtypes = ll, [mcfgp]
    lib.mmu_mmbt_param_table.argtypes = [mcfgp, C.POINTER(ParamEntry), i]
    lib.mmu_mmbt_workspace_bytes.restype, lib.mmu_mmbt_workspace_bytes.argtypes = ll, [mcfgp, i]
    lib.mmu_mmbt_forward.argtypes = [mcfgp, vp, minp, vp, ll, i, vp, vp]
    lib.mmu_mmbt_backward.argtypes = [mcfgp, vp, minp, vp, ll, vp, vp, vp]
    lib.mmu_bertadam_flat_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i, ll, f, f, f, f, f, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:
            fn.restype = C.c_int
    return lib


lib = _load()


def check(rc, what="mmu call"):
    if rc < 0:
        raise MMUError(f"{what} failed: {lib.mmu_error_string(int(rc)).decode()} (code {rc})")
    return rc


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()
