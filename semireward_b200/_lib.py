"""ctypes binding of libsrw_b200.so (include/srw.h).  The library is the product: if it cannot be loaded this module
raises — there is no CPU or PyTorch fallback anywhere in semireward_b200."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRW_B200_LIB") or os.path.join(_HERE, "lib", "libsrw_b200.so")   # the override is an A/B aid for kernel work

c_f32p = C.c_void_p
i64 = C.c_int64
i32 = C.c_int
f32 = C.c_float
vp = C.c_void_p

# enums (include/srw.h)
EPI_F32, EPI_PLANES, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_SPLITK = range(6)
GEMM_TCGEN05, GEMM_SIMT, GEMM_TCGEN05_1CTA = 0, 1, 2


class Dropout(C.Structure):
    _fields_ = [("seq_key", vp), ("seq_row", vp), ("site", C.c_uint32), ("p", C.c_double)]


class SplitArgs(C.Structure):
    _fields_ = [("x", vp), ("ldx", i64), ("rows", i32), ("cols", i32), ("row_scale", vp), ("rows_per_scale", i32),
                ("planes", vp), ("ldp", i64), ("plane_stride", i64), ("planes_t", vp), ("ldpt", i64), ("plane_stride_t", i64),
                ("colsum_out", vp), ("colsum_accumulate", i32), ("colsum_workspace", vp)]


class GemmArgs(C.Structure):
    _fields_ = [("M", i32), ("N", i32), ("K", i32),
                ("a", vp), ("lda", i64), ("a_plane_stride", i64), ("a_mn_major", i32),
                ("b", vp), ("ldb", i64), ("b_plane_stride", i64), ("b_mn_major", i32),
                ("epilogue", i32), ("bias", vp), ("resid", vp), ("ldr", i64), ("row_scale", vp), ("rows_per_scale", i32),
                ("aux", vp), ("ldaux", i64), ("out_f32", vp), ("ldo", i64), ("out_planes", vp), ("ldp", i64),
                ("out_plane_stride", i64), ("split_k", i32), ("workspace", vp), ("impl", i32), ("max_ctas", i32),
                ("drop", Dropout), ("drop_rows_per_seq", i32), ("a_seg_k", i32), ("a_seg_rows", i64)]


class SplitKReduceArgs(C.Structure):
    _fields_ = [("workspace", vp), ("split_k", i32), ("M", i32), ("N", i32), ("out", vp), ("ldo", i64), ("accumulate", i32)]


class ColsumArgs(C.Structure):
    _fields_ = [("x", vp), ("ldx", i64), ("planes", vp), ("ldp", i64), ("plane_stride", i64), ("row_scale", vp),
                ("rows_per_scale", i32), ("rows", i32), ("cols", i32), ("out", vp), ("accumulate", i32), ("workspace", vp)]


class FoldColsum(C.Structure):
    _fields_ = [("partial", vp), ("nparts", i32), ("stride_p", i64), ("cols", i32), ("out", vp), ("accumulate", i32)]


class GradFoldArgs(C.Structure):
    _fields_ = [("n_splitk", i32), ("splitk", SplitKReduceArgs * 4), ("n_colsum", i32), ("colsum", FoldColsum * 8)]


class LayerNormFwdArgs(C.Structure):
    _fields_ = [("x", vp), ("ldx", i64), ("rows", i32), ("cols", i32), ("eps", f32), ("gamma", vp), ("beta", vp),
                ("mean", vp), ("rstd", vp), ("y_planes", vp), ("ldp", i64), ("plane_stride", i64), ("y_f32", vp), ("ldy", i64)]


class LayerNormBwdArgs(C.Structure):
    _fields_ = [("dy", vp), ("lddy", i64), ("x", vp), ("ldx", i64), ("rows", i32), ("cols", i32), ("gamma", vp), ("mean", vp),
                ("rstd", vp), ("dx", vp), ("lddx", i64), ("accumulate_dx", i32), ("dgamma", vp), ("dbeta", vp),
                ("accumulate_dparams", i32), ("workspace", vp), ("dx_planes", vp), ("ldp", i64), ("plane_stride", i64),
                ("row_scale", vp), ("rows_per_scale", i32), ("colsum_out", vp), ("colsum_accumulate", i32),
                ("drop", Dropout), ("drop_rows_per_seq", i32)]


class AttnFwdArgs(C.Structure):
    _fields_ = [("B", i32), ("N", i32), ("H", i32), ("head_dim", i32), ("scale", f32), ("qkv", vp), ("ld_qkv", i64),
                ("qkv_plane_stride", i64), ("o", vp), ("ld_o", i64), ("o_plane_stride", i64), ("lse", vp),
                ("key_bias", vp), ("ld_bias", i64), ("kv_len", vp), ("drop", Dropout)]


class AttnBwdArgs(C.Structure):
    _fields_ = [("B", i32), ("N", i32), ("H", i32), ("head_dim", i32), ("scale", f32), ("qkv", vp), ("ld_qkv", i64),
                ("qkv_plane_stride", i64), ("o", vp), ("ld_o", i64), ("o_plane_stride", i64), ("d_o", vp), ("ld_do", i64),
                ("do_plane_stride", i64), ("lse", vp), ("delta", vp), ("dqkv", vp), ("ld_dqkv", i64), ("dqkv_plane_stride", i64),
                ("key_bias", vp), ("ld_bias", i64), ("kv_len", vp), ("drop", Dropout)]


class VitConfig(C.Structure):
    _fields_ = [("img_size", i32), ("patch_size", i32), ("in_chans", i32), ("embed_dim", i32), ("depth", i32),
                ("num_heads", i32), ("hidden_dim", i32), ("num_classes", i32), ("ln_eps", f32)]


class VitFwdArgs(C.Structure):
    _fields_ = [("cfg", C.POINTER(VitConfig)), ("params", C.POINTER(vp)), ("weight_planes", vp), ("x", vp), ("batch", i32),
                ("grad_batch", i32), ("drop_scale", vp), ("logits", vp), ("feat", vp), ("workspace", vp),
                ("workspace_bytes", i64), ("gemm_impl", i32), ("tokens_out", vp)]


class VitBwdArgs(C.Structure):
    _fields_ = [("cfg", C.POINTER(VitConfig)), ("params", C.POINTER(vp)), ("weight_planes", vp), ("x", vp), ("batch", i32),
                ("grad_batch", i32), ("drop_scale", vp), ("dlogits", vp), ("dfeat", vp), ("grads", C.POINTER(vp)),
                ("accumulate_grads", i32), ("workspace", vp), ("workspace_bytes", i64), ("gemm_impl", i32), ("block_lo", i32), ("block_hi", i32)]


class BertConfig(C.Structure):
    _fields_ = [("vocab_size", i32), ("max_position", i32), ("type_vocab", i32), ("hidden", i32), ("layers", i32), ("heads", i32),
                ("intermediate", i32), ("num_classes", i32), ("ln_eps", f32), ("p_hidden", C.c_double), ("p_attn", C.c_double),
                ("p_pooled", C.c_double)]


class BertFwdArgs(C.Structure):
    _fields_ = [("cfg", C.POINTER(BertConfig)), ("params", C.POINTER(vp)), ("weight_planes", vp), ("input_ids", vp), ("attention_mask", vp),
                ("batch", i32), ("seq_len", i32), ("grad_batch", i32), ("drop_seq_key", vp), ("drop_seq_row", vp), ("pool_len", vp),
                ("logits", vp), ("feat", vp), ("workspace", vp), ("workspace_bytes", i64), ("gemm_impl", i32)]


class BertBwdArgs(C.Structure):
    _fields_ = [("cfg", C.POINTER(BertConfig)), ("params", C.POINTER(vp)), ("weight_planes", vp), ("input_ids", vp), ("attention_mask", vp),
                ("batch", i32), ("seq_len", i32), ("grad_batch", i32), ("drop_seq_key", vp), ("drop_seq_row", vp), ("pool_len", vp),
                ("dlogits", vp), ("dfeat", vp), ("grads", C.POINTER(vp)), ("accumulate_grads", i32), ("workspace", vp), ("workspace_bytes", i64),
                ("gemm_impl", i32), ("layer_lo", i32), ("layer_hi", i32)]


HUBERT_MAX_CONV = 8


class HubertConfig(C.Structure):
    _fields_ = [("hidden", i32), ("layers", i32), ("heads", i32), ("intermediate", i32), ("num_classes", i32), ("conv_dim", i32),
                ("num_conv", i32), ("conv_kernel", i32 * HUBERT_MAX_CONV), ("conv_stride", i32 * HUBERT_MAX_CONV), ("pos_kernel", i32),
                ("pos_groups", i32), ("ln_eps", f32), ("p_feat_proj", C.c_double), ("p_hidden", C.c_double), ("p_attn", C.c_double),
                ("p_act", C.c_double), ("p_pooled", C.c_double)]


class HubertFwdArgs(C.Structure):
    _fields_ = [("cfg", C.POINTER(HubertConfig)), ("params", C.POINTER(vp)), ("weight_planes", vp), ("wav", vp), ("ld_wav", i64),
                ("batch", i32), ("samples", i32), ("grad_batch", i32), ("mask_time", vp), ("drop_seq_key", vp), ("drop_seq_row", vp),
                ("num_segments", i32), ("segment_start", vp), ("layer_skip", vp), ("logits", vp), ("feat", vp), ("workspace", vp),
                ("workspace_bytes", i64), ("gemm_impl", i32)]


class HubertBwdArgs(C.Structure):
    _fields_ = [("cfg", C.POINTER(HubertConfig)), ("params", C.POINTER(vp)), ("weight_planes", vp), ("wav", vp), ("ld_wav", i64),
                ("batch", i32), ("samples", i32), ("grad_batch", i32), ("mask_time", vp), ("drop_seq_key", vp), ("drop_seq_row", vp),
                ("num_segments", i32), ("segment_start", vp), ("layer_skip", vp), ("dlogits", vp), ("dfeat", vp), ("grads", C.POINTER(vp)),
                ("accumulate_grads", i32), ("workspace", vp), ("workspace_bytes", i64), ("gemm_impl", i32)]


class WrnConfig(C.Structure):
    _fields_ = [("num_classes", i32), ("depth", i32), ("widen", i32), ("img_size", i32), ("bn_momentum", f32), ("slope", f32)]


class WrnFwdArgs(C.Structure):
    _fields_ = [("cfg", C.POINTER(WrnConfig)), ("params", C.POINTER(vp)), ("bn_running_mean", C.POINTER(vp)), ("bn_running_var", C.POINTER(vp)),
                ("bn_num_batches_tracked", C.POINTER(vp)), ("weight_planes", vp), ("x", vp), ("batch", i32), ("training", i32), ("stat_repeats", i32),
                ("logits", vp), ("feat", vp), ("workspace", vp), ("workspace_bytes", i64), ("gemm_impl", i32),
                ("sync_fn", vp), ("sync_ctx", vp), ("sync_buf", vp), ("world_size", i32)]


class WrnBwdArgs(C.Structure):
    _fields_ = [("cfg", C.POINTER(WrnConfig)), ("params", C.POINTER(vp)), ("weight_planes", vp), ("batch", i32), ("grad_rows", i32), ("dlogits", vp),
                ("dfeat", vp), ("grads", C.POINTER(vp)), ("accumulate_grads", i32), ("workspace", vp), ("workspace_bytes", i64), ("gemm_impl", i32),
                ("sync_fn", vp), ("sync_ctx", vp), ("sync_buf", vp), ("world_size", i32)]


ALLREDUCE_SUM_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)   # srw_allreduce_sum_fn


class RewarderFwdArgs(C.Structure):
    _fields_ = [("B", i32), ("feature_dim", i32), ("label_rows", i32), ("rp", C.POINTER(vp)), ("feats", vp), ("ld_feats", i64),
                ("labels", vp), ("reward", vp), ("workspace", vp)]


class GeneratorFwdArgs(C.Structure):
    _fields_ = [("B", i32), ("feature_dim", i32), ("gp", C.POINTER(vp)), ("feats", vp), ("ld_feats", i64), ("labels", vp),
                ("workspace", vp)]


class RewarderTrainArgs(C.Structure):
    _fields_ = [("B", i32), ("feature_dim", i32), ("label_rows", i32), ("num_classes", i32), ("rp", C.POINTER(vp)),
                ("g", C.POINTER(vp)), ("m", C.POINTER(vp)), ("v", C.POINTER(vp)), ("feats", vp), ("ld_feats", i64),
                ("gen_labels", vp), ("true_labels", vp), ("lr", f32), ("step", i32), ("phase", i32), ("losses", vp), ("workspace", vp),
                ("loss_select", i32), ("g_add", C.POINTER(vp))]


class FlexMatchMaskArgs(C.Structure):
    _fields_ = [("B", i32), ("num_classes", i32), ("ulb_dest_len", i32), ("logits_w", vp), ("ld_logits", i64), ("idx_ulb", vp),
                ("p_cutoff", f32), ("thresh_warmup", i32), ("selected_label", vp), ("hist", vp), ("classwise_acc", vp),
                ("probs_w", vp), ("pseudo", vp), ("mask", vp), ("max_probs", vp)]


class SslLossArgs(C.Structure):
    _fields_ = [("B_lb", i32), ("B_ulb", i32), ("num_classes", i32), ("logits_lb", vp), ("logits_s", vp), ("ld_logits", i64),
                ("y_lb", vp), ("pseudo", vp), ("mask", vp), ("reward", vp), ("lambda_u", f32), ("mask2", vp), ("losses", vp),
                ("dlogits_lb", vp), ("dlogits_s", vp), ("ld_dlogits", i64),
                ("rp", C.POINTER(vp)), ("feats", vp), ("ld_feats", i64), ("feature_dim", i32), ("label_rows", i32), ("rew_workspace", vp), ("reward_out", vp)]


f64 = C.c_double


class FreeMatchMaskArgs(C.Structure):
    _fields_ = [("B", i32), ("num_classes", i32), ("logits_w", vp), ("ld_logits", i64), ("momentum", C.c_double), ("use_quantile", i32),
                ("clip_thresh", i32), ("time_p", vp), ("p_model", vp), ("label_hist", vp), ("probs_w", vp), ("pseudo", vp),
                ("pseudo_from_probs", i32), ("mask", vp), ("max_probs", vp), ("phase", i32), ("probs_all", vp), ("B_all", i32)]


class FreeMatchEntropyArgs(C.Structure):
    _fields_ = [("B", i32), ("num_classes", i32), ("mask", vp), ("logits_s", vp), ("ld_logits", i64), ("p_model", vp),
                ("label_hist", vp), ("lambda_e", f32), ("losses", vp), ("dlogits_s", vp), ("ld_dlogits", i64), ("accumulate", i32)]


class SoftMatchMaskArgs(C.Structure):
    _fields_ = [("B", i32), ("num_classes", i32), ("logits_w", vp), ("ld_logits", i64), ("momentum", C.c_double), ("n_sigma", i32),
                ("dist_align", i32), ("da_p_model", vp), ("da_p_target", vp), ("da_initialized", vp), ("prob_max_mu_t", vp),
                ("prob_max_var_t", vp), ("probs_w", vp), ("probs_aligned", vp), ("pseudo", vp), ("pseudo_from_probs", i32),
                ("mask", vp), ("max_probs", vp), ("phase", i32), ("probs_all", vp), ("B_all", i32), ("maxp_all", vp), ("n_all", i32)]


class AdamWRow(C.Structure):
    _fields_ = [("param", vp), ("grad", vp), ("exp_avg", vp), ("exp_avg_sq", vp), ("planes", vp), ("numel", i64),
                ("plane_stride", i64), ("cols", i32), ("ldp", i32), ("lr", f64), ("weight_decay", f64), ("first_block", i64)]


class AdamWArgs(C.Structure):
    _fields_ = [("num_tensors", i32), ("total_blocks", i64), ("table", vp), ("lr_factor", f64), ("beta1", f64), ("beta2", f64),
                ("eps", f64), ("step", i32), ("decoupled", i32)]


class SgdArgs(C.Structure):
    _fields_ = [("num_tensors", i32), ("total_blocks", i64), ("table", vp), ("lr_factor", f64), ("momentum", f64), ("nesterov", i32), ("first_step", i32)]


class EmaRow(C.Structure):
    _fields_ = [("param", vp), ("shadow", vp), ("numel", i64), ("first_block", i64)]


class EmaArgs(C.Structure):
    _fields_ = [("num_tensors", i32), ("total_blocks", i64), ("table", vp), ("decay", f64)]


ADAMW_BLOCK_ELEMS = 4096
PROF_GEMM, PROF_ATTN_FWD, PROF_ATTN_BWD, PROF_ADAMW, PROF_NUM = 0, 1, 2, 3, 4


class AugOpDesc(C.Structure):
    _fields_ = [("op", C.c_int32), ("ival", C.c_int32), ("alpha", f32), ("identity", C.c_int32), ("a", C.c_double * 6)]


class AugSample(C.Structure):
    _fields_ = [("src_index", i64), ("crop_top", C.c_int32), ("crop_left", C.c_int32), ("flip", C.c_int32), ("n_ops", C.c_int32),
                ("ops", AugOpDesc * 3), ("cut_x0", C.c_int32), ("cut_y0", C.c_int32), ("cut_x1", C.c_int32), ("cut_y1", C.c_int32)]


class AugmentArgs(C.Structure):
    _fields_ = [("src", vp), ("n_src", i64), ("img_size", i32), ("padding", i32), ("samples", vp), ("n", i32), ("mean", f32 * 3),
                ("std", f32 * 3), ("out", vp), ("out_u8", vp)]


class ProfileStats(C.Structure):
    _fields_ = [("launches", i64), ("total_ms", f64), ("flops", f64), ("bytes", f64)]


# every symbol include/srw.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("srw_version", i32, []),
    ("srw_last_error", C.c_char_p, []),
    ("srw_device_check", i32, [C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    ("srw_kernel_launches", i64, []),
    ("srw_profile_enable", i32, [i32]),
    ("srw_profile_collect", i32, [vp]),
    ("srw_split_planes", i32, [C.POINTER(SplitArgs), vp]),
    ("srw_gemm", i32, [C.POINTER(GemmArgs), vp]),
    ("srw_splitk_reduce", i32, [C.POINTER(SplitKReduceArgs), vp]),
    ("srw_colsum", i32, [C.POINTER(ColsumArgs), vp]),
    ("srw_colsum_nparts", i32, [i32]),
    ("srw_layernorm_bwd_nparts", i32, [i32]),
    ("srw_grad_fold", i32, [C.POINTER(GradFoldArgs), vp]),
    ("srw_layernorm_fwd", i32, [C.POINTER(LayerNormFwdArgs), vp]),
    ("srw_layernorm_bwd", i32, [C.POINTER(LayerNormBwdArgs), vp]),
    ("srw_attn_fwd", i32, [C.POINTER(AttnFwdArgs), vp]),
    ("srw_attn_bwd", i32, [C.POINTER(AttnBwdArgs), vp]),
    ("srw_attn_mask_prepare", i32, [vp, i32, i32, vp, i64, vp, vp]),
    ("srw_vit_weight_planes_bytes", i64, [C.POINTER(VitConfig)]),
    ("srw_vit_workspace_bytes", i64, [C.POINTER(VitConfig), i32, i32]),
    ("srw_vit_weight_plane_slot", i32, [C.POINTER(VitConfig), i32, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]),
    ("srw_vit_prepare_weights", i32, [C.POINTER(VitConfig), C.POINTER(vp), vp, vp]),
    ("srw_vit_forward", i32, [C.POINTER(VitFwdArgs), vp]),
    ("srw_vit_backward", i32, [C.POINTER(VitBwdArgs), vp]),
    ("srw_bert_weight_planes_bytes", i64, [C.POINTER(BertConfig)]),
    ("srw_bert_workspace_bytes", i64, [C.POINTER(BertConfig), i32, i32, i32]),
    ("srw_bert_weight_plane_slot", i32, [C.POINTER(BertConfig), i32, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]),
    ("srw_bert_prepare_weights", i32, [C.POINTER(BertConfig), C.POINTER(vp), vp, vp]),
    ("srw_bert_forward", i32, [C.POINTER(BertFwdArgs), vp]),
    ("srw_bert_backward", i32, [C.POINTER(BertBwdArgs), vp]),
    ("srw_hubert_frames", i32, [C.POINTER(HubertConfig), i32]),
    ("srw_hubert_weight_planes_bytes", i64, [C.POINTER(HubertConfig)]),
    ("srw_hubert_workspace_bytes", i64, [C.POINTER(HubertConfig), i32, i32, i32]),
    ("srw_hubert_prepare_weights", i32, [C.POINTER(HubertConfig), C.POINTER(vp), vp, vp]),
    ("srw_hubert_prepare_front", i32, [C.POINTER(HubertConfig), C.POINTER(vp), vp, vp]),
    ("srw_hubert_weight_plane_slot", i32, [C.POINTER(HubertConfig), i32, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]),
    ("srw_hubert_forward", i32, [C.POINTER(HubertFwdArgs), vp]),
    ("srw_hubert_backward", i32, [C.POINTER(HubertBwdArgs), vp]),
    ("srw_wrn_num_params", i32, [C.POINTER(WrnConfig)]),
    ("srw_wrn_weight_planes_bytes", i64, [C.POINTER(WrnConfig)]),
    ("srw_wrn_workspace_bytes", i64, [C.POINTER(WrnConfig), i32]),
    ("srw_wrn_prepare_weights", i32, [C.POINTER(WrnConfig), C.POINTER(vp), vp, vp]),
    ("srw_wrn_forward", i32, [C.POINTER(WrnFwdArgs), vp]),
    ("srw_wrn_backward", i32, [C.POINTER(WrnBwdArgs), vp]),
    ("srw_sgd_step", i32, [C.POINTER(SgdArgs), vp]),
    ("srw_set_graph_mode", i32, [i32]),
    ("srw_set_pdl_mode", i32, [i32]),
    ("srw_scale_inplace", i32, [vp, i64, vp, vp]),
    ("srw_rewarder_workspace_floats", i64, [i32, i32]),
    ("srw_rewarder_fwd", i32, [C.POINTER(RewarderFwdArgs), vp]),
    ("srw_generator_fwd", i32, [C.POINTER(GeneratorFwdArgs), vp]),
    ("srw_rewarder_train", i32, [C.POINTER(RewarderTrainArgs), vp]),
    ("srw_flexmatch_mask", i32, [C.POINTER(FlexMatchMaskArgs), vp]),
    ("srw_ssl_loss", i32, [C.POINTER(SslLossArgs), vp]),
    ("srw_freematch_mask", i32, [C.POINTER(FreeMatchMaskArgs), vp]),
    ("srw_freematch_entropy", i32, [C.POINTER(FreeMatchEntropyArgs), vp]),
    ("srw_softmatch_mask", i32, [C.POINTER(SoftMatchMaskArgs), vp]),
    ("srw_adamw_step", i32, [C.POINTER(AdamWArgs), vp]),
    ("srw_ema_step", i32, [C.POINTER(EmaArgs), vp]),
    ("srw_augment_batch", i32, [C.POINTER(AugmentArgs), vp]),
]

_lib = None


class SrwError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the library (building is __graft_entry__.build()'s job; a missing .so is a hard error)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SrwError(f"{LIB_PATH} not found: run `python -m semireward_b200.build` (or __graft_entry__.build()). "
                       "semireward_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().srw_last_error()
        raise SrwError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr_array(tensors):
    arr = (vp * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
