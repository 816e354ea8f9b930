"""ctypes binding of libstreamz_b200.so (the C ABI declared in include/streamz_b200.h).

There is no fallback of any kind: if the shared library is missing this module raises, and if no sm_100 GPU is
present ``szb_ctx_create`` fails with SZB_ERR_NO_DEVICE.  Nothing under ``oracle/`` is ever imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libstreamz_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_ALLOC, ERR_NCCL, ERR_IO, ERR_UNSUPPORTED = range(8)
STATUS_NAMES = ["OK", "INVALID", "CUDA", "NO_DEVICE", "ALLOC", "NCCL", "IO", "UNSUPPORTED"]


class StreamzError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"streamz_b200: SZB_ERR_{STATUS_NAMES[status] if 0 <= status < 8 else status}: {message}")
        self.status = status


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C streamz_b200/csrc`).  streamz_b200 has no CPU or PyTorch fallback.")
    return C.CDLL(LIB_PATH)


lib = _load()

vp, i32, u32, u64, f32, f64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_float, C.c_double
sz = C.c_size_t
P = C.POINTER

# name -> (restype, argtypes).  Keep in the order of include/streamz_b200.h; tests check the two agree.
SIGNATURES = {
    "szb_version": (C.c_char_p, []),
    "szb_last_error": (C.c_char_p, []),
    "szb_ctx_create": (i32, [i32, vp, P(vp)]),
    "szb_ctx_destroy": (None, [vp]),
    "szb_ctx_sync": (i32, [vp]),
    "szb_ctx_sm_count": (i32, [vp]),
    "szb_ctx_set_fused_resample": (i32, [vp, i32]),
    "szb_ctx_set_l2_ring": (i32, [vp, i32, i32]),
    "szb_ctx_launch_count": (u64, [vp]),
    "szb_ctx_graph_launch_count": (u64, [vp]),
    "szb_timer_start": (i32, [vp]),
    "szb_timer_stop": (i32, [vp, P(f32)]),
    "szb_kernel_timing": (i32, [vp, i32]),
    "szb_kernel_timing_read": (i32, [vp, P(f64), P(u64), i32]),
    "szb_dev_alloc": (i32, [vp, sz, P(vp)]),
    "szb_dev_free": (i32, [vp, vp]),
    "szb_memcpy_h2d": (i32, [vp, vp, vp, sz]),
    "szb_memcpy_d2h": (i32, [vp, vp, vp, sz]),
    "szb_host_alloc_pinned": (i32, [sz, P(vp)]),
    "szb_host_free_pinned": (i32, [vp]),
    "szb_table_mel": (i32, [vp]),
    "szb_table_dct": (i32, [vp]),
    "szb_table_resample_taps": (i32, [u32, vp, P(u32), P(u32)]),
    "szb_num_windows": (u64, [u64]),
    "szb_resample_out_len": (u64, [u64, u32]),
    "szb_downmix_to_mono": (i32, [vp, vp, u64, u32, vp, u64, P(u64)]),
    "szb_augment_params": (i32, [u64, u64, P(f32), P(f32), P(u64)]),
    "szb_augment": (i32, [vp, vp, u64, u64, vp]),
    "szb_augment_dev": (i32, [vp, vp, u64, u64, vp]),
    "szb_resample_to_44100": (i32, [vp, vp, u64, u32, vp, u64, P(u64)]),
    "szb_extract": (i32, [vp, vp, u64, vp, u64, P(u64)]),
    "szb_extract_range": (i32, [vp, vp, u64, u64, u64, vp, u64]),
    "szb_extract_range_dev": (i32, [vp, vp, u64, u64, u64, vp, u64]),
    "szb_extract_batch": (i32, [vp, vp, vp, u32, u32, vp, u64, vp]),
    "szb_extract_batch_dev": (i32, [vp, vp, vp, u32, u32, vp, u64, vp]),
    "szb_extract_batch_windows": (u64, [vp, u32, u32]),
    "szb_net_create": (i32, [vp, u32, u32, u32, u32, u64, P(vp)]),
    "szb_net_from_weights": (i32, [vp, u32, u32, u32, u32, vp, vp, vp, vp, vp, vp, P(vp)]),
    "szb_net_get_weights": (i32, [vp, vp, vp, vp, vp, vp, vp]),
    "szb_net_dims": (i32, [vp, P(u32 * 4)]),
    "szb_net_output_size": (u32, [vp]),
    "szb_net_add_output_class": (i32, [vp, vp, u64]),
    "szb_net_destroy": (None, [vp]),
    "szb_net_set_precision": (i32, [vp, i32]),
    "szb_net_get_precision": (i32, [vp]),
    "szb_net_record_training_file": (i32, [vp, u32, C.c_char_p]),
    "szb_net_file_list": (i32, [vp, u32, vp, sz, P(sz)]),
    "szb_net_forward": (i32, [vp, vp, u64, vp]),
    "szb_net_forward_dev": (i32, [vp, vp, u64, vp]),
    "szb_net_train_batch": (i32, [vp, vp, u64, vp, f32]),
    "szb_net_train_batch_labels": (i32, [vp, vp, vp, u64, f32, vp, P(f64), P(u64)]),
    "szb_net_train_epoch_dev": (i32, [vp, vp, vp, u64, vp, u64, u32, f32, f32, u64, u64, vp, P(f64), P(u64)]),
    "szb_net_train_epoch_steps_dev": (i32, [vp, vp, vp, u64, vp, u64, vp, u32, f32, f32, u64, u64, vp, P(f64), P(u64)]),
    "szb_dropout_keep_mask": (i32, [u64, u64, vp, u64, u32, f32, vp]),
    "szb_loop_seed": (u64, [u64, u32, u32]),
    "szb_shuffle_perm": (i32, [u64, u64, u64, vp]),
    "szb_lr_decay": (f32, [f32, i32]),
    "szb_net_pretrain_network": (i32, [vp, vp, u64, u32, u32, f32, f32, u32, u64, P(f64), P(u64)]),
    "szb_net_train_from_files": (i32, [vp, vp, vp, vp, u32, u32, f32, f32, u32, u64, P(f64), P(u64)]),
    "szb_identify_counts": (i32, [vp, vp, u64, f32, vp]),
    "szb_identify_counts_dev": (i32, [vp, vp, u64, f32, vp]),
    "szb_identify_counts_batch_dev": (i32, [vp, vp, vp, u32, f32, vp]),
    "szb_identify_sums": (i32, [vp, vp, u64, vp]),
    "szb_identify_speaker_list": (i32, [vp, vp, u64, f32, vp, u32, P(u32)]),
    "szb_net_embedding_size": (i32, [vp, P(u32)]),
    "szb_net_embed": (i32, [vp, vp, u64, i32, vp]),
    "szb_net_embedding_mean": (i32, [vp, vp, u64, vp]),
    "szb_net_embedding_median": (i32, [vp, vp, u64, i32, vp]),
    "szb_cosine_similarity": (f32, [vp, vp, u32]),
    "szb_match_embedding": (i32, [vp, vp, vp, u32, u32, f32, P(u64), P(f32)]),
    "szb_comm_unique_id": (i32, [vp]),
    "szb_comm_init": (i32, [vp, vp, i32, i32]),
    "szb_comm_destroy": (i32, [vp]),
    "szb_comm_world": (i32, [vp]),
    "szb_comm_peer_exchange": (i32, [vp, i32, P(i32)]),
    "szb_comm_peer_trace": (i32, [vp, i32, vp]),
    "szb_feature_cache_path": (i32, [C.c_char_p, vp, sz]),
    "szb_npy_write_f32": (i32, [C.c_char_p, vp, u64, u64]),
    "szb_npy_read_f32": (i32, [C.c_char_p, vp, u64, P(u64), P(u64)]),
    "szb_net_save": (i32, [vp, C.c_char_p, u32, u32]),
    "szb_net_load": (i32, [vp, C.c_char_p, P(vp), P(u32), P(u32)]),
    "szb_net_set_embeddings": (i32, [vp, vp, vp, vp, u32, u32]),
    "szb_net_get_embeddings": (i32, [vp, vp, vp, vp, u32, P(u32), P(u32)]),
    "szb_net_set_encoding_layer": (i32, [vp, vp, vp, u32, u32]),
    "szb_net_get_encoding_layer": (i32, [vp, vp, vp, u64, P(u32), P(u32)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


def check(status: int) -> None:
    if status != OK:
        raise StreamzError(status, (lib.szb_last_error() or b"").decode("utf-8", "replace"))


def ptr(a) -> C.c_void_p:
    """Raw pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    return C.c_void_p(a.ctypes.data)
