"""Python mirror of the reference's public Rust API for the hot path, over the C ABI (include/streamz_b200.h).

Names, argument meaning and error behaviour follow ``streamz_rs`` (streamz-rs/src/lib.rs) so the parity tests read
like the reference's own usage (main.rs:500-508, 658-666):

    FeatureExtractor::new / extract            lib.rs:239-263
    resample_to_44100 / downmix_to_mono        lib.rs:172-209
    SimpleNeuralNet::{new, forward, train, train_batch, output_size, add_output_class, save, load}   lib.rs:767-1282
    pretrain_from_features / train_from_feature_map                                                   lib.rs:582-665
    identify_speaker / _with_threshold / _with_threshold_feats / identify_speaker_list                lib.rs:1285-1411
    feature_cache_path / load_cached_features                                                         lib.rs:550-579

Every numeric operation runs in the CUDA library; this module only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N

DEFAULT_SAMPLE_RATE = 44100
WINDOW_SIZE = 800
MFCC_SIZE = 20
FEATURE_SIZE = 60
DEFAULT_DROPOUT = 0.2


class Context:
    """One CUDA stream + scratch on one device (szb_ctx).  Use from one thread at a time."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        h = C.c_void_p()
        N.check(N.lib.szb_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            N.lib.szb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("context is closed")
        return self._h

    def sync(self):
        N.check(N.lib.szb_ctx_sync(self.handle))

    @property
    def sm_count(self) -> int:
        return int(N.lib.szb_ctx_sm_count(self.handle))

    @property
    def launch_count(self) -> int:
        return int(N.lib.szb_ctx_launch_count(self.handle))

    def timer_start(self):
        N.check(N.lib.szb_timer_start(self.handle))

    def timer_stop(self) -> float:
        ms = C.c_float()
        N.check(N.lib.szb_timer_stop(self.handle, C.byref(ms)))
        return float(ms.value)

    def kernel_timing(self, enable: bool):
        N.check(N.lib.szb_kernel_timing(self.handle, 1 if enable else 0))

    def kernel_timing_read(self, reset: bool = True) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_uint64()
        N.check(N.lib.szb_kernel_timing_read(self.handle, C.byref(ms), C.byref(n), 1 if reset else 0))
        return float(ms.value), int(n.value)

    # device memory for the *_dev entry points
    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        N.check(N.lib.szb_dev_alloc(self.handle, int(nbytes), C.byref(p)))
        return int(p.value)

    def dev_free(self, dptr: int):
        N.check(N.lib.szb_dev_free(self.handle, C.c_void_p(dptr)))

    def h2d(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        N.check(N.lib.szb_memcpy_h2d(self.handle, C.c_void_p(dptr), N.ptr(arr), arr.nbytes))

    def d2h(self, arr: np.ndarray, dptr: int):
        assert arr.flags.c_contiguous
        N.check(N.lib.szb_memcpy_d2h(self.handle, N.ptr(arr), C.c_void_p(dptr), arr.nbytes))

    # multi-GPU
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        N.check(N.lib.szb_comm_init(self.handle, buf, int(rank), int(world)))


def _ctx_peer_exchange(self, enable, mode: str = "auto") -> bool:
    """Collective: switch the gradient exchange between NCCL all-reduces (default) and the all-reduce over NVLink peer memory
    fused into the update kernel (include/streamz_b200.h, szb_comm_peer_exchange); mode "auto" | "one-shot" | "two-shot" | "ll" (one-shot with the flag inside every 8-byte packet).
    Returns whether the peer exchange is active."""
    active = C.c_int32()
    code = {"auto": 1, "one-shot": 2, "two-shot": 3, "ll": 4}[mode] if enable else 0
    N.check(N.lib.szb_comm_peer_exchange(self.handle, code, C.byref(active)))
    return bool(active.value)


Context.comm_peer_exchange = _ctx_peer_exchange


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    N.check(N.lib.szb_comm_unique_id(buf))
    return bytes(buf)


_tls = threading.local()


def default_context() -> Context:
    """Thread-local context on cuda:0, the analogue of the reference's thread-local extractor (lib.rs:266-276)."""
    ctx = getattr(_tls, "ctx", None)
    if ctx is None:
        ctx = _tls.ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return ctx


def _i16(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int16))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def num_windows(n_samples: int) -> int:
    return int(N.lib.szb_num_windows(int(n_samples)))


def downmix_to_mono(samples, channels: int, ctx: Optional[Context] = None) -> np.ndarray:
    """lib.rs:172-183."""
    ctx = ctx or default_context()
    s = _i16(samples)
    ch = max(1, int(channels))
    out = np.empty((len(s) + ch - 1) // ch, dtype=np.int16)
    n = C.c_uint64()
    N.check(N.lib.szb_downmix_to_mono(ctx.handle, N.ptr(s), len(s), int(channels), N.ptr(out), len(out), C.byref(n)))
    return out[: n.value]


def augment_params(seed: int, n_samples: int) -> Tuple[float, float, int]:
    nl, gain, shift = C.c_float(), C.c_float(), C.c_uint64()
    N.check(N.lib.szb_augment_params(int(seed), int(n_samples), C.byref(nl), C.byref(gain), C.byref(shift)))
    return float(nl.value), float(gain.value), int(shift.value)


def augment(samples, seed: int = 0, ctx: Optional[Context] = None) -> np.ndarray:
    """lib.rs:103-116 with every random draw derived from ``seed``."""
    ctx = ctx or default_context()
    s = _i16(samples)
    out = np.empty_like(s)
    N.check(N.lib.szb_augment(ctx.handle, N.ptr(s), len(s), int(seed), N.ptr(out)))
    return out


def resample_to_44100(samples, from_rate: int, ctx: Optional[Context] = None) -> np.ndarray:
    """lib.rs:186-209 (returns a new i16 array; rate 44100 is a copy)."""
    ctx = ctx or default_context()
    s = _i16(samples)
    out = np.empty(int(N.lib.szb_resample_out_len(len(s), int(from_rate))) if from_rate != DEFAULT_SAMPLE_RATE else len(s),
                   dtype=np.int16)
    n = C.c_uint64()
    N.check(N.lib.szb_resample_to_44100(ctx.handle, N.ptr(s), len(s), int(from_rate), N.ptr(out), len(out), C.byref(n)))
    return out[: n.value]


class FeatureExtractor:
    """lib.rs:231-264.  The tables live in the GPU's constant memory (uploaded when the context is created)."""

    def __init__(self, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()

    def extract(self, samples) -> np.ndarray:
        """``extract(&self, &[i16]) -> Vec<Vec<f32>>``: ``[n, 60]`` float32; fewer than 800 samples -> ``[0, 60]``."""
        s = _i16(samples)
        n = num_windows(len(s))
        out = np.empty((n, FEATURE_SIZE), dtype=np.float32)
        got = C.c_uint64()
        N.check(N.lib.szb_extract(self.ctx.handle, N.ptr(s), len(s), N.ptr(out), n, C.byref(got)))
        assert got.value == n
        return out

    def extract_range(self, samples, w_begin: int, w_end: int) -> np.ndarray:
        """Rows ``w_begin .. w_end - 1`` of :meth:`extract` on the whole clip, bit for bit, computed from the samples those
        windows (and the two frames either side that the delta stencil reads) cover: the per-rank unit of the multi-GPU
        identification sweep."""
        s = _i16(samples)
        n = max(0, int(w_end) - int(w_begin))
        out = np.empty((n, FEATURE_SIZE), dtype=np.float32)
        N.check(N.lib.szb_extract_range(self.ctx.handle, N.ptr(s), len(s), int(w_begin), int(w_end), N.ptr(out), n))
        return out

    def extract_batch(self, clips: Sequence[np.ndarray], rate: int = DEFAULT_SAMPLE_RATE) -> List[np.ndarray]:
        """The rayon loop of main.rs:500-508 (and batch_resample, lib.rs:541-547, when ``rate != 44100``)."""
        feats, win_off = self.extract_packed(*pack_clips(clips), rate=rate)
        return [feats[win_off[i]: win_off[i + 1]] for i in range(len(clips))]

    def extract_packed(self, pcm: np.ndarray, clip_off: np.ndarray, rate: int = DEFAULT_SAMPLE_RATE,
                       out: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
        pcm = _i16(pcm)
        clip_off = np.ascontiguousarray(clip_off, dtype=np.uint64)
        n_clips = len(clip_off) - 1
        total = int(N.lib.szb_extract_batch_windows(N.ptr(clip_off), n_clips, int(rate)))
        if out is None:
            out = np.empty((total, FEATURE_SIZE), dtype=np.float32)
        win_off = np.zeros(n_clips + 1, dtype=np.uint64)
        N.check(N.lib.szb_extract_batch(self.ctx.handle, N.ptr(pcm), N.ptr(clip_off), n_clips, int(rate), N.ptr(out),
                                        out.shape[0], N.ptr(win_off)))
        return out[:total], win_off.astype(np.int64)


def pack_clips(clips: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate clips back to back; returns (pcm, clip_off[n+1]) in samples."""
    off = np.zeros(len(clips) + 1, dtype=np.uint64)
    pos = 0
    for i, c in enumerate(clips):
        pos += len(c)
        off[i + 1] = pos
    pcm = np.concatenate([_i16(c) for c in clips]) if clips else np.zeros(0, np.int16)
    return pcm, off


_thread_extractor = threading.local()


def with_thread_extractor(f):
    """lib.rs:271-276."""
    ex = getattr(_thread_extractor, "ex", None)
    if ex is None:
        ex = _thread_extractor.ex = FeatureExtractor()
    return f(ex)


class SimpleNeuralNet:
    """lib.rs:745-1282.  Weights live on the GPU; ``w1 .. b3`` properties download them."""

    def __init__(self, input: int, hidden1: int, hidden2: int, output: int, seed: int = 0, ctx: Optional[Context] = None,
                 _handle=None):
        self.ctx = ctx or default_context()
        if _handle is not None:
            self._h = _handle
        else:
            h = C.c_void_p()
            N.check(N.lib.szb_net_create(self.ctx.handle, input, hidden1, hidden2, output, int(seed) & (2 ** 64 - 1), C.byref(h)))
            self._h = h
        self.sample_rate, self.bits = DEFAULT_SAMPLE_RATE, 16

    new = classmethod(lambda cls, *a, **k: cls(*a, **k))

    @classmethod
    def from_weights(cls, w1, b1, w2, b2, w3, b3, ctx: Optional[Context] = None) -> "SimpleNeuralNet":
        ctx = ctx or default_context()
        w1, b1, w2, b2, w3, b3 = map(_f32, (w1, b1, w2, b2, w3, b3))
        assert w1.shape[1] == b1.shape[0] == w2.shape[0] and w2.shape[1] == b2.shape[0] == w3.shape[0] and w3.shape[1] == b3.shape[0]
        h = C.c_void_p()
        N.check(N.lib.szb_net_from_weights(ctx.handle, w1.shape[0], w1.shape[1], w2.shape[1], w3.shape[1], N.ptr(w1), N.ptr(b1),
                                           N.ptr(w2), N.ptr(b2), N.ptr(w3), N.ptr(b3), C.byref(h)))
        return cls(0, 0, 0, 0, ctx=ctx, _handle=h)

    def close(self):
        if getattr(self, "_h", None):
            N.lib.szb_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def dims(self) -> Tuple[int, int, int, int]:
        d = (C.c_uint32 * 4)()
        N.check(N.lib.szb_net_dims(self._h, C.byref(d)))
        return tuple(int(x) for x in d)

    def output_size(self) -> int:
        return int(N.lib.szb_net_output_size(self._h))

    PRECISIONS = {"fp32": 0, "3xtf32": 1, "tf32": 2}

    def set_precision(self, mode) -> "SimpleNeuralNet":
        """'3xtf32' (default, tensor cores, FP32-equivalent), 'tf32' (tensor cores, fastest) or 'fp32' (CUDA cores)."""
        N.check(N.lib.szb_net_set_precision(self._h, self.PRECISIONS.get(mode, mode)))
        return self

    def weights(self):
        i, h1, h2, c = self.dims
        w1, b1 = np.empty((i, h1), np.float32), np.empty(h1, np.float32)
        w2, b2 = np.empty((h1, h2), np.float32), np.empty(h2, np.float32)
        w3, b3 = np.empty((h2, c), np.float32), np.empty(c, np.float32)
        N.check(N.lib.szb_net_get_weights(self._h, N.ptr(w1), N.ptr(b1), N.ptr(w2), N.ptr(b2), N.ptr(w3), N.ptr(b3)))
        return w1, b1, w2, b2, w3, b3

    def add_output_class(self, new_col=None, seed: int = 0):
        col = None if new_col is None else _f32(new_col)
        N.check(N.lib.szb_net_add_output_class(self._h, N.ptr(col), int(seed)))

    def record_training_file(self, speaker: int, path: str):
        N.check(N.lib.szb_net_record_training_file(self._h, int(speaker), path.encode()))

    def file_lists(self) -> List[List[str]]:
        out = []
        for s in range(self.output_size()):
            n = C.c_size_t()
            N.check(N.lib.szb_net_file_list(self._h, s, None, 0, C.byref(n)))
            buf = C.create_string_buffer(n.value + 1)
            N.check(N.lib.szb_net_file_list(self._h, s, buf, n.value + 1, C.byref(n)))
            text = buf.value.decode()
            out.append(text.split("\n") if text else [])
        return out

    def forward(self, x) -> np.ndarray:
        """``forward(&self, &[f32]) -> Vec<f32>`` for one window, or batched over the rows of a 2-D array."""
        x = _f32(x)
        single = x.ndim == 1
        x2 = x.reshape(1, -1) if single else x
        assert x2.shape[1] == self.dims[0], "input size mismatch"   # ndarray's dot panics on a shape mismatch
        probs = np.empty((x2.shape[0], self.output_size()), dtype=np.float32)
        N.check(N.lib.szb_net_forward(self._h, N.ptr(x2), x2.shape[0], N.ptr(probs)))
        return probs[0] if single else probs

    def train(self, x, target, lr: float):
        """lib.rs:954-999: single-sample SGD = train_batch with one row."""
        self.train_batch(_f32(x).reshape(1, -1), target, lr)

    def train_batch(self, batch, target, lr: float):
        """``train_batch(&mut self, &[Vec<f32>], &[f32], f32)`` (lib.rs:1002): one target vector for the whole batch."""
        x = _f32(batch).reshape(-1, self.dims[0]) if len(batch) else np.zeros((0, self.dims[0]), np.float32)
        t = _f32(target)
        assert t.shape[0] == self.output_size(), "target size mismatch"
        N.check(N.lib.szb_net_train_batch(self._h, N.ptr(x), x.shape[0], N.ptr(t), float(lr)))

    def train_batch_labels(self, batch, labels, lr: float, keep=None) -> Tuple[float, int]:
        x = _f32(batch).reshape(-1, self.dims[0])
        lab = np.ascontiguousarray(labels, dtype=np.uint32)
        k = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
        loss, used = C.c_double(), C.c_uint64()
        N.check(N.lib.szb_net_train_batch_labels(self._h, N.ptr(x), N.ptr(lab), x.shape[0], float(lr), N.ptr(k), C.byref(loss),
                                                 C.byref(used)))
        return float(loss.value), int(used.value)

    def set_embeddings(self, embeds: Sequence[Tuple[np.ndarray, float, float]]):
        """lib.rs:869-872: (embedding, mean similarity, std similarity) per speaker; saved into model.npz."""
        n = len(embeds)
        e = _f32(np.stack([np.asarray(v[0], np.float32) for v in embeds])) if n else np.zeros((0, 0), np.float32)
        m = _f32([v[1] for v in embeds])
        sd = _f32([v[2] for v in embeds])
        N.check(N.lib.szb_net_set_embeddings(self._h, N.ptr(e) if n else None, N.ptr(m) if n else None, N.ptr(sd) if n else None, n,
                                             e.shape[1] if n else 0))

    def embeddings(self) -> List[Tuple[np.ndarray, float, float]]:
        """lib.rs:874-877."""
        n, dim = C.c_uint32(), C.c_uint32()
        N.check(N.lib.szb_net_get_embeddings(self._h, None, None, None, 0, C.byref(n), C.byref(dim)))
        if n.value == 0:
            return []
        e, m, sd = np.empty((n.value, dim.value), np.float32), np.empty(n.value, np.float32), np.empty(n.value, np.float32)
        N.check(N.lib.szb_net_get_embeddings(self._h, N.ptr(e), N.ptr(m), N.ptr(sd), n.value, C.byref(n), C.byref(dim)))
        return [(e[i].copy(), float(m[i]), float(sd[i])) for i in range(n.value)]

    def set_encoding_layer(self, w4, b4):
        """Optional hidden encoding layer (lib.rs:752-754): carried through save / load only."""
        b = _f32(b4)
        w = _f32(w4).reshape(-1, max(1, len(b))) if len(b) else np.zeros((0, 0), np.float32)
        N.check(N.lib.szb_net_set_encoding_layer(self._h, N.ptr(w) if len(b) else None, N.ptr(b) if len(b) else None,
                                                 w.shape[0] if len(b) else 0, len(b)))

    def encoding_layer(self):
        rows, n = C.c_uint32(), C.c_uint32()
        N.check(N.lib.szb_net_get_encoding_layer(self._h, None, None, 0, C.byref(rows), C.byref(n)))
        if n.value == 0:
            return None
        w, b = np.empty((rows.value, n.value), np.float32), np.empty(n.value, np.float32)
        N.check(N.lib.szb_net_get_encoding_layer(self._h, N.ptr(w), N.ptr(b), w.size, C.byref(rows), C.byref(n)))
        return w, b

    def save(self, path: str):
        N.check(N.lib.szb_net_save(self._h, path.encode(), int(self.sample_rate), int(self.bits)))

    @classmethod
    def load(cls, path: str, ctx: Optional[Context] = None) -> "SimpleNeuralNet":
        ctx = ctx or default_context()
        h, sr, bits = C.c_void_p(), C.c_uint32(), C.c_uint32()
        N.check(N.lib.szb_net_load(ctx.handle, path.encode(), C.byref(h), C.byref(sr), C.byref(bits)))
        net = cls(0, 0, 0, 0, ctx=ctx, _handle=h)
        net.sample_rate, net.bits = int(sr.value), int(bits.value)
        return net


class DeviceFeatures:
    """Feature windows (and labels) resident on the GPU for epoch training (szb_net_train_epoch_dev)."""

    def __init__(self, ctx: Context, feats: np.ndarray, labels: np.ndarray):
        feats, labels = _f32(feats), np.ascontiguousarray(labels, dtype=np.uint32)
        assert feats.shape[0] == labels.shape[0]
        self.ctx, self.n, self.n_in = ctx, feats.shape[0], feats.shape[1]
        self.d_feats = ctx.dev_alloc(max(1, feats.nbytes))
        self.d_labels = ctx.dev_alloc(max(1, labels.nbytes))
        if self.n:
            ctx.h2d(self.d_feats, feats)
            ctx.h2d(self.d_labels, labels)
        self.d_keep = None

    def set_keep_mask(self, keep: Optional[np.ndarray]):
        if self.d_keep is not None:
            self.ctx.dev_free(self.d_keep)
            self.d_keep = None
        if keep is not None:
            k = np.ascontiguousarray(keep, dtype=np.uint8)
            assert k.shape == (self.n, self.n_in)
            self.d_keep = self.ctx.dev_alloc(max(1, k.nbytes))
            self.ctx.h2d(self.d_keep, k)

    def close(self):
        for name in ("d_feats", "d_labels", "d_keep"):
            p = getattr(self, name, None)
            if p is not None:
                self.ctx.dev_free(p)
                setattr(self, name, None)


def train_epoch(net: SimpleNeuralNet, data: DeviceFeatures, perm, batch: int, lr: float, dropout: float = 0.0, seed: int = 0,
                stream: int = 0) -> Tuple[float, int]:
    """One epoch of lib.rs:599-622 over device-resident windows in the order ``perm``."""
    perm = np.ascontiguousarray(perm, dtype=np.uint32)
    loss, used = C.c_double(), C.c_uint64()
    N.check(N.lib.szb_net_train_epoch_dev(net._h, C.c_void_p(data.d_feats), C.c_void_p(data.d_labels), data.n, N.ptr(perm),
                                          len(perm), int(batch), float(lr), float(dropout), int(seed), int(stream),
                                          C.c_void_p(data.d_keep) if data.d_keep is not None else None, C.byref(loss), C.byref(used)))
    return float(loss.value), int(used.value)


def train_epoch_steps(net: SimpleNeuralNet, data: DeviceFeatures, perm, step_sizes, lr: float, dropout: float = 0.0, seed: int = 0,
                      stream: int = 0) -> Tuple[float, int]:
    """Multi-GPU form of :func:`train_epoch`: ``perm`` / ``step_sizes`` are this rank's slices of every global batch
    (``sharding.shard_batches``); returns the GLOBAL loss sum and window count."""
    perm = np.ascontiguousarray(perm, dtype=np.uint32)
    sizes = np.ascontiguousarray(step_sizes, dtype=np.uint32)
    loss, used = C.c_double(), C.c_uint64()
    N.check(N.lib.szb_net_train_epoch_steps_dev(net._h, C.c_void_p(data.d_feats), C.c_void_p(data.d_labels), data.n, N.ptr(perm),
                                                len(perm), N.ptr(sizes), len(sizes), float(lr), float(dropout), int(seed), int(stream),
                                                C.c_void_p(data.d_keep) if data.d_keep is not None else None, C.byref(loss), C.byref(used)))
    return float(loss.value), int(used.value)


def dropout_keep_mask(seed: int, stream: int, rows, n_in: int, prob: float) -> np.ndarray:
    rows = np.ascontiguousarray(rows, dtype=np.uint64)
    keep = np.empty((len(rows), n_in), dtype=np.uint8)
    N.check(N.lib.szb_dropout_keep_mask(int(seed), int(stream), N.ptr(rows), len(rows), int(n_in), float(prob), N.ptr(keep)))
    return keep.astype(bool)


def pretrain_from_features(net: SimpleNeuralNet, windows, target_class: int, num_classes: int, epochs: int, lr: float,
                           dropout: float, batch_size: int, rng: Optional[np.random.Generator] = None, seed: int = 0) -> float:
    """lib.rs:582-628.  The shuffle comes from ``rng`` (the reference uses an unseeded thread_rng), the dropout
    decisions from the library's counter RNG keyed by (seed, epoch).  Returns the mean loss over surviving windows."""
    windows = _f32(windows).reshape(-1, net.dims[0])
    assert num_classes == net.output_size(), "num_classes must equal the net's output size (forward slices to it)"
    n = windows.shape[0]
    if n == 0 or epochs == 0:
        return 0.0
    rng = rng or np.random.default_rng(seed)
    labels = np.full(n, int(target_class), dtype=np.uint32)
    data = DeviceFeatures(net.ctx, windows, labels)
    total, count = 0.0, 0
    try:
        for e in range(int(epochs)):
            perm = rng.permutation(n).astype(np.uint32)           # lib.rs:600-601
            l, c = train_epoch(net, data, perm, max(1, int(batch_size)), lr, dropout, seed=seed, stream=e)
            total += l
            count += c
    finally:
        data.close()
    return total / count if count else 0.0                        # lib.rs:623-627


def pretrain_network(net: SimpleNeuralNet, samples, target_class: int, num_classes: int, epochs: int, lr: float, dropout: float,
                     batch_size: int, extractor: Optional[FeatureExtractor] = None, seed: int = 0) -> float:
    """lib.rs:348-397: every epoch augments the clip, re-extracts its windows, shuffles and trains.  One C-ABI call
    (szb_net_pretrain_network): the augmented clip, its windows and the labels stay on the GPU.  Epoch ``e`` draws its
    augmentation, shuffle and dropout decisions from ``szb_loop_seed(seed, 0, e)``.  Returns the mean loss over the
    windows used, 0.0 when there were none (lib.rs:392-396)."""
    assert num_classes == net.output_size(), "num_classes must equal the net's output size (forward slices to it)"
    s = _i16(samples)
    loss, used = C.c_double(), C.c_uint64()
    N.check(N.lib.szb_net_pretrain_network(net._h, N.ptr(s), len(s), int(target_class), int(epochs), float(lr), float(dropout),
                                           max(1, int(batch_size)), int(seed) & (2 ** 64 - 1), C.byref(loss), C.byref(used)))
    return float(loss.value) / used.value if used.value else 0.0


def train_from_files(net: SimpleNeuralNet, files: Sequence[Tuple[str, np.ndarray, int]], num_speakers: int, epochs: int, lr: float,
                     dropout: float, batch_size: int, extractor: Optional[FeatureExtractor] = None, seed: int = 0) -> float:
    """lib.rs:668-732.  ``files`` holds (path, decoded 44.1 kHz mono samples, class): decoding and resampling
    (``load_and_resample_file``, lib.rs:696) stay on the caller's side.  Every (file, epoch) is one pretrain_network epoch at
    ``lr * 0.99**step`` (lib.rs:708-709), in file-major order -- one legal serialisation of the reference's rayon loop under
    its write lock; training files are recorded (lib.rs:723) and the dataset specs set (lib.rs:703-706).  Returns the mean
    loss over all windows used (the reference discards it)."""
    assert num_speakers == net.output_size(), "num_speakers must equal the net's output size"
    files = [(p, _i16(x), int(c)) for p, x, c in files]
    if not files or epochs <= 0:
        return 0.0
    pcm, off = pack_clips([x for _, x, _ in files])
    classes = np.array([c for _, _, c in files], dtype=np.uint32)
    net.sample_rate, net.bits = DEFAULT_SAMPLE_RATE, 16
    loss, used = C.c_double(), C.c_uint64()
    N.check(N.lib.szb_net_train_from_files(net._h, N.ptr(pcm), N.ptr(off), N.ptr(classes), len(files), int(epochs), float(lr),
                                           float(dropout), max(1, int(batch_size)), int(seed) & (2 ** 64 - 1), C.byref(loss),
                                           C.byref(used)))
    for path, _, cls in files:
        net.record_training_file(cls, path)
    return float(loss.value) / used.value if used.value else 0.0


def shuffle_perm(seed: int, stream: int, n: int) -> np.ndarray:
    """The library's seeded stand-in for ``windows.shuffle(&mut thread_rng)`` (lib.rs:370, 601)."""
    perm = np.zeros(max(1, int(n)), dtype=np.uint32)
    N.check(N.lib.szb_shuffle_perm(int(seed) & (2 ** 64 - 1), int(stream), int(n), N.ptr(perm)))
    return perm[:n]


def train_from_feature_map(net: SimpleNeuralNet, feature_map: Dict[str, np.ndarray], files: Iterable[Tuple[str, int]], epochs: int,
                           lr: float, dropout: float, batch_size: int, rng: Optional[np.random.Generator] = None, seed: int = 0) -> float:
    """lib.rs:632-665: files are trained one after the other, each for all its epochs; mean of per-file losses."""
    total, count = 0.0, 0
    for path, cls in files:
        wins = feature_map.get(path)
        if wins is None:
            continue
        total += pretrain_from_features(net, wins, cls, net.output_size(), epochs, lr, dropout, batch_size, rng=rng, seed=seed + count)
        net.record_training_file(cls, path)
        count += 1
    return total / count if count else 0.0


def identify_counts(net: SimpleNeuralNet, windows, threshold: float) -> np.ndarray:
    w = _f32(windows).reshape(-1, net.dims[0])
    counts = np.zeros(net.output_size(), dtype=np.uint64)
    N.check(N.lib.szb_identify_counts(net._h, N.ptr(w), w.shape[0], float(threshold), N.ptr(counts)))
    return counts.astype(np.int64)


def identify_counts_batch(net: SimpleNeuralNet, windows_per_clip: Sequence[np.ndarray], threshold: float) -> np.ndarray:
    """The histograms of identify_speaker_list (lib.rs:1389-1402) for many clips in one call: ``counts[c]`` equals
    ``identify_counts(net, windows_per_clip[c], threshold)``.  (szb_identify_counts_batch_dev: large forward batches over the
    concatenated windows with a per-window clip lookup.)"""
    n_in = net.dims[0]
    wins = [_f32(w).reshape(-1, n_in) for w in windows_per_clip]
    off = np.concatenate([[0], np.cumsum([len(w) for w in wins])]).astype(np.uint64)
    counts = np.zeros((len(wins), net.output_size()), dtype=np.uint32)
    if not wins:
        return counts
    flat = _f32(np.concatenate(wins)) if int(off[-1]) else np.zeros((0, n_in), np.float32)
    d = net.ctx.dev_alloc(max(1, flat.nbytes))
    try:
        if flat.nbytes:
            net.ctx.h2d(d, flat)
        N.check(N.lib.szb_identify_counts_batch_dev(net._h, C.c_void_p(d), N.ptr(off), len(wins), float(threshold), N.ptr(counts)))
    finally:
        net.ctx.dev_free(d)
    return counts


def identify_sums(net: SimpleNeuralNet, windows) -> np.ndarray:
    w = _f32(windows).reshape(-1, net.dims[0])
    sums = np.zeros(net.output_size(), dtype=np.float32)
    N.check(N.lib.szb_identify_sums(net._h, N.ptr(w), w.shape[0], N.ptr(sums)))
    return sums


def identify_counts_sharded(net: SimpleNeuralNet, sample, threshold: float, rank: int, world: int,
                            extractor: Optional[FeatureExtractor] = None) -> np.ndarray:
    """This rank's share of the histogram of identify_speaker_list (lib.rs:1389-1400) for one long clip: the clip's windows
    are cut into `world` contiguous ranges (``sharding.shard_windows``), rank r extracts and classifies range r only, and
    the per-class counts of all ranks ADD UP to the single-GPU histogram -- summed on the host, no collective (C <= 1000
    integers).  ``speakers_from_counts(sum)`` then gives the list."""
    from .sharding import shard_windows
    extractor = extractor or FeatureExtractor(net.ctx)
    s = _i16(sample)
    w0, w1 = shard_windows(num_windows(len(s)), world)[rank]
    if w1 <= w0:
        return np.zeros(net.output_size(), dtype=np.uint64)
    return identify_counts(net, extractor.extract_range(s, w0, w1), threshold)


def speakers_from_counts(counts) -> List[int]:
    """lib.rs:1402-1410: speakers with a non-zero count, by count descending, ties by ascending index (stable sort)."""
    counts = np.asarray(counts)
    idx = [i for i in range(len(counts)) if counts[i] > 0]
    return sorted(idx, key=lambda i: -int(counts[i]))


def _argmax_last(v: np.ndarray) -> int:
    return int(len(v) - 1 - np.argmax(v[::-1]))  # max_by keeps the last maximal element (lib.rs:1298-1301)


def identify_speaker(net: SimpleNeuralNet, sample, extractor: FeatureExtractor) -> int:
    """lib.rs:1285-1303."""
    sums = identify_sums(net, extractor.extract(sample))
    return _argmax_last(sums) if len(sums) else 0


def identify_speaker_with_threshold_feats(net: SimpleNeuralNet, windows, threshold: float) -> Optional[int]:
    """lib.rs:1346-1377."""
    w = _f32(windows).reshape(-1, net.dims[0])
    if net.output_size() <= 1 or w.shape[0] == 0:
        return None
    sums = identify_sums(net, w)
    best = _argmax_last(sums)
    return best if np.float32(sums[best]) / np.float32(w.shape[0]) >= np.float32(threshold) else None


def identify_speaker_with_threshold(net: SimpleNeuralNet, sample, threshold: float, extractor: FeatureExtractor) -> Optional[int]:
    """lib.rs:1307-1343."""
    if net.output_size() <= 1:
        return None
    return identify_speaker_with_threshold_feats(net, extractor.extract(sample), threshold)


def identify_speaker_list(net: SimpleNeuralNet, sample, threshold: float, extractor: Optional[FeatureExtractor] = None) -> List[int]:
    """lib.rs:1383-1411: extraction, forward, argmax/threshold histogram and the stable sort run behind one C call."""
    s = _i16(sample)
    cap = net.output_size()
    out = np.zeros(max(1, cap), dtype=np.uint32)
    n = C.c_uint32()
    N.check(N.lib.szb_identify_speaker_list(net._h, N.ptr(s), len(s), float(threshold), N.ptr(out), cap, C.byref(n)))
    return [int(x) for x in out[: n.value]]


# ---- embeddings and cosine matching ----------------------------------------------------------------

def embedding_size(net: SimpleNeuralNet) -> int:
    n = C.c_uint32()
    N.check(N.lib.szb_net_embedding_size(net._h, C.byref(n)))
    return int(n.value)


def embed(net: SimpleNeuralNet, x, relu2: bool = False) -> np.ndarray:
    """``embed`` (ReLU, tanh; lib.rs:895-900) or, with ``relu2``, ``forward_embedding`` (ReLU, ReLU; lib.rs:1073-1079)."""
    x = _f32(x)
    single = x.ndim == 1
    x2 = x.reshape(1, -1) if single else x
    out = np.empty((x2.shape[0], embedding_size(net)), dtype=np.float32)
    N.check(N.lib.szb_net_embed(net._h, N.ptr(x2), x2.shape[0], 1 if relu2 else 0, N.ptr(out)))
    return out[0] if single else out


def forward_embedding(net: SimpleNeuralNet, x) -> np.ndarray:
    return embed(net, x, relu2=True)


def extract_embedding_from_features(net: SimpleNeuralNet, feats) -> np.ndarray:
    """lib.rs:1453-1475."""
    w = _f32(feats).reshape(-1, net.dims[0])
    out = np.empty(embedding_size(net), dtype=np.float32)
    N.check(N.lib.szb_net_embedding_mean(net._h, N.ptr(w), w.shape[0], N.ptr(out)))
    return out


def median_embedding_from_features(net: SimpleNeuralNet, feats, relu2: bool = True) -> np.ndarray:
    """lib.rs:1478-1500."""
    w = _f32(feats).reshape(-1, net.dims[0])
    out = np.empty(embedding_size(net), dtype=np.float32)
    N.check(N.lib.szb_net_embedding_median(net._h, N.ptr(w), w.shape[0], 1 if relu2 else 0, N.ptr(out)))
    return out


def extract_embedding(net: SimpleNeuralNet, sample, extractor: FeatureExtractor) -> np.ndarray:
    """lib.rs:1418-1450: per-dimension median of ``embed`` over the clip's windows, normalised."""
    return median_embedding_from_features(net, extractor.extract(sample), relu2=False)


def cosine_similarity(a, b) -> float:
    a, b = _f32(a), _f32(b)
    return float(N.lib.szb_cosine_similarity(N.ptr(a), N.ptr(b), len(a)))


def identify_speaker_from_embedding(emb, speaker_embeddings: Dict[int, np.ndarray], threshold: float) -> Optional[int]:
    """lib.rs:1503-1529 (``None`` stands for ``usize::MAX``); the rule itself lives behind the C ABI (szb_match_embedding)."""
    if not speaker_embeddings:
        return None
    ids = np.array(list(speaker_embeddings.keys()), dtype=np.uint64)
    cents = _f32(np.stack([np.asarray(v, np.float32) for v in speaker_embeddings.values()]))
    e = _f32(emb)
    best, sim = C.c_uint64(), C.c_float()
    N.check(N.lib.szb_match_embedding(N.ptr(e), N.ptr(cents), N.ptr(ids), len(ids), cents.shape[1], float(threshold), C.byref(best),
                                      C.byref(sim)))
    return None if best.value == 2 ** 64 - 1 else int(best.value)


def _cosine_decision(emb, speaker_embeds, threshold: float) -> Optional[int]:
    best_idx, best_val = None, threshold
    for i, (mean, mean_sim, std_sim) in enumerate(speaker_embeds):
        sim = cosine_similarity(emb, mean)
        if sim < mean_sim - 2.0 * std_sim:
            continue
        factor = 0.3 if len(speaker_embeds) < 200 else 1.0
        if sim > 0.35 and (sim > mean_sim + std_sim * factor or sim > 0.5) and sim > best_val:
            best_val, best_idx = sim, i
    return best_idx


def identify_speaker_cosine_feats(net: SimpleNeuralNet, speaker_embeds, windows, threshold: float) -> Optional[int]:
    """lib.rs:1635-1661."""
    if not speaker_embeds:
        return None
    return _cosine_decision(extract_embedding_from_features(net, windows), speaker_embeds, threshold)


def identify_speaker_cosine(net: SimpleNeuralNet, speaker_embeds, sample, threshold: float, extractor: FeatureExtractor) -> Optional[int]:
    """lib.rs:1604-1632."""
    if not speaker_embeds:
        return None
    return _cosine_decision(extract_embedding(net, sample, extractor), speaker_embeds, threshold)


def compute_speaker_embeddings(net: SimpleNeuralNet, features_by_path: Dict[str, np.ndarray]):
    """lib.rs:1555-1599: per speaker (mean of the files' median embeddings, mean cosine to it, std of the cosines)."""
    out = []
    h2 = embedding_size(net)
    for files in net.file_lists():
        embeds = [median_embedding_from_features(net, features_by_path[p]) for p in files if p in features_by_path]
        if not embeds:
            out.append((np.zeros(h2, np.float32), 0.0, 0.0))
            continue
        mean = np.mean(np.stack(embeds), axis=0, dtype=np.float32)
        norm = np.sqrt((mean * mean).sum())
        mean = mean / norm if norm > 1e-6 else mean
        sims = np.array([cosine_similarity(e, mean) for e in embeds], np.float32)
        out.append((mean.astype(np.float32), float(sims.mean()), float(np.sqrt(((sims - sims.mean()) ** 2).mean()))))
    return out


def feature_cache_path(path: str) -> str:
    """lib.rs:550-555 (does not create the directory)."""
    buf = C.create_string_buffer(len(path) + 64)
    N.check(N.lib.szb_feature_cache_path(path.encode(), buf, len(buf)))
    return buf.value.decode()


def write_npy(path: str, arr):
    a = _f32(arr)
    assert a.ndim == 2
    N.check(N.lib.szb_npy_write_f32(path.encode(), N.ptr(a), a.shape[0], a.shape[1]))


def read_npy(path: str) -> np.ndarray:
    r, c = C.c_uint64(), C.c_uint64()
    N.check(N.lib.szb_npy_read_f32(path.encode(), None, 0, C.byref(r), C.byref(c)))
    out = np.empty((r.value, c.value), dtype=np.float32)
    N.check(N.lib.szb_npy_read_f32(path.encode(), N.ptr(out), out.size, C.byref(r), C.byref(c)))
    return out


def load_cached_features(path: str, samples_loader, extractor: FeatureExtractor) -> np.ndarray:
    """lib.rs:558-579: read ``feature_cache/<sanitised>.npy`` if present, else extract and (when non-empty) store it.
    Audio decoding stays on the host and is out of scope, so the caller supplies ``samples_loader(path) -> i16``."""
    cache = feature_cache_path(path)
    if os.path.exists(cache):
        return read_npy(cache)
    feats = extractor.extract(samples_loader(path))
    if len(feats):
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        try:
            write_npy(cache, feats)
        except N.StreamzError:
            pass  # `let _ = write_npy(..)`, lib.rs:576
    return feats
