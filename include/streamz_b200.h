/*
 * streamz_b200 -- C ABI of the B200-native (sm_100a) implementation of StreamZ's data-parallel hot path.
 *
 * This is the drop-in boundary.  The reference (Mycoearthdome/StreamZ, crate `streamz_rs`) has no FFI of its own:
 * its boundary is the public Rust API imported by the CLI at streamz-rs/src/main.rs:13-19.  Every entry point below
 * names the Rust item (file:line in streamz-rs/src/lib.rs) it replaces; INTEGRATION.md shows the `extern "C"` block
 * and the wrappers a maintainer adds to keep those Rust signatures.
 *
 * Conventions
 *   - plain pointers and sizes only; no CUDA, torch or C++ types in any signature;
 *   - every function returns szb_status (0 = ok) unless it is a pure size query; nothing throws or aborts across
 *     the ABI; szb_last_error() returns a human-readable message for the calling thread's last failure;
 *   - "host" entry points take host pointers and do their own H2D/D2H copies on the context's stream;
 *     "_dev" entry points take device pointers that live on the context's device (allocate with szb_dev_alloc);
 *   - outputs are caller-allocated; capacities are passed explicitly and SZB_ERR_INVALID is returned when too small;
 *   - a context owns one CUDA stream and scratch buffers: use it from one thread at a time; a net belongs to a
 *     context; read-only net calls on different contexts are independent;
 *   - there is NO CPU fallback: without a usable CUDA device szb_ctx_create fails with SZB_ERR_NO_DEVICE.
 */
#ifndef STREAMZ_B200_H
#define STREAMZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib.rs:25-36 */
#define SZB_SAMPLE_RATE 44100u /* DEFAULT_SAMPLE_RATE */
#define SZB_WINDOW_SIZE 800u   /* WINDOW_SIZE */
#define SZB_HOP_SIZE 400u      /* WINDOW_SIZE / 2, lib.rs:288 */
#define SZB_N_MELS 26u         /* N_MELS */
#define SZB_MFCC_SIZE 20u      /* MFCC_SIZE */
#define SZB_FEATURE_SIZE 60u   /* FEATURE_SIZE (WITH_DELTAS) */
#define SZB_DEFAULT_DROPOUT 0.2f
#define SZB_RESAMPLE_TAPS 16u

typedef int32_t szb_status;
enum {
    SZB_OK = 0,
    SZB_ERR_INVALID = 1,     /* bad argument / capacity too small / shape mismatch */
    SZB_ERR_CUDA = 2,        /* a CUDA runtime call failed */
    SZB_ERR_NO_DEVICE = 3,   /* no usable sm_100 device: there is no CPU fallback */
    SZB_ERR_ALLOC = 4,
    SZB_ERR_NCCL = 5,
    SZB_ERR_IO = 6,          /* file missing / malformed npy or npz */
    SZB_ERR_UNSUPPORTED = 7
};

typedef struct szb_ctx szb_ctx;
typedef struct szb_net szb_net;

/* ---- library / context ------------------------------------------------------------------------------------------ */
const char* szb_version(void);
const char* szb_last_error(void);
/* One context per (thread, device).  `stream` may be NULL (the context creates its own non-blocking stream) or an
 * existing cudaStream_t passed as void* (the context then launches on it and does not destroy it). */
szb_status szb_ctx_create(int32_t device, void* stream, szb_ctx** out);
void szb_ctx_destroy(szb_ctx* ctx);
szb_status szb_ctx_sync(szb_ctx* ctx);
int32_t szb_ctx_sm_count(const szb_ctx* ctx);
/* szb_extract_batch(_dev) with rate != 44100: enable = 1 runs the polyphase FIR inside the extraction kernel's staging -- every
 * 33-hop tile of the 44.1 kHz signal is produced in shared memory from the original-rate clip and never written to HBM; enable = 0
 * (the DEFAULT) runs resample_kernel -> extract_kernel with the 44.1 kHz i16 intermediate in HBM.  Identical results (bit for
 * bit, tested); the fused kernel moves 4x less DRAM traffic but measured 45 % slower on B200, hence the default.
 * Falls back to the two-kernel path for rates above 44.1 kHz or clips that do not start on 16-byte boundaries. */
szb_status szb_ctx_set_fused_resample(szb_ctx* ctx, int32_t enable);
/* OPTIONAL mode of the device-resident szb_extract_batch_dev with rate != 44100: the batch is processed in chunks whose
 * 44.1 kHz intermediate (chunk_mb MB) is written by the resampler into a two-slot ring that stays in the 126 MB L2 and is
 * read back by the extraction kernel from there, so it never makes the round trip through HBM.  streams = 2 runs the
 * resampler of chunk k + 1 on a second stream under the tail of extraction k.  chunk_mb = 0 (the DEFAULT): one chunk,
 * intermediate in HBM.  Results are identical in every setting; every chunked setting measured slower than the default on
 * B200, which is why it is off. */
szb_status szb_ctx_set_l2_ring(szb_ctx* ctx, int32_t chunk_mb, int32_t streams);
/* Number of kernels this context has launched since creation (bench.py reports it as gpu_launches). */
uint64_t szb_ctx_launch_count(const szb_ctx* ctx);
/* Of those, how many two-step training graphs were replayed with cudaGraphLaunch (small-batch epochs; each replay stands for
 * 22 kernels in the count above). */
uint64_t szb_ctx_graph_launch_count(const szb_ctx* ctx);
/* Device-time of the work enqueued between start and stop on the context's stream (CUDA events). */
szb_status szb_timer_start(szb_ctx* ctx);
szb_status szb_timer_stop(szb_ctx* ctx, float* elapsed_ms);
/* Accumulated device time (ms) and launches of the dominant extraction kernel since the last reset, measured with
 * CUDA events around each launch when enabled (used for the roofline figure; off by default). */
szb_status szb_kernel_timing(szb_ctx* ctx, int32_t enable);
szb_status szb_kernel_timing_read(szb_ctx* ctx, double* total_ms, uint64_t* launches, int32_t reset);

/* device memory helpers for FFI hosts that have no CUDA binding of their own */
szb_status szb_dev_alloc(szb_ctx* ctx, size_t bytes, void** dptr);
szb_status szb_dev_free(szb_ctx* ctx, void* dptr);
szb_status szb_memcpy_h2d(szb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
szb_status szb_memcpy_d2h(szb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
szb_status szb_host_alloc_pinned(size_t bytes, void** hptr);
szb_status szb_host_free_pinned(void* hptr);

/* ---- tables (FeatureExtractor::new, lib.rs:239-257) --------------------------------------------------------------- */
/* 26 x 401 mel bank as the reference holds it (f32), 20 x 26 DCT-II rows, polyphase taps [L][16] for `rate`. */
szb_status szb_table_mel(float* out_26x401);
szb_status szb_table_dct(float* out_20x26);
szb_status szb_table_resample_taps(uint32_t rate, float* out /* may be NULL */, uint32_t* L, uint32_t* M);

/* ---- front end ---------------------------------------------------------------------------------------------------- */
/* n = floor((len - 800) / 400) + 1, or 0 when len < 800 (lib.rs:289-291, 317). */
uint64_t szb_num_windows(uint64_t n_samples);
/* floor(n_in * 44100 / rate) (lib.rs:196). */
uint64_t szb_resample_out_len(uint64_t n_in, uint32_t rate);

/* downmix_to_mono (lib.rs:172-183): interleaved i16 -> mono, i32 sum / channels truncating toward zero. */
szb_status szb_downmix_to_mono(szb_ctx* ctx, const int16_t* interleaved, uint64_t n_samples, uint32_t channels,
                               int16_t* mono, uint64_t mono_cap, uint64_t* n_mono);

/* augment (lib.rs:103-116): circular shift < min(len, 800), gain U(0.95, 1.05), per-sample noise
 * U(-l, l) * 32767 with l ~ U(0, 0.005), clamp, truncating cast.  The reference draws from an unseeded thread_rng; here
 * every draw derives from `seed` (szb_augment_params returns the three clip-level draws). out must not alias in. */
szb_status szb_augment_params(uint64_t seed, uint64_t n_samples, float* noise_level, float* gain, uint64_t* shift);
szb_status szb_augment(szb_ctx* ctx, const int16_t* in, uint64_t n_samples, uint64_t seed, int16_t* out);
szb_status szb_augment_dev(szb_ctx* ctx, const int16_t* d_in, uint64_t n_samples, uint64_t seed, int16_t* d_out);

/* resample_to_44100 (lib.rs:186-209).  rate == 44100 copies.  Otherwise polyphase FIR (DESIGN.md "Resampler"),
 * output clamped to [-32768, 32767] and truncated toward zero like lib.rs:205-208. */
szb_status szb_resample_to_44100(szb_ctx* ctx, const int16_t* in, uint64_t n_in, uint32_t rate, int16_t* out,
                                 uint64_t out_cap, uint64_t* n_out);

/* FeatureExtractor::extract (lib.rs:261-263 -> 279-345): mono i16 @ 44.1 kHz -> [n][60] normalised MFCC+d+dd.
 * len < 800 gives n = 0 and SZB_OK (lib.rs:289). */
szb_status szb_extract(szb_ctx* ctx, const int16_t* pcm, uint64_t n_samples, float* feats, uint64_t cap_windows,
                       uint64_t* n_windows);

/* Windows [w_begin, w_end) of the clip `pcm` (44.1 kHz): rows identical, bit for bit, to rows w_begin .. w_end - 1 of
 * szb_extract on the whole clip (the delta stencil reads two more frames on either side of the range and clamps at the
 * real clip edges only).  Only the samples those frames cover are copied to the device.  This is the unit of the
 * multi-GPU identification sweep (BASELINE configs[4]): every rank takes a window range of the long clip
 * and the per-class counts are added on the host.  feats: [w_end - w_begin][60]. */
szb_status szb_extract_range(szb_ctx* ctx, const int16_t* pcm, uint64_t n_samples, uint64_t w_begin, uint64_t w_end, float* feats,
                             uint64_t cap_windows);
/* Device-output form: d_feats [w_end - w_begin][60] stays on the GPU (feeds szb_identify_counts_dev / szb_net_forward_dev). */
szb_status szb_extract_range_dev(szb_ctx* ctx, const int16_t* pcm, uint64_t n_samples, uint64_t w_begin, uint64_t w_end,
                                 float* d_feats, uint64_t cap_windows);

/* Batched form of the rayon loop at main.rs:500-508 (+ batch_resample, lib.rs:541-547, when rate != 44100):
 * clip c is pcm[clip_off[c] .. clip_off[c+1]) at `rate` Hz; its windows land at feats[win_off[c] .. win_off[c+1]).
 * win_off has n_clips + 1 entries and is written by the call.  With rate != 44100 the resampler runs fused in front
 * of the framing (the 44.1 kHz i16 signal -- identical to szb_resample_to_44100's -- never leaves the chip). */
szb_status szb_extract_batch(szb_ctx* ctx, const int16_t* pcm, const uint64_t* clip_off, uint32_t n_clips,
                             uint32_t rate, float* feats, uint64_t cap_windows, uint64_t* win_off);
/* Same with pcm and feats resident on the device; clip_off / win_off stay host arrays. */
szb_status szb_extract_batch_dev(szb_ctx* ctx, const int16_t* d_pcm, const uint64_t* clip_off, uint32_t n_clips,
                                 uint32_t rate, float* d_feats, uint64_t cap_windows, uint64_t* win_off);
/* Total windows the batch will produce (size query for the two calls above). */
uint64_t szb_extract_batch_windows(const uint64_t* clip_off, uint32_t n_clips, uint32_t rate);

/* ---- SimpleNeuralNet (lib.rs:745-1060) ----------------------------------------------------------------------------- */
/* SimpleNeuralNet::new (lib.rs:767-790): weights U(-0.5, 0.5), zero biases.  The reference draws from an unseeded
 * thread_rng; here the stream is a documented counter RNG keyed by `seed`. */
szb_status szb_net_create(szb_ctx* ctx, uint32_t n_in, uint32_t h1, uint32_t h2, uint32_t n_out, uint64_t seed,
                          szb_net** out);
/* Row-major w1[n_in][h1] b1[h1] w2[h1][h2] b2[h2] w3[h2][n_out] b3[n_out] (lib.rs:746-751). */
szb_status szb_net_from_weights(szb_ctx* ctx, uint32_t n_in, uint32_t h1, uint32_t h2, uint32_t n_out,
                                const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                                const float* b3, szb_net** out);
szb_status szb_net_get_weights(szb_net* net, float* w1, float* b1, float* w2, float* b2, float* w3, float* b3);
szb_status szb_net_dims(const szb_net* net, uint32_t dims[4]);
/* output_size (lib.rs:792) */
uint32_t szb_net_output_size(const szb_net* net);
/* add_output_class (lib.rs:797-821): w3 gains a column (given, or U(-0.5,0.5) from `seed` when NULL), b3 a zero. */
szb_status szb_net_add_output_class(szb_net* net, const float* new_col, uint64_t seed);
void szb_net_destroy(szb_net* net);
/* Arithmetic of the dense layers.  1 (default): tensor cores, tcgen05 kind::tf32 with the 3xTF32 split
 * (x = hi + lo; A_lo B_hi + A_hi B_lo + A_hi B_hi), FP32-equivalent for parity with the reference's FP32 `dot`.
 * 2: plain TF32 tensor cores (10-bit mantissa inputs, FP32 accumulate).  0: FP32 CUDA-core GEMM. */
szb_status szb_net_set_precision(szb_net* net, int32_t mode);
int32_t szb_net_get_precision(const szb_net* net);
/* record_training_file / file_lists (lib.rs:855-867): host-side bookkeeping saved into model.npz as
 * speaker_<i>_files.  szb_net_file_list returns the newline-joined list of `speaker` (len excludes the NUL). */
szb_status szb_net_record_training_file(szb_net* net, uint32_t speaker, const char* path);
szb_status szb_net_file_list(const szb_net* net, uint32_t speaker, char* out, size_t cap, size_t* len);

/* forward (lib.rs:880-891), batched: x [B][n_in] -> probs [B][n_out]. */
szb_status szb_net_forward(szb_net* net, const float* x, uint64_t B, float* probs);
szb_status szb_net_forward_dev(szb_net* net, const float* d_x, uint64_t B, float* d_probs);

/* train_batch (lib.rs:1002-1060): one mean-gradient SGD step; `target` is ONE [n_out] vector shared by the batch,
 * exactly the reference's signature.  B == 0 is a no-op. */
szb_status szb_net_train_batch(szb_net* net, const float* x, uint64_t B, const float* target, float lr);
/* Superset: per-window labels (label >= n_out -> all-zero target, lib.rs:592-595), optional dropout decisions
 * keep[B][n_in] (0 = zero the input, lib.rs:119-129), windows left all-zero are skipped and excluded from the
 * divisor (lib.rs:607-609, 1047).  loss_sum / n_used (may be NULL) return the summed -ln(max(p[label],1e-12)) computed
 * with the pre-update weights (lib.rs:610-617) and the number of windows used. */
szb_status szb_net_train_batch_labels(szb_net* net, const float* x, const uint32_t* labels, uint64_t B, float lr,
                                      const uint8_t* keep, double* loss_sum, uint64_t* n_used);
/* One epoch of pretrain_from_features' loop (lib.rs:599-622) with the features resident on the device: rows are
 * visited in the order perm[0..n_perm) (host array; the shuffle of lib.rs:601), in chunks of `batch`.  Dropout
 * decisions come from d_keep [n][n_in] when non-NULL, else from the counter RNG (seed, stream) when dropout > 0. */
szb_status szb_net_train_epoch_dev(szb_net* net, const float* d_feats, const uint32_t* d_labels, uint64_t n,
                                   const uint32_t* perm, uint64_t n_perm, uint32_t batch, float lr, float dropout,
                                   uint64_t seed, uint64_t stream, const uint8_t* d_keep, double* loss_sum,
                                   uint64_t* n_used);
/* Same with explicit per-step row counts (step i trains on the next step_sizes[i] rows of perm; 0 is allowed).  This is
 * the multi-GPU form: each rank passes its slice of every global batch (streamz_b200.sharding.shard_batches) and the
 * step all-reduces [gradient | window count | loss] over the communicator of szb_comm_init, so all ranks must pass the
 * same n_steps.  loss_sum / n_used are then global. */
szb_status szb_net_train_epoch_steps_dev(szb_net* net, const float* d_feats, const uint32_t* d_labels, uint64_t n,
                                         const uint32_t* perm, uint64_t n_perm, const uint32_t* step_sizes, uint32_t n_steps,
                                         float lr, float dropout, uint64_t seed, uint64_t stream, const uint8_t* d_keep,
                                         double* loss_sum, uint64_t* n_used);
/* The dropout stream above, on the host, for callers that need the decisions (tests, oracle): keep[n_rows][n_in]. */
szb_status szb_dropout_keep_mask(uint64_t seed, uint64_t stream, const uint64_t* rows, uint64_t n_rows, uint32_t n_in,
                                 float prob, uint8_t* keep);

/* ---- raw-audio training loops (lib.rs:348-397, 668-732) ------------------------------------------ */
/* Seed of (file, epoch) inside the two loops below; the shuffled visiting order of n rows for (seed, stream) -- Fisher-Yates
 * from the back, j = splitmix64(key ^ i) % (i + 1), the library's stand-in for `windows.shuffle(&mut thread_rng)`
 * (lib.rs:370, 601); lr * 0.99f32.powi(step) with powi evaluated by squaring in float32 (lib.rs:709). */
uint64_t szb_loop_seed(uint64_t seed, uint32_t file, uint32_t epoch);
szb_status szb_shuffle_perm(uint64_t seed, uint64_t stream, uint64_t n, uint32_t* perm);
float szb_lr_decay(float lr, int32_t step);
/* pretrain_network (lib.rs:348-397): `epochs` times  augment (lib.rs:368) -> extract (369) -> shuffle (370) -> chunks of
 * `batch` with input dropout, all-zero windows skipped, loss with the pre-update weights, train_batch (371-390).  pcm is a
 * HOST clip at 44.1 kHz; everything after its upload stays on the device.  Epoch e uses szb_loop_seed(seed, 0, e) for the
 * augmentation draws, the shuffle and the dropout stream.  Returns the summed loss and the number of windows used
 * (the reference returns their ratio, 0.0 when no window was used, lib.rs:392-396). */
szb_status szb_net_pretrain_network(szb_net* net, const int16_t* pcm, uint64_t n_samples, uint32_t target_class, uint32_t epochs,
                                    float lr, float dropout, uint32_t batch, uint64_t seed, double* loss_sum, uint64_t* n_used);
/* train_from_files (lib.rs:668-732) on decoded, resampled clips (decode stays on the host, lib.rs:696): file f is
 * pcm[clip_off[f] .. clip_off[f+1]) with class classes[f]; each file runs `epochs` single-epoch pretrain_network calls with
 * lr * 0.99^step, step counting every (file, epoch) in order (lib.rs:708-709).  The reference interleaves files
 * nondeterministically (rayon + write lock); this is the file-major serialisation.  The caller records the training files
 * (szb_net_record_training_file, lib.rs:723) and the dataset specs (lib.rs:703-706). */
szb_status szb_net_train_from_files(szb_net* net, const int16_t* pcm, const uint64_t* clip_off, const uint32_t* classes,
                                    uint32_t n_files, uint32_t epochs, float lr, float dropout, uint32_t batch, uint64_t seed,
                                    double* loss_sum, uint64_t* n_used);

/* ---- aggregation (lib.rs:1285-1411) -------------------------------------------------------------------------------- */
/* Per-class count of windows whose LAST-index argmax probability is >= threshold (lib.rs:1391-1402). */
szb_status szb_identify_counts(szb_net* net, const float* feats, uint64_t n_windows, float threshold,
                               uint64_t* counts /* [n_out] */);
szb_status szb_identify_counts_dev(szb_net* net, const float* d_feats, uint64_t n_windows, float threshold,
                                   uint64_t* counts /* host [n_out] */);
/* Batched form for many clips (the identification loop of a whole shard, BASELINE configs[3]): windows of clip c are rows
 * [win_off[c], win_off[c+1]) of d_feats -- the layout szb_extract_batch_dev leaves -- and counts[c][k] is the histogram of
 * lib.rs:1389-1402 for clip c.  One pass of large forward batches with a per-window clip lookup instead of one launch
 * sequence and one synchronising read-back per clip.  counts: HOST [n_clips][n_out] u32; win_off: host, n_clips + 1. */
szb_status szb_identify_counts_batch_dev(szb_net* net, const float* d_feats, const uint64_t* win_off, uint32_t n_clips,
                                         float threshold, uint32_t* counts);
/* Per-class sum of window probabilities (lib.rs:1290-1297, 1319-1328). */
szb_status szb_identify_sums(szb_net* net, const float* feats, uint64_t n_windows, float* sums /* [n_out] */);
/* identify_speaker_list (lib.rs:1383-1411): extraction + forward + histogram + stable sort by count descending. */
szb_status szb_identify_speaker_list(szb_net* net, const int16_t* pcm, uint64_t n_samples, float threshold,
                                     uint32_t* speakers, uint32_t cap, uint32_t* n_speakers);

/* ---- embeddings (SURVEY.md 8(f) "next" row N1; lib.rs:893-905, 1073-1079, 1413-1540) -------------------------------- */
/* embedding_size (lib.rs:903) */
szb_status szb_net_embedding_size(const szb_net* net, uint32_t* size);
/* Second hidden layer for B windows, out [B][h2]: relu2 = 0 is `embed` (ReLU, tanh; lib.rs:895-900), relu2 = 1 is
 * `forward_embedding` (ReLU, ReLU; lib.rs:1073-1079). */
szb_status szb_net_embed(szb_net* net, const float* x, uint64_t B, int32_t relu2, float* out);
/* extract_embedding_from_features (lib.rs:1453-1475): mean of forward_embedding over the windows, L2-normalised. */
szb_status szb_net_embedding_mean(szb_net* net, const float* feats, uint64_t n_windows, float* out /* [h2] */);
/* Per-dimension median over the windows, L2-normalised: relu2 = 1 is median_embedding_from_features (lib.rs:1478-1500),
 * relu2 = 0 the reduction of extract_embedding (lib.rs:1418-1450).  No windows: zero vector. */
szb_status szb_net_embedding_median(szb_net* net, const float* feats, uint64_t n_windows, int32_t relu2, float* out);
/* cosine_similarity (lib.rs:1531-1540) */
float szb_cosine_similarity(const float* a, const float* b, uint32_t n);
/* identify_speaker_from_embedding (lib.rs:1503-1529): best cosine match among n centroids [n][dim] (ids may be NULL: 0..n-1);
 * *best_id = UINT64_MAX (usize::MAX) when the best similarity does not exceed the threshold (x 0.7 below 20 speakers). */
szb_status szb_match_embedding(const float* emb, const float* centroids, const uint64_t* ids, uint32_t n, uint32_t dim,
                               float threshold, uint64_t* best_id, float* best_sim);

/* ---- multi-GPU: batch-parallel training, one NCCL all-reduce of the flattened gradient per step --------------------- */
szb_status szb_comm_unique_id(uint8_t id[128]);
szb_status szb_comm_init(szb_ctx* ctx, const uint8_t id[128], int32_t rank, int32_t world);
szb_status szb_comm_destroy(szb_ctx* ctx);
int32_t szb_comm_world(const szb_ctx* ctx);
/* Gradient exchange over NVLink peer memory, fused into the update kernel (collective call: every rank, same `enable`): every
 * rank's exchange region is mapped into all ranks with CUDA IPC and the SGD kernel itself all-reduces the step's gradient with
 * posted peer stores and flag rounds -- no NCCL call inside a step.  enable: 0 = off (NCCL all-reduces),
 * 1 = on, protocol chosen by size; 2 = on, ONE-SHOT (every rank stores its whole vector into every peer, one flag round,
 * every rank adds the vectors in rank order while it updates); 3 = on, TWO-SHOT (scatter slices, owner reduces in rank order
 * and broadcasts, two flag rounds); 4 = on, PACKETS (one-shot, every element travels as an 8-byte {value, step} packet that is
 * its own flag: no flag round, twice the bytes; tensor-core update kernel only, otherwise chosen by size).  Replicas stay
 * bit-identical with all of them.  *active reports whether the mapping succeeded
 * on all ranks (otherwise every rank stays on NCCL). */
szb_status szb_comm_peer_exchange(szb_ctx* ctx, int32_t enable, int32_t* active);

/* Diagnostics of the fused exchange: enable != 0 starts (and clears) the recording of the phase times of the update kernel's
 * CTA 0; ns[0..5] = mean nanoseconds per step in: scatter stores, first publish, wait for flag1, reduce + broadcast, second
 * publish, wait for flag2; ns[7] = steps recorded since the last call. */
szb_status szb_comm_peer_trace(szb_ctx* ctx, int32_t enable, double* ns /* [8], may be NULL */);

/* ---- on-disk formats (host code) ----------------------------------------------------------------------------------- */
/* feature_cache/<sanitised path>.npy (lib.rs:550-579): C-order <f4 [n][60]. */
szb_status szb_feature_cache_path(const char* audio_path, char* out, size_t cap);
szb_status szb_npy_write_f32(const char* path, const float* data, uint64_t rows, uint64_t cols);
szb_status szb_npy_read_f32(const char* path, float* data, uint64_t cap_elems, uint64_t* rows, uint64_t* cols);
/* model.npz (lib.rs:1081-1282): stored (uncompressed) zip of npy members w1 b1 w2 b2 sample_rate bits num_speakers
 * w3_k b3_k (k = 1..C, one column each) speaker_i_files.  Loader also accepts the legacy dense w3 / b3 pair. */
szb_status szb_net_save(szb_net* net, const char* path, uint32_t sample_rate, uint32_t bits);
szb_status szb_net_load(szb_ctx* ctx, const char* path, szb_net** out, uint32_t* sample_rate, uint32_t* bits);
/* set_embeddings / embeddings (lib.rs:869-877): per-speaker (embedding[dim], mean similarity, std similarity) triples that
 * model.npz carries as speaker_embeddings [n][dim], speaker_mean_sims [n], speaker_std_sims [n] (lib.rs:1114-1127,
 * 1253-1264); the CLI sets them right before saving (main.rs:845-856).  get: pass all three outputs NULL for a size query. */
szb_status szb_net_set_embeddings(szb_net* net, const float* emb, const float* mean_sims, const float* std_sims, uint32_t n,
                                  uint32_t dim);
szb_status szb_net_get_embeddings(const szb_net* net, float* emb, float* mean_sims, float* std_sims, uint32_t cap_n, uint32_t* n,
                                  uint32_t* dim);
/* The optional hidden encoding layer w4 [rows][n] (row-major) / b4 [n] (lib.rs:752-754; npz members w4_k / b4_k,
 * lib.rs:1099-1108, 1168-1226).  Out of scope for the hot path: carried through load -> save so a
 * reference-written model keeps it, never evaluated.  n = 0 removes it; get with w4 = b4 = NULL is a size query. */
szb_status szb_net_set_encoding_layer(szb_net* net, const float* w4, const float* b4, uint32_t rows, uint32_t n);
szb_status szb_net_get_encoding_layer(const szb_net* net, float* w4, float* b4, uint64_t cap_elems, uint32_t* rows, uint32_t* n);

#ifdef __cplusplus
}
#endif
#endif /* STREAMZ_B200_H */
