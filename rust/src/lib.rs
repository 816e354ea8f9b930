//! Drop-in host shim: the hot-path subset of `streamz_rs` (streamz-rs/src/lib.rs) on top of libstreamz_b200.
//!
//! NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no cargo/rustc there).  The tested equivalents of this file are
//! `include/streamz_rs.hpp` (C++) and `streamz_b200/api.py` (Python), which wrap the same C entry points in the same way.
//! Signatures below are the reference's; comments give the reference line each one replaces.
#![allow(clippy::too_many_arguments)]
use std::error::Error;
use std::ffi::{c_char, c_void, CStr, CString};
use std::os::raw::c_int;

pub const DEFAULT_SAMPLE_RATE: u32 = 44100; // lib.rs:25
pub const WINDOW_SIZE: usize = 800; // lib.rs:26
pub const MFCC_SIZE: usize = 20; // lib.rs:28
pub const FEATURE_SIZE: usize = 60; // lib.rs:30-34
pub const DEFAULT_DROPOUT: f32 = 0.2; // lib.rs:36

#[repr(C)]
pub struct SzbCtx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct SzbNet {
    _p: [u8; 0],
}

extern "C" {
    fn szb_last_error() -> *const c_char;
    fn szb_ctx_create(device: i32, stream: *mut c_void, out: *mut *mut SzbCtx) -> c_int;
    fn szb_ctx_destroy(ctx: *mut SzbCtx);
    fn szb_num_windows(n_samples: u64) -> u64;
    fn szb_resample_out_len(n_in: u64, rate: u32) -> u64;
    fn szb_downmix_to_mono(ctx: *mut SzbCtx, inp: *const i16, n: u64, ch: u32, out: *mut i16, cap: u64, n_out: *mut u64) -> c_int;
    fn szb_resample_to_44100(ctx: *mut SzbCtx, inp: *const i16, n: u64, rate: u32, out: *mut i16, cap: u64, n_out: *mut u64) -> c_int;
    fn szb_extract(ctx: *mut SzbCtx, pcm: *const i16, n: u64, feats: *mut f32, cap: u64, n_win: *mut u64) -> c_int;
    fn szb_extract_batch(ctx: *mut SzbCtx, pcm: *const i16, clip_off: *const u64, n_clips: u32, rate: u32, feats: *mut f32,
                         cap: u64, win_off: *mut u64) -> c_int;
    fn szb_extract_batch_windows(clip_off: *const u64, n_clips: u32, rate: u32) -> u64;
    fn szb_net_create(ctx: *mut SzbCtx, n_in: u32, h1: u32, h2: u32, n_out: u32, seed: u64, out: *mut *mut SzbNet) -> c_int;
    fn szb_net_destroy(net: *mut SzbNet);
    fn szb_net_dims(net: *const SzbNet, dims: *mut u32) -> c_int;
    fn szb_net_output_size(net: *const SzbNet) -> u32;
    fn szb_net_add_output_class(net: *mut SzbNet, new_col: *const f32, seed: u64) -> c_int;
    fn szb_net_record_training_file(net: *mut SzbNet, speaker: u32, path: *const c_char) -> c_int;
    fn szb_net_forward(net: *mut SzbNet, x: *const f32, b: u64, probs: *mut f32) -> c_int;
    fn szb_net_train_batch(net: *mut SzbNet, x: *const f32, b: u64, target: *const f32, lr: f32) -> c_int;
    fn szb_net_train_epoch_dev(net: *mut SzbNet, d_feats: *const f32, d_labels: *const u32, n: u64, perm: *const u32, n_perm: u64,
                               batch: u32, lr: f32, dropout: f32, seed: u64, stream: u64, d_keep: *const u8, loss: *mut f64,
                               used: *mut u64) -> c_int;
    fn szb_dev_alloc(ctx: *mut SzbCtx, bytes: usize, p: *mut *mut c_void) -> c_int;
    fn szb_dev_free(ctx: *mut SzbCtx, p: *mut c_void) -> c_int;
    fn szb_memcpy_h2d(ctx: *mut SzbCtx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    fn szb_identify_sums(net: *mut SzbNet, feats: *const f32, n: u64, sums: *mut f32) -> c_int;
    fn szb_identify_counts(net: *mut SzbNet, feats: *const f32, n: u64, thr: f32, counts: *mut u64) -> c_int;
    fn szb_extract_range(ctx: *mut SzbCtx, pcm: *const i16, n: u64, w_begin: u64, w_end: u64, feats: *mut f32, cap: u64) -> c_int;
    fn szb_augment(ctx: *mut SzbCtx, inp: *const i16, n: u64, seed: u64, out: *mut i16) -> c_int;
    fn szb_net_embedding_size(net: *const SzbNet, size: *mut u32) -> c_int;
    fn szb_net_embed(net: *mut SzbNet, x: *const f32, b: u64, relu2: i32, out: *mut f32) -> c_int;
    fn szb_net_embedding_mean(net: *mut SzbNet, feats: *const f32, n: u64, out: *mut f32) -> c_int;
    fn szb_net_embedding_median(net: *mut SzbNet, feats: *const f32, n: u64, relu2: i32, out: *mut f32) -> c_int;
    fn szb_cosine_similarity(a: *const f32, b: *const f32, n: u32) -> f32;
    fn szb_identify_speaker_list(net: *mut SzbNet, pcm: *const i16, n: u64, thr: f32, out: *mut u32, cap: u32, n_out: *mut u32) -> c_int;
    fn szb_net_save(net: *mut SzbNet, path: *const c_char, sample_rate: u32, bits: u32) -> c_int;
    fn szb_net_load(ctx: *mut SzbCtx, path: *const c_char, out: *mut *mut SzbNet, sr: *mut u32, bits: *mut u32) -> c_int;
    fn szb_match_embedding(emb: *const f32, centroids: *const f32, ids: *const u64, n: u32, dim: u32, threshold: f32, best_id: *mut u64,
                           best_sim: *mut f32) -> c_int;
    fn szb_net_set_embeddings(net: *mut SzbNet, emb: *const f32, mean: *const f32, std: *const f32, n: u32, dim: u32) -> c_int;
    fn szb_net_get_embeddings(net: *const SzbNet, emb: *mut f32, mean: *mut f32, std: *mut f32, cap_n: u32, n: *mut u32, dim: *mut u32) -> c_int;
    fn szb_net_pretrain_network(net: *mut SzbNet, pcm: *const i16, n: u64, class: u32, epochs: u32, lr: f32, dropout: f32, batch: u32, seed: u64,
                                loss: *mut f64, used: *mut u64) -> c_int;
    fn szb_net_train_from_files(net: *mut SzbNet, pcm: *const i16, clip_off: *const u64, classes: *const u32, n_files: u32, epochs: u32, lr: f32,
                                dropout: f32, batch: u32, seed: u64, loss: *mut f64, used: *mut u64) -> c_int;
}

fn check(status: c_int) -> Result<(), Box<dyn Error>> {
    if status == 0 {
        Ok(())
    } else {
        let msg = unsafe { CStr::from_ptr(szb_last_error()) }.to_string_lossy().into_owned();
        Err(format!("streamz_b200 error {}: {}", status, msg).into())
    }
}

struct Ctx(*mut SzbCtx);
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { szb_ctx_destroy(self.0) }
    }
}
thread_local! {
    // one CUDA context (stream + scratch) per thread, like EXTRACTOR_TLS (lib.rs:266-268)
    static CTX: Ctx = {
        let mut p = std::ptr::null_mut();
        check(unsafe { szb_ctx_create(0, std::ptr::null_mut(), &mut p) }).expect("no usable B200: streamz_b200 has no CPU fallback");
        Ctx(p)
    };
}
fn ctx() -> *mut SzbCtx {
    CTX.with(|c| c.0)
}

/// lib.rs:172-183
pub fn downmix_to_mono(samples: &[i16], channels: usize) -> Vec<i16> {
    let ch = channels.max(1);
    let mut out = vec![0i16; (samples.len() + ch - 1) / ch];
    let mut n = 0u64;
    check(unsafe { szb_downmix_to_mono(ctx(), samples.as_ptr(), samples.len() as u64, channels as u32, out.as_mut_ptr(), out.len() as u64, &mut n) })
        .expect("downmix");
    out.truncate(n as usize);
    out
}

/// lib.rs:186-209
pub fn resample_to_44100(samples: &[i16], from_rate: u32) -> Result<Vec<i16>, Box<dyn Error>> {
    let cap = if from_rate == DEFAULT_SAMPLE_RATE { samples.len() as u64 } else { unsafe { szb_resample_out_len(samples.len() as u64, from_rate) } };
    let mut out = vec![0i16; cap as usize];
    let mut n = 0u64;
    check(unsafe { szb_resample_to_44100(ctx(), samples.as_ptr(), samples.len() as u64, from_rate, out.as_mut_ptr(), cap, &mut n) })?;
    out.truncate(n as usize);
    Ok(out)
}

/// lib.rs:231-264 -- the tables live in the GPU's constant memory
pub struct FeatureExtractor;
impl FeatureExtractor {
    pub fn new() -> Self {
        let _ = ctx();
        FeatureExtractor
    }
    /// lib.rs:261-263
    pub fn extract(&self, samples: &[i16]) -> Vec<Vec<f32>> {
        let n = unsafe { szb_num_windows(samples.len() as u64) } as usize;
        let mut flat = vec![0f32; n * FEATURE_SIZE];
        let mut got = 0u64;
        check(unsafe { szb_extract(ctx(), samples.as_ptr(), samples.len() as u64, flat.as_mut_ptr(), n as u64, &mut got) }).expect("extract");
        flat.chunks(FEATURE_SIZE).map(|c| c.to_vec()).collect()
    }
    /// main.rs:500-508 as ONE call (replaces the rayon loop; `rate != 44100` also replaces batch_resample, lib.rs:541)
    pub fn extract_batch(&self, clips: &[Vec<i16>], rate: u32) -> Vec<Vec<Vec<f32>>> {
        let mut off = vec![0u64; clips.len() + 1];
        let mut pcm = Vec::new();
        for (i, c) in clips.iter().enumerate() {
            pcm.extend_from_slice(c);
            off[i + 1] = pcm.len() as u64;
        }
        let total = unsafe { szb_extract_batch_windows(off.as_ptr(), clips.len() as u32, rate) } as usize;
        let mut flat = vec![0f32; total * FEATURE_SIZE];
        let mut woff = vec![0u64; clips.len() + 1];
        check(unsafe { szb_extract_batch(ctx(), pcm.as_ptr(), off.as_ptr(), clips.len() as u32, rate, flat.as_mut_ptr(), total as u64, woff.as_mut_ptr()) })
            .expect("extract_batch");
        (0..clips.len())
            .map(|i| flat[woff[i] as usize * FEATURE_SIZE..woff[i + 1] as usize * FEATURE_SIZE].chunks(FEATURE_SIZE).map(|c| c.to_vec()).collect())
            .collect()
    }
}

/// lib.rs:271-276
pub fn with_thread_extractor<F, R>(f: F) -> R
where
    F: FnOnce(&FeatureExtractor) -> R,
{
    f(&FeatureExtractor::new())
}

/// lib.rs:745-1282
pub struct SimpleNeuralNet {
    net: *mut SzbNet,
    sample_rate: u32,
    bits: u16,
}
unsafe impl Send for SimpleNeuralNet {}
impl Drop for SimpleNeuralNet {
    fn drop(&mut self) {
        unsafe { szb_net_destroy(self.net) }
    }
}
impl SimpleNeuralNet {
    /// lib.rs:767
    pub fn new(input: usize, hidden1: usize, hidden2: usize, output: usize) -> Self {
        let seed = std::time::SystemTime::now().duration_since(std::time::UNIX_EPOCH).map(|d| d.as_nanos() as u64).unwrap_or(0);
        let mut net = std::ptr::null_mut();
        check(unsafe { szb_net_create(ctx(), input as u32, hidden1 as u32, hidden2 as u32, output as u32, seed, &mut net) }).expect("net");
        Self { net, sample_rate: DEFAULT_SAMPLE_RATE, bits: 16 }
    }
    pub fn output_size(&self) -> usize {
        unsafe { szb_net_output_size(self.net) as usize }
    }
    fn input_size(&self) -> usize {
        let mut d = [0u32; 4];
        unsafe { szb_net_dims(self.net, d.as_mut_ptr()) };
        d[0] as usize
    }
    /// lib.rs:797-821
    pub fn add_output_class(&mut self) {
        check(unsafe { szb_net_add_output_class(self.net, std::ptr::null(), self.output_size() as u64 + 1) }).expect("add_output_class");
    }
    pub fn set_dataset_specs(&mut self, sample_rate: u32, bits: u16) {
        self.sample_rate = sample_rate;
        self.bits = bits;
    }
    /// lib.rs:855-862
    pub fn record_training_file(&mut self, class: usize, path: &str) {
        let c = CString::new(path).unwrap();
        let _ = unsafe { szb_net_record_training_file(self.net, class as u32, c.as_ptr()) };
    }
    /// lib.rs:880-891
    pub fn forward(&self, bits: &[f32]) -> Vec<f32> {
        assert_eq!(bits.len(), self.input_size());
        let mut p = vec![0f32; self.output_size()];
        check(unsafe { szb_net_forward(self.net, bits.as_ptr(), 1, p.as_mut_ptr()) }).expect("forward");
        p
    }
    /// lib.rs:954-999
    pub fn train(&mut self, bits: &[f32], target: &[f32], lr: f32) {
        self.train_batch(&[bits.to_vec()], target, lr)
    }
    /// lib.rs:1002-1060
    pub fn train_batch(&mut self, batch: &[Vec<f32>], target: &[f32], lr: f32) {
        if batch.is_empty() {
            return;
        }
        assert_eq!(target.len(), self.output_size());
        let flat: Vec<f32> = batch.iter().flat_map(|v| v.iter().copied()).collect();
        check(unsafe { szb_net_train_batch(self.net, flat.as_ptr(), batch.len() as u64, target.as_ptr(), lr) }).expect("train_batch");
    }
    /// lib.rs:1081-1130
    pub fn save(&self, path: &str) -> Result<(), Box<dyn Error>> {
        let c = CString::new(path)?;
        check(unsafe { szb_net_save(self.net, c.as_ptr(), self.sample_rate, self.bits as u32) })
    }
    /// lib.rs:869-872: (embedding, mean similarity, std similarity) per speaker; saved with the model (lib.rs:1114-1127)
    pub fn set_embeddings(&mut self, embeds: Vec<(Vec<f32>, f32, f32)>) {
        let dim = embeds.first().map(|e| e.0.len()).unwrap_or(0);
        let flat: Vec<f32> = embeds.iter().flat_map(|e| e.0.iter().copied()).collect();
        let mean: Vec<f32> = embeds.iter().map(|e| e.1).collect();
        let std: Vec<f32> = embeds.iter().map(|e| e.2).collect();
        check(unsafe { szb_net_set_embeddings(self.net, flat.as_ptr(), mean.as_ptr(), std.as_ptr(), embeds.len() as u32, dim as u32) })
            .expect("set_embeddings");
    }
    /// lib.rs:874-877 (returned by value: the arrays live behind the C ABI)
    pub fn embeddings(&self) -> Vec<(Vec<f32>, f32, f32)> {
        let (mut n, mut dim) = (0u32, 0u32);
        check(unsafe { szb_net_get_embeddings(self.net, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), 0, &mut n, &mut dim) })
            .expect("embeddings");
        let (mut flat, mut mean, mut std) = (vec![0f32; (n * dim) as usize], vec![0f32; n as usize], vec![0f32; n as usize]);
        if n > 0 {
            check(unsafe { szb_net_get_embeddings(self.net, flat.as_mut_ptr(), mean.as_mut_ptr(), std.as_mut_ptr(), n, &mut n, &mut dim) })
                .expect("embeddings");
        }
        (0..n as usize).map(|i| (flat[i * dim as usize..(i + 1) * dim as usize].to_vec(), mean[i], std[i])).collect()
    }
    /// lib.rs:1132-1282
    pub fn load(path: &str) -> Result<Self, Box<dyn Error>> {
        let c = CString::new(path)?;
        let (mut net, mut sr, mut bits) = (std::ptr::null_mut(), 0u32, 0u32);
        check(unsafe { szb_net_load(ctx(), c.as_ptr(), &mut net, &mut sr, &mut bits) })?;
        Ok(Self { net, sample_rate: sr, bits: bits as u16 })
    }
}

/// lib.rs:582-628.  Shuffle: Fisher-Yates from a splitmix64 stream seeded by the clock (the reference uses thread_rng).
pub fn pretrain_from_features(net: &mut SimpleNeuralNet, windows: &[Vec<f32>], target_class: usize, num_classes: usize, epochs: usize,
                              lr: f32, dropout: f32, batch_size: usize) -> f32 {
    assert_eq!(num_classes, net.output_size());
    if windows.is_empty() || epochs == 0 {
        return 0.0;
    }
    let n = windows.len();
    let flat: Vec<f32> = windows.iter().flat_map(|v| v.iter().copied()).collect();
    let labels = vec![target_class as u32; n];
    let (mut d_feats, mut d_labels) = (std::ptr::null_mut(), std::ptr::null_mut());
    unsafe {
        check(szb_dev_alloc(ctx(), flat.len() * 4, &mut d_feats)).expect("alloc");
        check(szb_dev_alloc(ctx(), n * 4, &mut d_labels)).expect("alloc");
        check(szb_memcpy_h2d(ctx(), d_feats, flat.as_ptr() as *const c_void, flat.len() * 4)).expect("h2d");
        check(szb_memcpy_h2d(ctx(), d_labels, labels.as_ptr() as *const c_void, n * 4)).expect("h2d");
    }
    let mut state = std::time::SystemTime::now().duration_since(std::time::UNIX_EPOCH).map(|d| d.as_nanos() as u64).unwrap_or(1);
    let mut next = || {
        state = state.wrapping_add(0x9E3779B97F4A7C15);
        let mut z = state;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
        z ^ (z >> 31)
    };
    let seed = next();
    let (mut total, mut count) = (0f64, 0u64);
    let mut perm: Vec<u32> = (0..n as u32).collect();
    for e in 0..epochs {
        for i in (1..n).rev() {
            perm.swap(i, (next() % (i as u64 + 1)) as usize); // lib.rs:601
        }
        let (mut loss, mut used) = (0f64, 0u64);
        check(unsafe {
            szb_net_train_epoch_dev(net.net, d_feats as *const f32, d_labels as *const u32, n as u64, perm.as_ptr(), n as u64,
                                    batch_size.max(1) as u32, lr, dropout, seed, e as u64, std::ptr::null(), &mut loss, &mut used)
        })
        .expect("train_epoch");
        total += loss;
        count += used;
    }
    unsafe {
        szb_dev_free(ctx(), d_feats);
        szb_dev_free(ctx(), d_labels);
    }
    if count > 0 { (total / count as f64) as f32 } else { 0.0 }
}

/// lib.rs:348-397: per epoch augment -> extract -> shuffle -> dropout / train_batch chunks, one C call, device-resident.
/// `_extractor` is kept for the reference's signature; the draws the reference takes from thread_rng derive from the clock.
pub fn pretrain_network(net: &mut SimpleNeuralNet, samples: &[i16], target_class: usize, num_classes: usize, epochs: usize, lr: f32,
                        dropout: f32, batch_size: usize, _extractor: &FeatureExtractor) -> f32 {
    assert_eq!(num_classes, net.output_size());
    let seed = std::time::SystemTime::now().duration_since(std::time::UNIX_EPOCH).map(|d| d.as_nanos() as u64).unwrap_or(0);
    let (mut loss, mut used) = (0f64, 0u64);
    check(unsafe {
        szb_net_pretrain_network(net.net, samples.as_ptr(), samples.len() as u64, target_class as u32, epochs as u32, lr, dropout,
                                 batch_size.max(1) as u32, seed, &mut loss, &mut used)
    })
    .expect("pretrain_network");
    if used > 0 { (loss / used as f64) as f32 } else { 0.0 }
}

/// lib.rs:668-732 on already decoded clips (decode + resample stay with the caller, lib.rs:696): every (file, epoch) is one
/// pretrain_network epoch at lr * 0.99^step; file-major order (one serialisation of the reference's rayon loop under its lock).
pub fn train_from_files(net: std::sync::Arc<std::sync::RwLock<SimpleNeuralNet>>, files: &[(&str, &[i16], usize)], _total_files: usize,
                        num_speakers: usize, epochs: usize, lr: f32, dropout: f32, batch_size: usize, _extractor: &FeatureExtractor)
                        -> Result<(), Box<dyn Error>> {
    let mut guard = net.write().map_err(|_| "poisoned lock")?;
    assert_eq!(num_speakers, guard.output_size());
    guard.sample_rate = DEFAULT_SAMPLE_RATE;                                       // lib.rs:703-706
    guard.bits = 16;
    let mut off = vec![0u64];
    let mut pcm: Vec<i16> = Vec::new();
    for (_, s, _) in files {
        pcm.extend_from_slice(s);
        off.push(pcm.len() as u64);
    }
    let classes: Vec<u32> = files.iter().map(|f| f.2 as u32).collect();
    let seed = std::time::SystemTime::now().duration_since(std::time::UNIX_EPOCH).map(|d| d.as_nanos() as u64).unwrap_or(0);
    let (mut loss, mut used) = (0f64, 0u64);
    check(unsafe {
        szb_net_train_from_files(guard.net, pcm.as_ptr(), off.as_ptr(), classes.as_ptr(), files.len() as u32, epochs as u32, lr, dropout,
                                 batch_size.max(1) as u32, seed, &mut loss, &mut used)
    })?;
    for (path, _, class) in files {
        guard.record_training_file(*class, path);                                   // lib.rs:723
    }
    Ok(())
}

/// lib.rs:632-665
pub fn train_from_feature_map(net: &mut SimpleNeuralNet, feature_map: &std::collections::HashMap<String, Vec<Vec<f32>>>,
                              files: &[(&str, usize)], epochs: usize, lr: f32, dropout: f32, batch_size: usize) -> f32 {
    let (mut total, mut count) = (0f32, 0usize);
    for &(path, class) in files {
        if let Some(wins) = feature_map.get(path) {
            total += pretrain_from_features(net, wins, class, net.output_size(), epochs, lr, dropout, batch_size);
            net.record_training_file(class, path);
            count += 1;
        }
    }
    if count > 0 { total / count as f32 } else { 0.0 }
}

fn prob_sums(net: &SimpleNeuralNet, windows: &[Vec<f32>]) -> Vec<f32> {
    let flat: Vec<f32> = windows.iter().flat_map(|v| v.iter().copied()).collect();
    let mut sums = vec![0f32; net.output_size()];
    check(unsafe { szb_identify_sums(net.net, flat.as_ptr(), windows.len() as u64, sums.as_mut_ptr()) }).expect("identify_sums");
    sums
}
fn argmax_last(v: &[f32]) -> usize {
    v.iter().enumerate().max_by(|a, b| a.1.partial_cmp(b.1).unwrap()).map(|(i, _)| i).unwrap_or(0)
}

/// lib.rs:1285-1303
pub fn identify_speaker(net: &SimpleNeuralNet, sample: &[i16], extractor: &FeatureExtractor) -> usize {
    argmax_last(&prob_sums(net, &extractor.extract(sample)))
}
/// lib.rs:1346-1377
pub fn identify_speaker_with_threshold_feats(net: &SimpleNeuralNet, windows: &[Vec<f32>], threshold: f32) -> Option<usize> {
    if net.output_size() <= 1 || windows.is_empty() {
        return None;
    }
    let sums = prob_sums(net, windows);
    let best = argmax_last(&sums);
    if sums[best] / windows.len() as f32 >= threshold { Some(best) } else { None }
}
/// lib.rs:1307-1343
pub fn identify_speaker_with_threshold(net: &SimpleNeuralNet, sample: &[i16], threshold: f32, extractor: &FeatureExtractor) -> Option<usize> {
    if net.output_size() <= 1 {
        return None;
    }
    identify_speaker_with_threshold_feats(net, &extractor.extract(sample), threshold)
}
/// lib.rs:1383-1411 -- extraction, forward, argmax/threshold histogram and the stable sort run behind one C call
pub fn identify_speaker_list(net: &SimpleNeuralNet, sample: &[i16], threshold: f32, _extractor: &FeatureExtractor) -> Vec<usize> {
    let mut out = vec![0u32; net.output_size().max(1)];
    let mut n = 0u32;
    check(unsafe { szb_identify_speaker_list(net.net, sample.as_ptr(), sample.len() as u64, threshold, out.as_mut_ptr(), out.len() as u32, &mut n) })
        .expect("identify_speaker_list");
    out[..n as usize].iter().map(|&i| i as usize).collect()
}

// ---- embeddings, augmentation, window-range extraction ------------------------------------

fn flatten(windows: &[Vec<f32>]) -> Vec<f32> {
    windows.iter().flat_map(|v| v.iter().copied()).collect()
}

impl SimpleNeuralNet {
    /// lib.rs:903
    pub fn embedding_size(&self) -> usize {
        let mut n = 0u32;
        check(unsafe { szb_net_embedding_size(self.net, &mut n) }).expect("embedding_size");
        n as usize
    }
    /// lib.rs:895-900 (ReLU, tanh)
    pub fn embed(&self, bits: &[f32]) -> Vec<f32> {
        let mut out = vec![0f32; self.embedding_size()];
        check(unsafe { szb_net_embed(self.net, bits.as_ptr(), 1, 0, out.as_mut_ptr()) }).expect("embed");
        out
    }
    /// lib.rs:1073-1079 (ReLU, ReLU)
    pub fn forward_embedding(&self, input: &[f32]) -> Vec<f32> {
        let mut out = vec![0f32; self.embedding_size()];
        check(unsafe { szb_net_embed(self.net, input.as_ptr(), 1, 1, out.as_mut_ptr()) }).expect("forward_embedding");
        out
    }
}

/// lib.rs:1453-1475: mean of forward_embedding over the windows, L2-normalised
pub fn extract_embedding_from_features(net: &SimpleNeuralNet, feats: &[Vec<f32>]) -> Vec<f32> {
    let flat = flatten(feats);
    let mut out = vec![0f32; net.embedding_size()];
    check(unsafe { szb_net_embedding_mean(net.net, flat.as_ptr(), feats.len() as u64, out.as_mut_ptr()) }).expect("embedding_mean");
    out
}
/// lib.rs:1478-1500: per-dimension median of forward_embedding, L2-normalised
pub fn median_embedding_from_features(net: &SimpleNeuralNet, feats: &[Vec<f32>]) -> Vec<f32> {
    let flat = flatten(feats);
    let mut out = vec![0f32; net.embedding_size()];
    check(unsafe { szb_net_embedding_median(net.net, flat.as_ptr(), feats.len() as u64, 1, out.as_mut_ptr()) }).expect("embedding_median");
    out
}
/// lib.rs:1418-1450
pub fn extract_embedding(net: &SimpleNeuralNet, sample: &[i16], extractor: &FeatureExtractor) -> Vec<f32> {
    let feats = extractor.extract(sample);
    let flat = flatten(&feats);
    let mut out = vec![0f32; net.embedding_size()];
    check(unsafe { szb_net_embedding_median(net.net, flat.as_ptr(), feats.len() as u64, 0, out.as_mut_ptr()) }).expect("extract_embedding");
    out
}
/// lib.rs:1531-1540
pub fn cosine_similarity(a: &[f32], b: &[f32]) -> f32 {
    unsafe { szb_cosine_similarity(a.as_ptr(), b.as_ptr(), a.len().min(b.len()) as u32) }
}
/// lib.rs:1503-1529; the matching rule runs behind the C ABI (`szb_match_embedding`), this wrapper only flattens the map.
pub fn identify_speaker_from_embedding(emb: &[f32], speaker_embeddings: &std::collections::HashMap<usize, Vec<f32>>, threshold: f32) -> usize {
    let ids: Vec<u64> = speaker_embeddings.keys().map(|&k| k as u64).collect();
    let flat: Vec<f32> = speaker_embeddings.values().flat_map(|v| v.iter().copied()).collect();
    let mut found = u64::MAX;
    check(unsafe {
        szb_match_embedding(emb.as_ptr(), flat.as_ptr(), ids.as_ptr(), ids.len() as u32, emb.len() as u32, threshold, &mut found, std::ptr::null_mut())
    })
    .expect("match_embedding");
    if found == u64::MAX { usize::MAX } else { found as usize }
}
/// lib.rs:103-116 with the draws derived from `seed` (the reference uses thread_rng)
pub fn augment(samples: &[i16], seed: u64) -> Vec<i16> {
    let mut out = vec![0i16; samples.len()];
    check(unsafe { szb_augment(ctx(), samples.as_ptr(), samples.len() as u64, seed, out.as_mut_ptr()) }).expect("augment");
    out
}

impl FeatureExtractor {
    /// Rows `w_begin .. w_end` of `extract`, bit for bit, from the samples that range (and its two-frame halo) covers:
    /// the per-GPU unit of the identification sweep over one long clip (one process per GPU, counts added on the host).
    pub fn extract_range(&self, samples: &[i16], w_begin: usize, w_end: usize) -> Vec<Vec<f32>> {
        let n = w_end.saturating_sub(w_begin);
        let mut flat = vec![0f32; n * FEATURE_SIZE];
        check(unsafe { szb_extract_range(ctx(), samples.as_ptr(), samples.len() as u64, w_begin as u64, w_end as u64, flat.as_mut_ptr(), n as u64) })
            .expect("extract_range");
        flat.chunks(FEATURE_SIZE).map(|c| c.to_vec()).collect()
    }
}
/// Per-class window counts of lib.rs:1389-1400 for already extracted windows (add the ranks' vectors, then sort as lib.rs:1402-1410).
pub fn identify_counts(net: &SimpleNeuralNet, windows: &[Vec<f32>], threshold: f32) -> Vec<u64> {
    let flat = flatten(windows);
    let mut counts = vec![0u64; net.output_size()];
    check(unsafe { szb_identify_counts(net.net, flat.as_ptr(), windows.len() as u64, threshold, counts.as_mut_ptr()) }).expect("identify_counts");
    counts
}
