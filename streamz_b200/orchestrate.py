"""The training / evaluation run of the reference CLI on top of the library (streamz-rs/src/main.rs).

Only the orchestration of the hot path is mirrored -- which clips are extracted together, which windows train which class
in which order, when a class is added -- so that a maintainer can see the whole `main()` flow expressed in C-ABI calls:

    main.rs:490-510   batch_resample + the rayon extraction loop          -> extract_feature_map (ONE batched call per rate)
    main.rs:517-519   burn-in limit                                        -> burn_in_limit
    main.rs:651-668   initial training over the labelled files             -> initial_training (train_from_feature_map)
    main.rs:750-835   incremental pass: embedding, speaker assignment,     -> incremental_training
                      class growth during burn-in, 5 epochs per file
    main.rs:560-640   evaluation by cosine similarity to saved embeddings  -> evaluate

File decoding, the WAV cache, list files, progress bars and the hidden-payload layer stay with the host application.
The reference runs the incremental pass under rayon with locks, so its file order is not defined; here files are visited
in list order (one of the orders the reference can produce), and every random draw is seeded."""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import api

DEFAULT_CONF_THRESHOLD = 0.8     # main.rs:26
DEFAULT_BURN_IN_FRAC = 0.2       # main.rs:29
TRAIN_EPOCHS = 100               # main.rs:32
DROPOUT_PROB = api.DEFAULT_DROPOUT   # main.rs:34 (= lib.rs:36)
BATCH_SIZE = 8                   # main.rs:36
INCREMENTAL_EPOCHS = 5           # main.rs:810
MIN_WINDOWS = 5                  # main.rs:756


def burn_in_limit(dataset_size: int, override: Optional[int] = None) -> int:
    """main.rs:517-519: ceil(n * 0.2) clamped to [10, 50] unless --burn-in-limit is given."""
    if override is not None:
        return int(override)
    return min(50, max(10, int(math.ceil(np.float32(dataset_size) * np.float32(DEFAULT_BURN_IN_FRAC)))))


def count_speakers(files: Sequence[Tuple[str, Optional[int]]]) -> int:
    """main.rs:129-135."""
    return len({c for _, c in files if c is not None})


def extract_feature_map(clips: Dict[str, Tuple[np.ndarray, int]], extractor: api.FeatureExtractor) -> Dict[str, np.ndarray]:
    """main.rs:490-510: `clips[path] = (mono i16 PCM, sample rate)` -> `feature_map[path] = [n, 60]`.  All clips of one rate
    go through a single szb_extract_batch call (resampling included) instead of one `extract` per rayon task."""
    by_rate: Dict[int, List[str]] = {}
    for path, (_, rate) in clips.items():
        by_rate.setdefault(int(rate), []).append(path)
    out: Dict[str, np.ndarray] = {}
    for rate, paths in by_rate.items():
        feats = extractor.extract_batch([clips[p][0] for p in paths], rate)
        out.update(zip(paths, feats))
    return out


def initial_training(net: api.SimpleNeuralNet, feature_map: Dict[str, np.ndarray], train_files: Sequence[Tuple[str, Optional[int]]],
                     epochs: int = TRAIN_EPOCHS, lr: float = 0.01, dropout: float = DROPOUT_PROB, batch_size: int = BATCH_SIZE,
                     seed: int = 0) -> Optional[float]:
    """main.rs:651-668: every labelled file, in list order, for `epochs` epochs each; None when nothing is labelled."""
    refs = [(p, c) for p, c in train_files if c is not None]
    if not refs:
        return None
    return api.train_from_feature_map(net, feature_map, refs, epochs, lr, dropout, batch_size, seed=seed)


def incremental_training(net: api.SimpleNeuralNet, train_files: List[List], feature_map: Dict[str, np.ndarray], limit: int,
                         conf_threshold: float = DEFAULT_CONF_THRESHOLD, dropout: float = DROPOUT_PROB, batch_size: int = BATCH_SIZE,
                         epochs: int = INCREMENTAL_EPOCHS, seed: int = 0, new_columns: Optional[Iterable[np.ndarray]] = None,
                         speaker_embeddings: Optional[Dict[int, np.ndarray]] = None) -> dict:
    """main.rs:750-835.  `train_files` is a list of [path, class or None] and is updated in place with the assigned classes
    (the reference rewrites train_files.txt from it, main.rs:866-872).  `new_columns`: explicit output-layer columns for the
    classes added on the way (the reference draws them from thread_rng; default: the library's seeded initialiser).
    Returns the bookkeeping of the run: losses, per-speaker embeddings, the decision taken for every file."""
    cols = iter(new_columns) if new_columns is not None else None
    speaker_features: Dict[int, List[np.ndarray]] = {}
    embeds: Dict[int, np.ndarray] = dict(speaker_embeddings or {})
    total_loss, count, log = 0.0, 0, []

    def grow() -> int:
        label = net.output_size()
        col = next(cols) if cols is not None else None
        net.add_output_class(col, seed=seed + label)
        return label

    for entry in train_files:
        path, cls = entry[0], entry[1]
        windows = feature_map.get(path)
        if windows is None:                                      # main.rs:829 "Missing audio"
            log.append((path, "missing", None))
            continue
        if len(windows) < MIN_WINDOWS:                           # main.rs:756-760
            log.append((path, "too short", None))
            continue
        emb = api.extract_embedding_from_features(net, windows)  # main.rs:763-767 (already unit length; normalize is idempotent)
        burn = count < limit                                     # main.rs:769-770
        threshold = 0.5 if burn else conf_threshold              # main.rs:771-775
        if burn and cls is None:                                 # main.rs:778-785: a new class for every unlabelled burn-in file
            speaker = grow()
            net.record_training_file(speaker, path)
            how = "new (burn-in)"
        elif cls is not None:                                    # main.rs:786-787
            speaker, how = int(cls), "labelled"
        else:                                                    # main.rs:788-798
            matched = api.identify_speaker_from_embedding(emb, embeds, threshold)
            if matched is None or matched >= net.output_size():
                speaker, how = grow(), "new (no match)"
            else:
                speaker, how = int(matched), "matched"
        entry[1] = speaker
        lr = 0.05 if count < 1000 else 0.01                      # main.rs:801
        loss = api.pretrain_from_features(net, windows, speaker, net.output_size(), epochs, lr, dropout, batch_size,
                                          seed=seed * 7919 + count)      # main.rs:804-813
        net.record_training_file(speaker, path)                  # main.rs:814
        total_loss += loss
        speaker_features.setdefault(speaker, []).append(emb)     # main.rs:818-823
        embeds[speaker] = average_vectors(speaker_features[speaker])
        count += 1                                               # main.rs:825 (recompute_embeddings every 100 files re-derives the same averages)
        log.append((path, how, speaker))
    return {"total_loss": total_loss, "count": count, "mean_loss": total_loss / count if count else None,
            "speaker_features": speaker_features, "speaker_embeddings": embeds, "log": log}


def average_vectors(vectors: Sequence[np.ndarray]) -> np.ndarray:
    """lib.rs:144-159: element-wise mean, then L2 normalisation when the norm exceeds 1e-6 (lib.rs:132-139)."""
    if not len(vectors):
        return np.zeros(0, np.float32)
    avg = np.zeros_like(np.asarray(vectors[0], np.float32))
    for v in vectors:
        avg = avg + np.asarray(v, np.float32)
    avg = avg / np.float32(len(vectors))
    norm = np.sqrt((avg * avg).sum(dtype=np.float32))
    return avg / norm if norm > 1e-6 else avg


def evaluate(net: api.SimpleNeuralNet, speaker_embeddings: Dict[int, np.ndarray], target_files: Sequence[Tuple[str, int]],
             feature_map: Dict[str, np.ndarray], conf_threshold: float = DEFAULT_CONF_THRESHOLD) -> dict:
    """main.rs:583-640: a file is assigned to the most similar stored speaker embedding above the threshold."""
    tp = fp = fn = correct = 0
    for path, true_class in target_files:
        windows = feature_map.get(path)
        if windows is None:
            continue
        emb = api.extract_embedding_from_features(net, windows)
        best_id, best_sim = None, -np.inf
        for sid, centroid in speaker_embeddings.items():
            sim = api.cosine_similarity(emb, centroid)
            if sim > conf_threshold and sim > best_sim:
                best_sim, best_id = sim, sid
        if best_id == true_class:
            correct += 1; tp += 1
        elif best_id is None:
            fn += 1
        else:
            fp += 1
    total = max(1, len(target_files))
    precision, recall = tp / max(1, tp + fp), tp / max(1, tp + fn)
    return {"accuracy": correct / total, "precision": precision, "recall": recall,
            "f1": 2 * precision * recall / max(precision + recall, 1e-6)}


def training_run(clips: Dict[str, Tuple[np.ndarray, int]], train_files: List[List], net: Optional[api.SimpleNeuralNet] = None,
                 ctx: Optional[api.Context] = None, initial_epochs: int = TRAIN_EPOCHS, burn_in: Optional[int] = None,
                 conf_threshold: float = DEFAULT_CONF_THRESHOLD, seed: int = 0, model_path: Optional[str] = None) -> dict:
    """The non-eval branch of main(): extract everything once, create the net when there is no model (main.rs:641-649),
    initial training, incremental pass, final speaker embeddings, optional save (main.rs:838-858)."""
    ctx = ctx or api.default_context()
    extractor = api.FeatureExtractor(ctx)
    feature_map = extract_feature_map(clips, extractor)
    limit = burn_in_limit(len(train_files), burn_in)
    fresh = net is None
    if fresh:
        n_spk = count_speakers(train_files)
        if n_spk == 0:                                           # main.rs:643-647
            n_spk = 1
            train_files[0][1] = 0
        net = api.SimpleNeuralNet(api.FEATURE_SIZE, 512, 256, max(1, n_spk), seed=seed, ctx=ctx)
    init_loss = initial_training(net, feature_map, train_files, epochs=initial_epochs, seed=seed) if fresh else None
    inc = incremental_training(net, train_files, feature_map, limit, conf_threshold=conf_threshold, seed=seed)
    final = api.compute_speaker_embeddings(net, feature_map)
    net.set_embeddings(final)                                    # main.rs:854: saved with the model (lib.rs:1114-1127)
    if model_path:
        net.save(model_path)
    return {"net": net, "feature_map": feature_map, "burn_in_limit": limit, "initial_loss": init_loss, "incremental": inc,
            "speaker_embeddings": final, "train_files": train_files}
