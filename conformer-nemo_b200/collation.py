"""Length-bucketed batching / collation on the host: what feeds the sharded forward from real
manifests instead of synthetic tensors.  Mirrors, for inference,

* manifest parsing and filtering   common/parts/preprocessing/manifest.py:34-112 (``item_iter`` / ``__parse_item``),
                                   common/parts/preprocessing/collections.py:88-190 (``AudioText``: min / max duration,
                                   ``max_number``, ``do_sort_by_duration``, entries whose parser returns None dropped)
* padding rules                    ``_speech_collate_fn`` (asr/data/audio_to_text.py:48-99): signals zero-padded on the right
                                   to the longest of the batch, tokens padded with ``pad_id``, 4- or 5-tuples, no-audio batches
* chunked bucketing                ``BucketingIterator`` (asr/data/audio_to_text.py:1515-1533)

and adds the B200 side the reference leaves to ``DataLoader(pin_memory=True)``: ``CollationService`` plans length-bucketed
sub-batches per rank with the same cost model as the sharded forward (sharding.plan_shards), collates each sub-batch
directly into a ring of pinned host buffers and copies it to the device on a side stream, so the copy of sub-batch i+1
overlaps the forward of sub-batch i.  Host code only: no kernels, no collective.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Any, Callable, Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import torch

from .sharding import ShardPlan, plan_shards


# ---- manifests ------------------------------------------------------------------------------------------------------------

@dataclass
class ManifestEntry:
    """The fields of ``AudioText.OUTPUT_TYPE`` (collections.py:91-94)."""
    id: int
    audio_file: str
    duration: float
    text_tokens: List[int]
    offset: Optional[float]
    text_raw: str
    speaker: Optional[int] = None
    orig_sr: Optional[int] = None
    lang: Optional[str] = None


def _parse_item(line: str, manifest_file: str) -> dict:
    item = json.loads(line)
    if "audio_filename" in item:
        item["audio_file"] = item.pop("audio_filename")
    elif "audio_filepath" in item:
        item["audio_file"] = item.pop("audio_filepath")
    else:  # manifest.py:86-89
        raise ValueError(f"Manifest file {manifest_file} has invalid json line structure: {line} without proper audio file key.")
    item["audio_file"] = os.path.expanduser(item["audio_file"])
    if "duration" not in item:  # manifest.py:93-96
        raise ValueError(f"Manifest file {manifest_file} has invalid json line structure: {line} without proper duration key.")
    if "text" in item:
        pass
    elif "text_filepath" in item:
        with open(item.pop("text_filepath"), "r") as f:
            item["text"] = f.read().replace("\n", "")
    elif "normalized_text" in item:
        item["text"] = item["normalized_text"]
    return dict(audio_file=item["audio_file"], duration=item["duration"], text=item.get("text", ""),
                offset=item.get("offset", None), speaker=item.get("speaker", None), orig_sr=item.get("orig_sample_rate", None),
                lang=item.get("lang", None))


def read_manifest(manifest_files: Union[str, Sequence[str]], parser: Optional[Callable[[str], Optional[List[int]]]] = None,
                  min_duration: Optional[float] = None, max_duration: Optional[float] = None, max_number: Optional[int] = None,
                  do_sort_by_duration: bool = False) -> List[ManifestEntry]:
    """JSON-lines manifests -> entries, with the reference's filters in the reference's order.  ``manifest_files`` may be
    a comma-separated string like ``manifest_filepath`` in the configs (audio_to_text.py:273 splits on ',')."""
    if isinstance(manifest_files, str):
        manifest_files = manifest_files.split(",")
    data: List[ManifestEntry] = []
    k = -1
    done = False
    for manifest_file in manifest_files:
        if done:
            break
        with open(os.path.expanduser(manifest_file), "r") as f:
            for line in f:
                k += 1  # ids count every line, filtered or not (manifest.py:69-72)
                item = _parse_item(line, manifest_file)
                duration = item["duration"]
                if min_duration is not None and duration < min_duration:
                    continue
                if max_duration is not None and duration > max_duration:
                    continue
                text = item["text"]
                tokens = (parser(text) if parser is not None else []) if text != "" else []
                if tokens is None:  # collections.py:160-163
                    continue
                data.append(ManifestEntry(k, item["audio_file"], duration, list(tokens), item["offset"], text, item["speaker"],
                                          item["orig_sr"], item["lang"]))
                if len(data) == max_number:  # collections.py:172-173
                    done = True
                    break
    if do_sort_by_duration:
        data.sort(key=lambda e: e.duration)
    return data


# ---- collation ---------------------------------------------------------------------------------------------------------------

def speech_collate(batch: Sequence[tuple], pad_id: int, audio_out: Optional[torch.Tensor] = None):
    """``_speech_collate_fn`` (audio_to_text.py:48-99): batch of (signal, signal_len, tokens, tokens_len[, sample_id]).
    ``audio_out``: optional preallocated (>= B, >= max_len) buffer (e.g. pinned) the padded signals are written into; the
    returned ``audio_signal`` is then a view of it."""
    packed = list(zip(*batch))
    if len(packed) == 5:
        _, audio_lengths, _, tokens_lengths, sample_ids = packed
    elif len(packed) == 4:
        sample_ids = None
        _, audio_lengths, _, tokens_lengths = packed
    else:
        raise ValueError("Expects 4 or 5 tensors in the batch!")
    has_audio = audio_lengths[0] is not None
    max_audio_len = max(audio_lengths).item() if has_audio else 0
    max_tokens_len = max(tokens_lengths).item()
    n = len(batch)
    audio_signal = None
    if has_audio:
        first = batch[0][0]
        if audio_out is not None:
            if audio_out.dim() != 2 or audio_out.size(0) < n or audio_out.size(1) < max_audio_len:
                raise ValueError(f"audio_out {tuple(audio_out.shape)} cannot hold a ({n}, {max_audio_len}) batch")
            audio_signal = audio_out[:n, :max_audio_len]
        else:
            audio_signal = torch.empty(n, max_audio_len, dtype=first.dtype)
    tokens = torch.full((n, max_tokens_len), pad_id, dtype=batch[0][2].dtype)
    for i, b in enumerate(batch):
        sig, sig_len, tokens_i, tokens_i_len = b[:4]
        # the reference pads by (max - stated length) and then stacks, so a stated length that differs from the tensor's
        # size fails in torch.stack there; same condition, same exception class here
        if has_audio:
            m = sig.numel()
            if m != sig_len.item():
                raise RuntimeError(f"stack expects each tensor to be equal size: signal {i} has {m} samples, length says {sig_len.item()}")
            audio_signal[i, :m] = sig
            audio_signal[i, m:] = 0
        k = tokens_i.numel()
        if k != tokens_i_len.item():
            raise RuntimeError(f"stack expects each tensor to be equal size: tokens {i} has {k} entries, length says {tokens_i_len.item()}")
        tokens[i, :k] = tokens_i
    audio_lengths = torch.stack(audio_lengths) if has_audio else None
    tokens_lengths = torch.stack(tokens_lengths)
    if sample_ids is None:
        return audio_signal, audio_lengths, tokens, tokens_lengths
    return audio_signal, audio_lengths, tokens, tokens_lengths, torch.tensor(sample_ids, dtype=torch.int32)


class BucketingIterator:
    """audio_to_text.py:1515-1533: consecutive chunks of ``bucketing_batch_size`` samples; the last one may be short."""

    def __init__(self, wrapped_iter: Iterable, bucketing_batch_size: int):
        self.wrapped_iter = iter(wrapped_iter)
        self.bucketing_batch_size = bucketing_batch_size

    def __iter__(self):
        return self

    def __next__(self):
        batches = []
        for _ in range(self.bucketing_batch_size):
            try:
                batches.append(next(self.wrapped_iter))
            except StopIteration:
                break
        if not batches:
            raise StopIteration
        return batches


# ---- the service -----------------------------------------------------------------------------------------------------------------

def feature_frames(n_samples: int, hop: int = 160) -> int:
    """Frames the preprocessor makes of ``n_samples`` samples: floor(len / hop) + 1 (features.py:347-353)."""
    return n_samples // hop + 1


@dataclass
class StagedBatch:
    indices: List[int]                    # positions in the caller's entry list
    audio_signal: torch.Tensor            # (B, L) on `device` (or pinned host memory when device is None)
    audio_lengths: torch.Tensor           # (B,) int64, same place
    tokens: torch.Tensor                  # (B, U) host
    tokens_lengths: torch.Tensor          # (B,) host
    sample_ids: torch.Tensor              # (B,) int32 host = manifest ids
    ready: Any = None                     # torch.cuda.Event recorded after the copies (None on the host)
    slot: int = 0
    audio_lengths_host: List[int] = field(default_factory=list)   # the same lengths as host integers (no sync to get them)

    def feature_lengths_host(self, hop: int = 160) -> List[int]:
        """Frames the preprocessor makes of every utterance (features.py:347-353), as host integers: what
        ``ConformerEncoder.forward(..., length_host=)`` needs to run a ragged sub-batch in its packed layout."""
        return [feature_frames(n, hop) for n in self.audio_lengths_host]


class CollationService:
    """Plans and stages one rank's share of a list of utterances.

    ``lengths``: samples per utterance (e.g. ``int(duration * sample_rate)`` from the manifest) -- the plan is made from
    lengths alone, identically on every rank, with no communication.
    ``load(i) -> (signal 1-D float tensor, tokens 1-D int tensor)`` reads utterance i (decoding audio files is the caller's
    business: the reference does it in ``WaveformFeaturizer``, out of scope here).
    Iterating yields ``StagedBatch`` es in plan order; with a CUDA ``device`` the signals are collated into one of
    ``depth`` pinned buffers and copied on a side stream -- call ``wait(batch)`` before using them on the compute stream.
    A pinned slot is overwritten only after its previous copy has completed.
    """

    def __init__(self, lengths: Sequence[int], load: Callable[[int], Tuple[torch.Tensor, torch.Tensor]], n_ranks: int = 1,
                 rank: int = 0, max_batch: int = 64, bucket_frames="auto", hop: int = 160, pad_id: int = 0, device=None,
                 depth: int = 2, sample_ids: Optional[Sequence[int]] = None, dtype=torch.float32):
        if not 0 <= rank < n_ranks:
            raise ValueError(f"rank {rank} outside [0, {n_ranks})")
        self.lengths = [int(v) for v in lengths]
        self.load = load
        self.pad_id = pad_id
        self.rank = rank
        self.sample_ids = list(sample_ids) if sample_ids is not None else list(range(len(self.lengths)))
        frames = [feature_frames(v, hop) for v in self.lengths]
        self.plan: ShardPlan = plan_shards(frames, n_ranks, max_batch=max_batch, bucket_frames=bucket_frames)
        self.batches: List[List[int]] = self.plan.batches[rank]
        self.device = torch.device(device) if device is not None else None
        self.depth = max(1, depth)
        self.dtype = dtype
        rows = max((len(b) for b in self.batches), default=0)
        cols = max((self.lengths[i] for b in self.batches for i in b), default=0)
        self._use_cuda = self.device is not None and self.device.type == "cuda"
        pin = self._use_cuda and torch.cuda.is_available()
        # flat slots: every sub-batch is collated into the CONTIGUOUS (n, max_len) prefix view of its slot, so the
        # host-to-device copy is one cudaMemcpyAsync from pinned memory (a strided view of a (rows, cols) slot would make
        # torch stage a pageable contiguous temporary first and copy synchronously)
        self._audio = [torch.empty(rows * cols, dtype=dtype, pin_memory=pin) for _ in range(self.depth)]
        self._lens = [torch.empty(rows, dtype=torch.int64, pin_memory=pin) for _ in range(self.depth)]
        self._busy = [None] * self.depth  # event of the last copy out of each pinned slot
        self._stream = torch.cuda.Stream(self.device) if self._use_cuda else None

    def __len__(self):
        return len(self.batches)

    def padding_fraction(self) -> float:
        """Share of the staged samples that is padding (what the length buckets keep small)."""
        real = sum(self.lengths[i] for b in self.batches for i in b)
        padded = sum(len(b) * max(self.lengths[i] for i in b) for b in self.batches)
        return 0.0 if padded == 0 else 1.0 - real / padded

    def _collate(self, idx: List[int], slot: int):
        items = []
        for i in idx:
            sig, tok = self.load(i)
            if sig.dim() != 1 or sig.numel() != self.lengths[i]:
                raise ValueError(f"utterance {i}: load() returned {tuple(sig.shape)}, expected ({self.lengths[i]},)")
            items.append((sig.to(self.dtype), torch.tensor(sig.numel()), tok, torch.tensor(tok.numel()), self.sample_ids[i]))
        n, max_len = len(idx), max(self.lengths[i] for i in idx)
        return speech_collate(items, self.pad_id, audio_out=self._audio[slot][: n * max_len].view(n, max_len))

    def __iter__(self) -> Iterator[StagedBatch]:
        for n, idx in enumerate(self.batches):
            slot = n % self.depth
            if self._busy[slot] is not None:
                self._busy[slot].synchronize()  # the previous copy out of this pinned slot has completed
                self._busy[slot] = None
            audio, lens, tokens, tok_lens, ids = self._collate(idx, slot)
            self._lens[slot][:len(idx)] = lens
            lens = self._lens[slot][:len(idx)]
            ready = None
            if self._use_cuda:
                with torch.cuda.stream(self._stream):
                    assert audio.is_contiguous() and audio.is_pinned()  # one asynchronous copy straight from the slot
                    audio = audio.to(self.device, non_blocking=True)
                    lens = lens.to(self.device, non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(self._stream)
                self._busy[slot] = ready
            yield StagedBatch(list(idx), audio, lens, tokens, tok_lens, ids, ready, slot, [self.lengths[i] for i in idx])

    def wait(self, batch: StagedBatch, stream=None):
        """Makes ``stream`` (default: the current stream) wait for the batch's copies and tells the caching allocator that
        the batch's device tensors are used there."""
        if batch.ready is not None:
            stream = stream or torch.cuda.current_stream(self.device)
            stream.wait_event(batch.ready)
            batch.audio_signal.record_stream(stream)
            batch.audio_lengths.record_stream(stream)
