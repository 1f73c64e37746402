"""Batched, on-device streaming loader behind the reference's `streaming_dataloader(...)` signature
(src/pretraining/dataloader/pretraining_dataloader.py:304-382).

Same data semantics as the reference StreamingDataset (:186-301): chunk files
`data/nasa_power/processed/weather_dataset_weekly_{id}.pt` (TensorDataset(weather, coords, index)),
per-rank contiguous chunk partition, year = 1984 + ((idx*365 + t) * interval) / 365 in float32, mask drawn
for the WHOLE chunk, then `randperm` shuffle, then the `< cutoff_year` filter, then batches of
`batch_size` that run across chunk boundaries; the RNG is consumed in the same order (mask, randperm), and
the masks are bit-identical to the reference's torch.rand-based functions for the same generator state
(CUDA: Philox replay kernels wm_mask_bert / wm_mask_former; CPU: torch.rand itself).
What changed is the mechanics: no per-sample Python loop, no per-sample host sync, no per-sample collate, and the
NEXT chunk file is read, pinned and copied to the device by a helper thread on a side stream while the current
chunk's batches train (the reference blocks the training loop for every torch.load).
"""
import logging
import os
import random
import struct
import threading
import zipfile
from typing import Iterator, List, Optional, Tuple

import torch

from ...utils.constants import DATA_DIR, DRY_RUN, DRY_RUN_TRAIN_CHUNK_IDS, NUM_DATASET_PARTS, VALIDATION_CHUNK_IDS

random.seed(1234)
logger = logging.getLogger(__name__)


class _PinnedPool:
    """Page-locked staging buffers shared by all loaders of the process (train + validation loaders are alive together,
    a new pair per epoch): pinning 185 MB costs more than copying it, so a buffer is pinned once and re-used. A buffer
    handed out again waits for the device copy that last read it."""

    def __init__(self):
        self._lock = threading.Lock()
        self._bufs = []  # [tensor(uint8, pinned), last-use event or None, busy]

    def acquire(self, nbytes: int):
        with self._lock:
            for b in self._bufs:
                if not b[2] and b[0].numel() >= nbytes:
                    b[2] = True
                    break
            else:
                b = [torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8).pin_memory(), None, True]
                self._bufs.append(b)
        if b[1] is not None:
            b[1].synchronize()
        return b

    def release(self, b, event):
        with self._lock:
            b[1], b[2] = event, False


_PINNED = _PinnedPool()


def _stored_payload_span(path: str, nbytes: int) -> Optional[int]:
    """Byte offset, inside a torch.save zip file, of the one STORED record of exactly `nbytes` bytes (the weather tensor's
    storage: every other record of a chunk file is smaller). None if the file is not laid out like that."""
    try:
        with zipfile.ZipFile(path) as zf:
            hits = [i for i in zf.infolist() if i.file_size == nbytes and i.compress_type == zipfile.ZIP_STORED and "/data/" in i.filename]
        if len(hits) != 1:
            return None
        with open(path, "rb") as f:
            f.seek(hits[0].header_offset)
            head = f.read(30)
        if len(head) != 30 or head[:4] != b"PK\x03\x04":
            return None
        name_len, extra_len = struct.unpack("<HH", head[26:30])
        return hits[0].header_offset + 30 + name_len + extra_len
    except Exception:  # noqa: BLE001
        return None


class StreamingDataset(torch.utils.data.IterableDataset):
    """Iterating yields BATCHES (weather [B,365,31], coords [B,2], year [B,365], interval [B,1], mask [B,365,31])."""

    def __init__(self, file_paths, num_input_features=None, num_output_features=None, shuffle=False,
                 masking_function: Optional[str] = None, masking_prob: float = 0.15, n_masked_features: int = 1,
                 rank: int = 0, cutoff_year: float = 2002.0, batch_size: int = 1):
        self.file_paths = list(file_paths)
        self.num_input_features = num_input_features
        self.num_output_features = num_output_features
        self.shuffle = shuffle
        self.masking_prob = masking_prob
        self.n_masked_features = n_masked_features
        self.rank = rank
        self.cutoff_year = cutoff_year
        self.batch_size = batch_size
        self.device = f"cuda:{rank}" if torch.cuda.is_available() else "cpu"
        table = {"weatherbert": self.weatherbert_masking_function,
                 "weatherformer": self.weatherformer_masking_function,
                 "simmtm": self.simmtm_masking_function}
        if masking_function not in table:
            raise ValueError(f"Masking function {masking_function} is not valid")
        self.masking_function = table[masking_function]
        # The first chunk starts loading NOW (no RNG involved): BaseTrainer builds the train and validation loaders at
        # the top of an epoch, so the validation loader's first file is read while the training epoch runs and the
        # training loader's while the previous epoch's bookkeeping finishes.
        self._first = None
        if len(self.file_paths) >= 3 and torch.device(self.device).type == "cuda":
            if os.path.exists(self.file_paths[1]):
                self._first = self._start_prefetch(self.file_paths[1])

    # ---- masks (reference :56-84) ------------------------------------------------------------------
    def weatherbert_masking_function(self, seq_len, n_features, batch_size):
        if torch.device(self.device).type == "cuda":
            from ... import ops
            return ops.mask_bert(seq_len, n_features, batch_size, self.masking_prob, device=self.device)
        return torch.rand(batch_size, seq_len, n_features, device=self.device) < self.masking_prob

    def weatherformer_masking_function(self, seq_len, n_features, batch_size):
        if torch.device(self.device).type == "cuda":
            from ... import ops
            return ops.mask_former(seq_len, n_features, batch_size, self.n_masked_features, device=self.device)
        order = torch.argsort(torch.rand(batch_size, n_features, device=self.device), dim=-1)
        return (order < self.n_masked_features).unsqueeze(1).expand(-1, seq_len, -1)

    def simmtm_masking_function(self, seq_len, n_features, batch_size):
        """Contiguous time segments (geometric lengths, mean 5), the same positions for every feature, trimmed to
        int(seq_len * masking_prob) positions per sample (reference :86-184). Not a hot kernel -- once per chunk
        -- so it stays torch ops on the loader's device, drawing from the generator in the reference's order
        (Geometric.sample, rand, and rand_like only when a sample has to be trimmed): same seed, same masks."""
        dev = self.device
        target = int(seq_len * self.masking_prob)
        if target == 0:
            return torch.zeros(batch_size, seq_len, n_features, dtype=torch.bool, device=dev)
        nseg = max(1, target // 5 + 5)  # candidate segments per sample
        total = batch_size * nseg
        length = torch.distributions.Geometric(probs=torch.tensor(1 / 5, device=dev)).sample((total,)).int()
        length = torch.clamp(length, min=1, max=seq_len)
        room = torch.clamp(seq_len - length, min=0)
        start = (torch.rand(total, device=dev) * (room + 1).float()).long()
        length, start = length.view(batch_size, nseg), start.view(batch_size, nseg)
        order = torch.argsort(start, dim=-1)
        start, length = torch.gather(start, -1, order), torch.gather(length, -1, order)
        end = start + length
        prev_end = torch.cat([torch.zeros(batch_size, 1, device=dev, dtype=torch.long), end[:, :-1]], dim=-1)
        keep_seg = (start >= prev_end).unsqueeze(-1)  # drop a segment that starts inside its predecessor
        pos = torch.arange(seq_len, device=dev)[None, None, :]
        hit = ((pos >= start.unsqueeze(-1)) & (pos < end.unsqueeze(-1)) & keep_seg).any(dim=1)  # [B, S]
        too_many = hit.sum(dim=1) > target
        if too_many.any():  # keep a random subset of exactly `target` masked positions in those samples
            r = torch.where(hit, torch.rand_like(hit.float()), torch.inf)
            rank = torch.argsort(torch.argsort(r, dim=1), dim=1)
            hit = torch.where(too_many.unsqueeze(1), rank < target, hit)
        return hit.unsqueeze(-1).expand(-1, -1, n_features)

    # ---- chunk -> tensors -------------------------------------------------------------------------
    def _load_chunk(self, path) -> Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """File -> (weather, coords, index) float32 on self.device. On CUDA the file is read to the host, pinned and
        copied on a side stream; the returned tensors carry an event (`_ready`) the consumer stream waits on. No RNG
        is touched here, so running it one chunk ahead in a helper thread leaves the mask / randperm order alone."""
        on_cuda = torch.device(self.device).type == "cuda"
        data = None
        if on_cuda:
            try:  # map the file instead of reading it: the only copy of the payload is then the one to the device
                data = torch.load(path, weights_only=False, map_location="cpu", mmap=True)
            except Exception:  # noqa: BLE001 -- legacy (non-zip) files cannot be mapped
                data = None
        if data is None:
            data = torch.load(path, weights_only=False, map_location="cpu" if on_cuda else self.device)
        if hasattr(data, "tensors"):
            weather, coords, index = data.tensors[:3]
        else:
            if len(data) == 0:
                return None
            weather = torch.stack([s[0] for s in data])
            coords = torch.stack([s[1] for s in data])
            index = torch.stack([s[2] for s in data])
        if weather.shape[0] == 0:
            return None
        if not on_cuda:
            return weather.to(self.device).float(), coords.to(self.device).float(), index.to(self.device).float()
        from ...graph_step import CAPTURE_LOCK  # no CUDA work of this helper thread while a training step is being recorded

        with CAPTURE_LOCK, torch.cuda.device(self.device):
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            with torch.cuda.stream(self._copy_stream):
                big = self._weather_via_pinned(path, weather)
                # (fallback) pageable -> device on the side stream: the driver stages the copy itself and blocks only this
                # helper thread, with the GIL released
                out = [big if big is not None else weather.to(self.device).float(),
                       coords.to(self.device).float(), index.to(self.device).float()]
                ready = torch.cuda.Event()
                ready.record(self._copy_stream)
                if big is not None:
                    _PINNED.release(self._staging, ready)
        # The cutoff-year filter almost never drops anything; whether it does is decided HERE, on the host copy of the
        # index columns and with the same float32 operations in the same order as the device expression in
        # _chunk_samples (IEEE mul / add / div on both sides: identical results), so that the training loop does not have
        # to read a flag back from the GPU -- that read drained the launch queue once per chunk.
        seq_len = weather.shape[1]
        idx = index.float()
        t = torch.arange(seq_len, dtype=torch.float32)
        years_host = 1984.0 + ((idx[:, 0:1] * 365 + t) * idx[:, 1:2]) / 365
        keep_all = bool((years_host.max(dim=1).values < self.cutoff_year).all())
        if os.environ.get("WM_LOADER_HOST_CUTOFF", "1") == "0":  # (A/B switch: let the device decide, with its host sync)
            keep_all = None
        return out[0], out[1], out[2], ready, keep_all

    def _weather_via_pinned(self, path, weather) -> Optional[torch.Tensor]:
        """The weather tensor of a mapped chunk file, read with plain read() calls straight into a re-used page-locked
        buffer and copied to the device from there (current stream). Touching the mapping instead -- 45,000 page faults
        for 185 MB, then the driver's staged pageable copy -- took 0.16-0.44 s per chunk depending on the box, which
        nothing hides for the first chunk of an epoch. None: layout not recognised, the caller copies the mapped tensor."""
        if weather.dtype != torch.float32 or not weather.is_contiguous():
            return None
        nbytes_storage = weather.untyped_storage().nbytes()
        off = _stored_payload_span(path, nbytes_storage)
        if off is None:
            return None
        nbytes = weather.numel() * 4
        off += weather.storage_offset() * 4
        buf = _PINNED.acquire(nbytes)
        try:
            view = memoryview(buf[0].numpy())[:nbytes]
            got = 0
            with open(path, "rb", buffering=0) as f:
                f.seek(off)
                while got < nbytes:
                    k = f.readinto(view[got:])
                    if not k:
                        raise IOError(f"short read from {path}")
                    got += k
            host = buf[0][:nbytes].view(torch.float32).view(weather.shape)
            dev = host.to(self.device, non_blocking=True)
        except Exception:  # noqa: BLE001
            _PINNED.release(buf, None)
            return None
        self._staging = buf
        return dev

    def _start_prefetch(self, path):
        box = {}

        def work():
            try:
                box["v"] = self._load_chunk(path)
            except BaseException as e:  # re-raised in the consumer
                box["e"] = e

        th = threading.Thread(target=work, daemon=True)
        th.start()
        return th, box

    def _chunk_samples(self, path, loaded="load"):
        if loaded == "load":
            loaded = self._load_chunk(path)
        if loaded is None:
            return None
        if len(loaded) == 5:  # CUDA: order this stream after the side-stream copy; the chunk then belongs to it
            torch.cuda.current_stream(torch.device(self.device)).wait_event(loaded[3])
            for t in loaded[:3]:
                t.record_stream(torch.cuda.current_stream(torch.device(self.device)))
        keep_all = loaded[4] if len(loaded) == 5 else None  # (CUDA prefetch path: decided on the host, see _load_chunk)
        weather, coords, index = loaded[:3]
        n, seq_len, n_features = weather.shape
        interval = index[:, 1:2].contiguous()
        t = torch.arange(seq_len, dtype=torch.float32, device=self.device)
        # same float32 op order as the reference's per-sample expression (:251-256)
        years = 1984.0 + ((index[:, 0:1] * 365 + t) * interval) / 365
        mask = self.masking_function(seq_len, n_features, n)
        if self.shuffle and n > 1:
            perm = torch.randperm(n, device=self.device)
            weather, coords, years, interval, mask = weather[perm], coords[perm], years[perm], interval[perm], mask[perm]
        if keep_all is not True:
            keep = years.max(dim=1).values < self.cutoff_year
            if not bool(keep.all()):  # one host sync per chunk (the reference syncs once per sample)
                weather, coords, years, interval, mask = weather[keep], coords[keep], years[keep], interval[keep], mask[keep]
        return weather, coords, years, interval, mask

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        carry: Optional[List[torch.Tensor]] = None
        n_groups = len(self.file_paths) // 3
        weekly = [self.file_paths[3 * gi + 1] for gi in range(n_groups)]  # the weekly file of each triple (:198)
        pending, self._first = (self._first or self._start_prefetch(weekly[0])) if n_groups else None, None
        for gi in range(n_groups):
            th, box = pending
            th.join()
            if "e" in box:
                raise box["e"]
            pending = self._start_prefetch(weekly[gi + 1]) if gi + 1 < n_groups else None  # overlaps this chunk's batches
            parts = self._chunk_samples(weekly[gi], loaded=box["v"])
            if parts is not None:
                parts = list(parts)
                if carry is not None:
                    parts = [torch.cat([c, p], 0) for c, p in zip(carry, parts)]
                    carry = None
                n = parts[0].shape[0]
                full = (n // self.batch_size) * self.batch_size
                for s in range(0, full, self.batch_size):
                    yield tuple(p[s:s + self.batch_size] for p in parts)
                if full < n:
                    carry = [p[full:].contiguous() for p in parts]
            if gi % 10 == 0 and self.rank == 0:
                logger.info(f"Dataloader iterated over [{gi + 1}/{n_groups}] chunks")
        if carry is not None and carry[0].shape[0] > 0:
            yield tuple(carry)  # last partial batch (the reference DataLoader has drop_last=False)


def chunk_ids_for(split: str, world_size: int = 1, rank: int = 0) -> List[int]:
    """Chunk ids of a split after the reference's per-rank partition (:311-341)."""
    if DRY_RUN:
        train_ids, val_ids = DRY_RUN_TRAIN_CHUNK_IDS, VALIDATION_CHUNK_IDS[:4]
    else:
        train_ids, val_ids = set(range(NUM_DATASET_PARTS)).difference(VALIDATION_CHUNK_IDS), VALIDATION_CHUNK_IDS
    ids = list(train_ids if split.lower() == "train" else val_ids)
    if world_size > 1:
        per_rank = len(ids) // world_size
        ids = ids[: per_rank * world_size][rank * per_rank:(rank + 1) * per_rank]
    return ids


def streaming_dataloader(batch_size, split="train", shuffle=False, num_input_features=None, num_output_features=None,
                         masking_function: Optional[str] = None, masking_prob: float = 0.15, n_masked_features: int = 1,
                         world_size: int = 1, rank: int = 0):
    ids = chunk_ids_for(split, world_size, rank)
    monthly, weekly, daily = list(ids), list(ids), list(ids)
    if shuffle:  # two draws from Python's RNG, like the reference (:349-351)
        random.shuffle(weekly)
        random.shuffle(daily)
    base = DATA_DIR + "nasa_power/processed/"
    paths: List[str] = []
    for m, w, d in zip(monthly, weekly, daily):
        paths += [base + f"weather_dataset_monthly_{m}.pt", base + f"weather_dataset_weekly_{w}.pt",
                  base + f"weather_dataset_daily_{d}.pt"]
    dataset = StreamingDataset(paths, num_input_features=num_input_features, num_output_features=num_output_features,
                               shuffle=shuffle, masking_function=masking_function, masking_prob=masking_prob,
                               n_masked_features=n_masked_features, rank=rank, batch_size=batch_size)
    # batch_size=None: the dataset already yields whole on-device batches
    return torch.utils.data.DataLoader(dataset, batch_size=None, pin_memory=False, num_workers=0)
