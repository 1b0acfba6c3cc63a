"""Packed feature store + sharded loader -- SURVEY.md 8(f) row 3 (the data format either side of the path).

The reference keeps gallery features in Python structures: a ``dict`` name -> feature tensor rebuilt on every
evaluation (run/test/test_fiq.py:88), ``itemgetter`` + ``torch.stack`` per batch (:104-107) and one ``torch.load``
per image for the 13 patch features (dataloader/fashioniq.py:69-70,97-98).  This module defines the format the CUDA
path consumes directly:

    <dir>/meta.json        {"rows": N, "dim": D, "patches": P | 0, "dtype": "bf16", "version": 1}
    <dir>/global.bf16      N x D  row-major little-endian bfloat16 (unit-norm fused gallery features = the B operand
                           of ern_sim_topk; row i has global id i)
    <dir>/local.bf16       N x P x D bfloat16 patch features (optional; input of VisualSR)
    <dir>/names.txt        one name per row (ids are row numbers; repeated names = Fashion200k captions)

Rows are addressed by integer id everywhere; names are only factorised once on the host.  ``load_shard`` streams the
row block of one rank (``sharded.shard_bounds``) through pinned staging buffers straight into device memory, so a
100M x 640 gallery (128 GB) never has to fit in host RAM.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .sharded import shard_bounds

VERSION = 1


def _as_bf16_bits(t: torch.Tensor) -> np.ndarray:
    """Tensor -> uint16 array holding the bfloat16 bit patterns (round-to-nearest-even from fp32)."""
    return t.detach().to("cpu").to(torch.bfloat16).contiguous().view(torch.int16).numpy().view(np.uint16)


class FeatureStore:
    def __init__(self, path: str):
        self.path = path
        with open(os.path.join(path, "meta.json")) as f:
            self.meta = json.load(f)
        if self.meta.get("version") != VERSION or self.meta.get("dtype") != "bf16":
            raise ValueError(f"unsupported feature store {self.meta}")
        self.rows, self.dim, self.patches = int(self.meta["rows"]), int(self.meta["dim"]), int(self.meta["patches"])
        def mapped(name, shape):
            if self.rows == 0:                                     # an empty file cannot be memory-mapped
                return np.zeros(shape, dtype=np.uint16)
            return np.memmap(os.path.join(path, name), dtype=np.uint16, mode="r", shape=shape)

        self._global = mapped("global.bf16", (self.rows, self.dim))
        self._local = mapped("local.bf16", (self.rows, self.patches, self.dim)) if self.patches else None
        self._names: Optional[List[str]] = None

    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def save(path: str, features: torch.Tensor, names: Optional[Sequence[str]] = None,
             local_features: Optional[torch.Tensor] = None, chunk_rows: int = 1 << 18) -> "FeatureStore":
        os.makedirs(path, exist_ok=True)
        rows, dim = features.shape
        patches = 0 if local_features is None else int(local_features.shape[1])
        if names is not None and len(names) != rows:
            raise ValueError("names must have one entry per row")
        with open(os.path.join(path, "global.bf16"), "wb") as f:
            for s in range(0, rows, chunk_rows):
                f.write(_as_bf16_bits(features[s:s + chunk_rows]).tobytes())
        if local_features is not None:
            with open(os.path.join(path, "local.bf16"), "wb") as f:
                for s in range(0, rows, max(1, chunk_rows // max(patches, 1))):
                    f.write(_as_bf16_bits(local_features[s:s + max(1, chunk_rows // max(patches, 1))]).tobytes())
        with open(os.path.join(path, "names.txt"), "w") as f:
            for i in range(rows):
                f.write((names[i] if names is not None else str(i)).replace("\n", " ") + "\n")
        with open(os.path.join(path, "meta.json"), "w") as f:
            json.dump({"rows": rows, "dim": dim, "patches": patches, "dtype": "bf16", "version": VERSION}, f)
        return FeatureStore(path)

    @staticmethod
    def import_patch_dir(path: str, patch_dir: str, names: Sequence[str], features: torch.Tensor,
                         pattern: str = "{}.pth", workers: int = 8, chunk_rows: int = 4096) -> "FeatureStore":
        """Pack the reference's on-disk patch features -- one ``torch.save``d ``[P, D]`` tensor per image at
        ``<patch_dir>/<name>.pth`` (dataloader/fashioniq.py:69-70,97-98: ``fashion-iq/fashion_local13/``;
        dataloader/cirr.py:55-56,85-86: ``cirr_dataset/cirr_local_13/``) -- together with the ``[N, D]`` global features
        of the same images into one store.  Files are read by a thread pool and written in row chunks, so the
        ``[N, P, D]`` tensor never has to exist in host memory."""
        from concurrent.futures import ThreadPoolExecutor

        rows, dim = features.shape
        if len(names) != rows:
            raise ValueError("names must have one entry per row")
        os.makedirs(path, exist_ok=True)

        def read(name: str) -> torch.Tensor:
            t = torch.load(os.path.join(patch_dir, pattern.format(name)), map_location="cpu")
            t = t.reshape(-1, t.shape[-1]).float()
            if t.shape[1] != dim:
                raise ValueError(f"{name}: patch feature dim {t.shape[1]} != {dim}")
            return t

        patches = int(read(names[0]).shape[0]) if rows else 0
        with ThreadPoolExecutor(max_workers=max(1, workers)) as pool, \
                open(os.path.join(path, "local.bf16"), "wb") as f:
            for s in range(0, rows, chunk_rows):
                block = list(pool.map(read, names[s:s + chunk_rows]))
                for nm, t in zip(names[s:s + chunk_rows], block):
                    if t.shape[0] != patches:
                        raise ValueError(f"{nm}: {t.shape[0]} patches, expected {patches}")
                f.write(_as_bf16_bits(torch.stack(block)).tobytes())
        with open(os.path.join(path, "global.bf16"), "wb") as f:
            for s in range(0, rows, 1 << 18):
                f.write(_as_bf16_bits(features[s:s + (1 << 18)]).tobytes())
        with open(os.path.join(path, "names.txt"), "w") as f:
            for nm in names:
                f.write(str(nm).replace("\n", " ") + "\n")
        with open(os.path.join(path, "meta.json"), "w") as f:
            json.dump({"rows": rows, "dim": dim, "patches": patches, "dtype": "bf16", "version": VERSION}, f)
        return FeatureStore(path)

    # ------------------------------------------------------------------------------------------------
    @property
    def names(self) -> List[str]:
        if self._names is None:
            with open(os.path.join(self.path, "names.txt")) as f:
                self._names = [line.rstrip("\n") for line in f]
        return self._names

    def name_to_row(self) -> Dict[str, int]:
        """name -> row id; a repeated name keeps its LAST row, as ``dict(zip(index_names, index_features))``
        does in the reference (run/test/test_fiq.py:88)."""
        return {nm: i for i, nm in enumerate(self.names)}

    def _to_device(self, arr: np.memmap, begin: int, end: int, device, chunk_rows: int) -> torch.Tensor:
        shape = (end - begin,) + tuple(arr.shape[1:])
        out = torch.empty(shape, dtype=torch.bfloat16, device=device)
        if end == begin:
            return out
        pin = torch.cuda.is_available() and torch.device(device).type == "cuda"
        stages = [torch.empty((chunk_rows,) + tuple(arr.shape[1:]), dtype=torch.int16, pin_memory=pin) for _ in range(2)]
        events = [None, None]
        # the copies run on the DESTINATION device's current stream: record / synchronize on that stream, not on the
        # stream of whatever device happens to be current in this process
        stream = torch.cuda.current_stream(out.device) if pin else None
        for n, s in enumerate(range(begin, end, chunk_rows)):
            e = min(s + chunk_rows, end)
            st = stages[n & 1]
            if events[n & 1] is not None:
                events[n & 1].synchronize()                      # staging buffer free again
            st[:e - s].numpy().view(np.uint16)[...] = arr[s:e]
            if pin:
                with torch.cuda.device(out.device), torch.cuda.stream(stream):
                    out[s - begin:e - begin].view(torch.int16).copy_(st[:e - s], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(stream)
                events[n & 1] = ev
            else:
                out[s - begin:e - begin].view(torch.int16).copy_(st[:e - s])
        if pin:
            stream.synchronize()
        return out

    def load_shard(self, rank: int = 0, world_size: int = 1, device="cuda", chunk_rows: int = 1 << 16
                   ) -> Tuple[torch.Tensor, int]:
        """bf16 gallery rows owned by ``rank`` on ``device`` and the global id of its first row."""
        begin, end = shard_bounds(self.rows, world_size, rank)
        return self._to_device(self._global, begin, end, device, chunk_rows), begin

    def load_local(self, rows: Sequence[int], device="cuda") -> torch.Tensor:
        """Patch features [len(rows), P, D] (bf16) of the given row ids -- integer-id gather instead of one
        ``torch.load`` per image."""
        if self._local is None:
            raise ValueError("store has no patch features")
        idx = np.asarray(rows, dtype=np.int64)
        host = torch.from_numpy(np.ascontiguousarray(self._local[idx]).view(np.int16))
        return host.to(device).view(torch.bfloat16)

    def gather(self, rows: Sequence[int], device="cuda") -> torch.Tensor:
        """Global features [len(rows), D] (bf16) of the given row ids (replaces itemgetter + torch.stack)."""
        idx = np.asarray(rows, dtype=np.int64)
        host = torch.from_numpy(np.ascontiguousarray(self._global[idx]).view(np.int16))
        return host.to(device).view(torch.bfloat16)


def stream_topk(store: "FeatureStore", queries: torch.Tensor, k: int, *, chunk_rows: int = 1 << 22, rank: int = 0,
                world_size: int = 1, rank_by: int = 0):
    """Exact top-k over a gallery that does not have to fit in device memory: the rank's row block is streamed from
    the store in ``chunk_rows`` pieces (pinned staging -> device), each piece is scored with ``ops.sim_topk`` using
    its global ``id_offset`` and folded into the running best ``[Q,k]`` keys with ``ops.topk_merge`` -- the same
    (value, id) wire keys and merge kernel the multi-GPU exchange uses, so the result is identical to scoring the
    whole block at once.  Returns ``(values, ids, keys)``."""
    from . import ops
    begin, end = shard_bounds(store.rows, world_size, rank)
    dev = queries.device
    best = None
    for s in range(begin, max(end, begin + 1), chunk_rows):
        e = min(s + chunk_rows, end)
        piece = store._to_device(store._global, s, e, dev, min(chunk_rows, 1 << 16))
        _, _, keys, _ = ops.sim_topk(queries, piece, k, id_offset=s, rank_by=rank_by, want_keys=True)
        best = keys if best is None else ops.topk_merge(torch.stack((best, keys)), k)[2]
    return ops.topk_merge(best.unsqueeze(0), k)
