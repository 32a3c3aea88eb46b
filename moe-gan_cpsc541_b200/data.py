"""On-disk formats of the reference, so the B200 trainer is interchangeable with its artefacts.

Dataset: two `.npy` files written by data_processing/data_processing_pipeline.py (:249-252) - images float32
[N,3,64,64] in [-1,1] (:37-42) and CLIP text embeddings float32 [N,512] (:162) - read here memory-mapped and fed
through a double-buffered pinned-memory prefetcher (H2D copies on a side stream overlap the previous step).
Checkpoint: `{'generator': state_dict, 'discriminator': state_dict}` as saved by moegan/sagemaker_train.py:297-301.
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import numpy as np
import torch

from .host_model import load_reference_checkpoint


class NpyPairDataset(torch.utils.data.Dataset):
    def __init__(self, images_npy: str, embeddings_npy: str):
        self.images = np.load(images_npy, mmap_mode="r")
        self.embeddings = np.load(embeddings_npy, mmap_mode="r")
        if self.images.shape[0] != self.embeddings.shape[0]:
            raise ValueError(f"{images_npy}: {self.images.shape[0]} images but {self.embeddings.shape[0]} embeddings")
        if self.images.ndim != 4 or self.images.shape[1] != 3:
            raise ValueError(f"{images_npy}: expected [N,3,H,W], got {self.images.shape}")

    def __len__(self) -> int:
        return int(self.images.shape[0])

    def __getitem__(self, i: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return torch.from_numpy(np.array(self.images[i], dtype=np.float32)), torch.from_numpy(np.array(self.embeddings[i], dtype=np.float32))


class PinnedPrefetcher:
    """Yields (images, embeddings) batches on `device`; rank `rank` of `world` takes every world-th batch (data parallel).

    Two pinned staging buffers alternate: while the consumer trains on batch i, batch i+1 is assembled on the host and
    copied with `non_blocking=True` on a side stream.  drop_last=True (static shapes for CUDA-graph replay).
    """

    def __init__(self, dataset: NpyPairDataset, batch_size: int, device: torch.device, shuffle: bool = True, seed: int = 0,
                 rank: int = 0, world: int = 1):
        self.ds, self.bs, self.device = dataset, batch_size, torch.device(device)
        self.shuffle, self.seed, self.rank, self.world = shuffle, seed, rank, world
        self.epoch = 0
        self.cuda = self.device.type == "cuda"
        n_img = dataset.images.shape[1:]
        n_emb = dataset.embeddings.shape[1:]
        mk = lambda shape: torch.empty((batch_size, *shape), dtype=torch.float32, pin_memory=self.cuda)
        self._stage = [(mk(n_img), mk(n_emb)) for _ in range(2)]
        self._stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        # one event per staging slot, recorded after the slot's H2D copies: the HOST waits on it before it overwrites the
        # pinned buffer (stream ordering alone does not stop the CPU, which can run several steps ahead of the GPU, from
        # rewriting a buffer whose non_blocking copy is still queued)
        self._copied = [torch.cuda.Event() if self.cuda else None for _ in range(2)]
        self._in_flight = [False, False]

    def __len__(self) -> int:
        return (len(self.ds) // self.bs) // self.world

    def _fill(self, slot: int, idx: np.ndarray):
        img, emb = self._stage[slot]
        if self.cuda and self._in_flight[slot]:
            self._copied[slot].synchronize()                   # the previous upload out of this slot has finished
            self._in_flight[slot] = False
        order = np.sort(idx)                                   # sorted reads are kinder to the memory map
        img.copy_(torch.from_numpy(np.ascontiguousarray(self.ds.images[order], dtype=np.float32)))
        emb.copy_(torch.from_numpy(np.ascontiguousarray(self.ds.embeddings[order], dtype=np.float32)))
        if not self.cuda:
            return img.clone(), emb.clone()
        with torch.cuda.stream(self._stream):
            out = img.to(self.device, non_blocking=True), emb.to(self.device, non_blocking=True)
            self._copied[slot].record(self._stream)
        self._in_flight[slot] = True
        return out

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        n = len(self.ds)
        perm = np.random.RandomState(self.seed + self.epoch).permutation(n) if self.shuffle else np.arange(n)
        self.epoch += 1
        batches = [perm[i * self.bs:(i + 1) * self.bs] for i in range(n // self.bs)][self.rank::self.world][:len(self)]
        nxt = self._fill(0, batches[0]) if batches else None
        for i in range(len(batches)):
            cur = nxt
            if self.cuda:
                torch.cuda.current_stream(self.device).wait_stream(self._stream)   # batch i has landed
            if i + 1 < len(batches):
                if self.cuda:
                    self._stream.wait_stream(torch.cuda.current_stream(self.device))  # staging slot is free again
                nxt = self._fill((i + 1) & 1, batches[i + 1])
            yield cur


def save_checkpoint(path: str, generator: torch.nn.Module, discriminator: Optional[torch.nn.Module] = None) -> None:
    """Reference format (sagemaker_train.py:297-301).  The discriminator's weight-norm parameters are written under the
    legacy names (weight_g / weight_v) the reference's `torch.nn.utils.weight_norm` modules expect."""
    ckpt = {"generator": {k: v.detach().cpu() for k, v in generator.state_dict().items()}}
    if discriminator is not None:
        sd = {}
        for k, v in discriminator.state_dict().items():
            k = k.replace("parametrizations.weight.original0", "weight_g").replace("parametrizations.weight.original1", "weight_v")
            sd[k] = v.detach().cpu()
        ckpt["discriminator"] = sd
    torch.save(ckpt, path)


def load_checkpoint(path: str, generator: torch.nn.Module, discriminator: Optional[torch.nn.Module] = None) -> None:
    load_reference_checkpoint(generator, discriminator, torch.load(path, map_location="cpu"))
