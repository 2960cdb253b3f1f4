"""HBM-resident batching of the reference's training examples.

The reference keeps every clip in host RAM (`SingleSpeakerDataset`, tts/dataloader.py:18-90: `code = npy / 1023`, phoneme id
list, length) and builds each batch in Python (`TTS_SingleSpkr_Collate_Fn.__call__`, :145-188):

    code            = Normalize(0.5, 0.5)(FloatTensor(codes / 1023))            [B, 8, T] fp32 in [-1, 1]
    cmu_sequence_id = ids padded with 0 / truncated to max_seq_length           [B, max_seq_length] int32
    attention_mask  = 1 where an id is present                                  [B, max_seq_length] int32

`GpuBatcher` holds the codes of the whole set as int16 [N, 8, T] and the padded id / mask tables on the device once; a batch
is an index gather plus the `pt_codes_affine` kernel -- no host work, no H2D copy per step (LJSpeech: 13 100 clips x 8 x 900
int16 = 189 MB).  The text front-end (cleaners, CMUdict lookup, `intersperse`) is CPU string work outside the tensor path:
callers pass phoneme id lists, or a `tokenizer(text) -> list[int]` to `from_tar`, which reads the reference's on-disk format
(`<name>.npy` int64 [1, 8, T] or [8, T], `<name>.len.txt`, `<name>.txt` / `<name>.normalized.txt`; generate_code.py:61-84).
"""
from __future__ import annotations

import io
import tarfile
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops


class GpuBatcher:
    def __init__(self, codes: Sequence[np.ndarray], cmu_sequences: Sequence[Sequence[int]], max_seq_length: int,
                 code_lengths: Optional[Sequence[float]] = None, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ops._lib.PtError("GpuBatcher keeps the dataset in HBM; there is no CPU path")
        if len(codes) != len(cmu_sequences) or len(codes) == 0:
            raise ValueError("codes and cmu_sequences must be non-empty and of equal length")
        arrs = [np.asarray(c).reshape(-1, np.asarray(c).shape[-1]) for c in codes]        # [1, 8, T] or [8, T] -> [8, T]
        q, t = arrs[0].shape
        if any(a.shape != (q, t) for a in arrs):
            raise ValueError("all clips must share one [Q, T] shape (generate_code.py pads every clip to the same duration)")
        stacked = np.stack(arrs)
        if stacked.min() < 0 or stacked.max() > 1023:
            raise ValueError("codes must lie in [0, 1023]")
        self.codes = torch.from_numpy(stacked.astype(np.int16)).to(dev)                  # [N, Q, T] int16
        n = len(arrs)
        ids = torch.zeros(n, max_seq_length, dtype=torch.int32)
        mask = torch.zeros(n, max_seq_length, dtype=torch.int32)
        for i, seq in enumerate(cmu_sequences):                                          # _collate_batch_helpler, dataloader.py:123-137
            k = min(len(seq), max_seq_length)
            ids[i, :k] = torch.tensor(list(seq[:k]), dtype=torch.int32)
            mask[i, :k] = 1
        self.ids, self.mask = ids.to(dev), mask.to(dev)
        self.code_lengths = list(code_lengths) if code_lengths is not None else [float(t)] * n
        self.max_seq_length = max_seq_length

    def __len__(self) -> int:
        return self.codes.shape[0]

    def batch(self, indices) -> Dict[str, torch.Tensor]:
        """The tensors of TTS_SingleSpkr_Collate_Fn's dict for the clips `indices` (a device or host index tensor / list)."""
        idx = torch.as_tensor(indices, device=self.codes.device, dtype=torch.long)
        codes = self.codes.index_select(0, idx).to(torch.int64)
        return {"code": ops.codes_affine(codes), "cmu_sequence_id": self.ids.index_select(0, idx),
                "attention_mask": self.mask.index_select(0, idx), "code_length": [self.code_lengths[int(i)] for i in idx.tolist()]}

    def epoch(self, batch_size: int, shuffle: bool = False, generator: Optional[torch.Generator] = None):
        """Batches of one pass over the set (DataLoader(dataset, batch_size, shuffle) semantics, last batch kept)."""
        n = len(self)
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        for i in range(0, n, batch_size):
            yield self.batch(order[i:i + batch_size])

    @classmethod
    def from_tar(cls, path: str, tokenizer: Callable[[str], List[int]], max_seq_length: int, device="cuda") -> "GpuBatcher":
        """Read the reference's processed tar (generate_code.py output).  `tokenizer` maps the (normalised, when present) text of a
        clip to its phoneme id list -- in the reference: intersperse(text_to_sequence(text, ["english_cleaners"], cmu_dict), len(symbols))."""
        codes, seqs, lens = [], [], []
        with tarfile.open(path, "r") as tf:
            names = {m.name for m in tf.getmembers()}
            for name in sorted(n for n in names if n.endswith(".npy")):
                codes.append(np.load(io.BytesIO(tf.extractfile(name).read())))
                norm = name.replace(".npy", ".normalized.txt")
                text = tf.extractfile(norm if norm in names else name.replace(".npy", ".txt")).read().decode()
                seqs.append(tokenizer(text))
                lens.append(float(tf.extractfile(name.replace(".npy", ".len.txt")).read().decode()))
        return cls(codes, seqs, max_seq_length, code_lengths=lens, device=device)
