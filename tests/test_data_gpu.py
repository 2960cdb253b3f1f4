"""GPU: the HBM-resident batcher against the oracle's restatement of the reference's dataset + collate (tts/dataloader.py)."""
import io
import os
import tarfile
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _clips(n, T, seed):
    rs = np.random.RandomState(seed)
    codes = [rs.randint(0, 1024, size=(1, 8, T)).astype(np.int64) for _ in range(n)]     # generate_code.py saves [1, 8, T]
    seqs = [rs.randint(1, 150, size=rs.randint(5, 80)).tolist() for _ in range(n)]
    return codes, seqs


def test_batcher_matches_reference_collate_bit_exact(cuda):
    import ref_model
    from prompt_tts_b200.data import GpuBatcher
    codes, seqs = _clips(11, 96, 0)
    gb = GpuBatcher(codes, seqs, max_seq_length=40, device=cuda)       # 40 < some sequence lengths: truncation path
    idx = [7, 0, 3, 10, 3]
    out = gb.batch(idx)
    code, ids, mask = ref_model.collate([codes[i].squeeze(0) for i in idx], [seqs[i] for i in idx], 40)
    assert torch.equal(out["code"].cpu(), code)                        # fp32, bit for bit
    assert torch.equal(out["cmu_sequence_id"].cpu(), ids) and torch.equal(out["attention_mask"].cpu(), mask)
    assert out["code"].dtype == torch.float32 and out["cmu_sequence_id"].dtype == torch.int32
    seen = sum(b["code"].shape[0] for b in gb.epoch(4))
    assert seen == len(gb) == 11


def test_batcher_reads_the_reference_tar_format(cuda):
    from prompt_tts_b200.data import GpuBatcher
    codes, seqs = _clips(5, 64, 1)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "set_processed.tar")
        with tarfile.open(path, "w") as tf:
            def add(name, data):
                ti = tarfile.TarInfo(name)
                ti.size = len(data)
                tf.addfile(ti, io.BytesIO(data))
            for i, c in enumerate(codes):
                buf = io.BytesIO()
                np.save(buf, c)
                add(f"clip{i}.npy", buf.getvalue())
                add(f"clip{i}.len.txt", str(64 / 75).encode())
                add(f"clip{i}.txt", " ".join(map(str, seqs[i])).encode())
                if i % 2 == 0:
                    add(f"clip{i}.normalized.txt", " ".join(map(str, seqs[i][::-1])).encode())   # the normalised text wins when present
        gb = GpuBatcher.from_tar(path, tokenizer=lambda t: [int(x) for x in t.split()], max_seq_length=100, device=cuda)
    assert len(gb) == 5
    b = gb.batch([0, 1])
    assert b["cmu_sequence_id"][0, :len(seqs[0])].tolist() == seqs[0][::-1]
    assert b["cmu_sequence_id"][1, :len(seqs[1])].tolist() == seqs[1]
    assert torch.equal(b["code"].cpu(), ((torch.from_numpy(np.stack([codes[0][0], codes[1][0]])).float() / 1023) - 0.5) / 0.5)
