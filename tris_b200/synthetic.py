"""Synthetic RefCOCOg-shaped batches: img ~ N(0,1) fp32 [B,3,S,S]; word_ids int32 [B,L] = SOT, n~U{3..17}
random tokens, EOT, zero padding; neg_word_ids [B,negs,L].  Same tensor contract as dataset/ReferDataset.py:190-229."""
from __future__ import annotations

import torch

SOT, EOT = 49406, 49407


def _sentences(n, max_len, g):
    ids = torch.zeros((n, max_len), dtype=torch.int32)
    lens = torch.randint(3, min(17, max_len - 2) + 1, (n,), generator=g)
    body = torch.randint(1, SOT, (n, max_len), generator=g, dtype=torch.int32)
    pos = torch.arange(max_len).unsqueeze(0)
    ids = torch.where((pos >= 1) & (pos <= lens.unsqueeze(1)), body, ids)
    ids[:, 0] = SOT
    ids[torch.arange(n), lens + 1] = EOT
    return ids


def synthetic_batch(batch, size=320, max_len=20, negatives=3, seed=1234, pin=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    img = torch.randn((batch, 3, size, size), generator=g, dtype=torch.float32)
    word_ids = _sentences(batch, max_len, g)
    neg = _sentences(batch * negatives, max_len, g).reshape(batch, negatives, max_len) if negatives > 0 else None
    if pin:
        img, word_ids = img.pin_memory(), word_ids.pin_memory()
        neg = neg.pin_memory() if neg is not None else None
    return img, word_ids, neg
