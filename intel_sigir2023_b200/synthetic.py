"""Synthetic Tmall-schema batches in the reference's collate layout.

The batch dict produced here has the keys, dtypes and padding conventions of
``BaseModel.Dataset.collate_batch`` (reference IntEL/src/models/BaseModel.py:121-142)
fed by the feed-dict builders (BaseModel.py:158-197, GeneralSeq.py:35-54,
IntEL.py:220-239): ragged fields right-padded with 0, ``scores`` float64 [B,L,K]
per-session min-max normalised (BaseModel.py:173), ``ranking`` 3/2/1 pay/fav/click,
0 true-negative, -1 unlabeled (BaseModel.py:177-185), dense float64 ``his_intents``
and one-hot ``his_item_int``.

Everything is generated with torch on the requested device from a seeded
``torch.Generator`` so large batches (B=4096, I=1071, H=20 -> 1.4 GB of dense
history intents) are cheap to build; there are no datasets on the GPU box.
Statistics follow SURVEY.md section 8(d): raw base scores ~N(mu_k, sigma_k) with
the toy-set moments, pay~Poisson(0.26), fav~Poisson(0.94), click~1+Poisson(2.8),
positives capped at n/2, intent vectors with 1-8 non-zeros.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch

# toy-set moments of the raw base scores: pCTR, pCVR, pFVR, extra
_SCORE_MU = (10.1, 1.7, 0.5, 0.0)
_SCORE_SD = (6.1, 3.1, 3.0, 3.0)


@dataclass
class CorpusSpec:
    """Sizes of the synthetic corpus (what the reference's reader would report)."""
    n_item: int = 1000        # max_iid; item ids 1..n_item, table rows n_item+1
    n_class: int = 357        # product(itemfnum) rows of item_embeddings
    n_user: int = 500         # max_uid; table rows n_user+1
    n_ctx: int = 931          # product(contextfnum) rows of context_embeddings
    model_num: int = 3        # K basic lists
    intent_num: int = 1071    # I
    history_max: int = 20     # H cap (--history_max)

    @property
    def item_rows(self) -> int:
        return self.n_item + 1

    @property
    def user_rows(self) -> int:
        return self.n_user + 1


@dataclass
class BatchSpec:
    batch_size: int = 8
    max_len: int = 12          # L: every batch is padded to the longest session
    min_len: int = 12          # sessions draw n_b ~ U{min_len..max_len}
    max_nnz: int = 8           # non-zeros per intent vector (1..max_nnz)
    unlabeled_frac: float = 0.1
    phase: str = "train"
    force_full: bool = True    # make at least one session have n_b == max_len


def _poisson(rate: float, shape, gen: torch.Generator, device) -> torch.Tensor:
    return torch.poisson(torch.full(shape, rate, device=device), generator=gen).long()


def _intent_pairs(rows: int, I: int, max_nnz: int, gen, device):
    """(idx int64 [rows, nnz], val float64 [rows, nnz]): 1..max_nnz positive entries per row summing to 1
    (an index may repeat inside a row; repeated entries add up)."""
    max_nnz = max(1, min(max_nnz, I))
    nnz = torch.randint(1, max_nnz + 1, (rows,), generator=gen, device=device)
    idx = torch.randint(0, I, (rows, max_nnz), generator=gen, device=device)
    val = torch.rand(rows, max_nnz, generator=gen, device=device, dtype=torch.float64) + 0.05
    keep = torch.arange(max_nnz, device=device)[None, :] < nnz[:, None]
    val = val * keep
    val = val / val.sum(dim=1, keepdim=True).clamp_min(1e-30)
    return idx, val


def _densify(idx: torch.Tensor, val: torch.Tensor, I: int) -> torch.Tensor:
    out = torch.zeros(idx.shape[0], I, dtype=torch.float64, device=idx.device)
    if idx.shape[0]:
        out.scatter_add_(1, idx, val)
    return out


def _intent_rows(rows: int, I: int, max_nnz: int, gen, device) -> torch.Tensor:
    """Dense float64 [rows, I] vectors with 1..max_nnz positive entries summing to 1."""
    idx, val = _intent_pairs(rows, I, max_nnz, gen, device)
    return _densify(idx, val, I)


def make_batch(corpus: CorpusSpec, spec: BatchSpec, seed: int = 0,
               device: str | torch.device = "cpu", layout: str = "dense") -> Dict[str, object]:
    """One collated batch; same keys as the reference's DataLoader would yield.

    layout: "dense" = the reference API (float64 [B,H,I] ``his_intents`` / one-hot ``his_item_int``);
    "compact" = the opt-in index form of the same two tensors (``his_intents_idx`` int32 [B,H,nz] +
    ``his_intents_val`` float32, ``his_item_int_idx`` int32 [B,H2,1] + ``his_item_int_val``), which is what a
    device-side batch builder would emit; "both" = both, from the same draws."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    B, L, K, I = spec.batch_size, spec.max_len, corpus.model_num, corpus.intent_num
    ar = torch.arange(L, device=dev)

    n = torch.randint(spec.min_len, spec.max_len + 1, (B,), generator=gen, device=dev)
    if spec.force_full:
        n[0] = L
    valid = ar[None, :] < n[:, None]

    # items: Zipf-ish popularity via a squared uniform, ids 1..n_item; pad id 0
    u = torch.rand(B, L, generator=gen, device=dev)
    i_id = (1 + (u * u * corpus.n_item).long()).clamp_(1, corpus.n_item) * valid
    # a fixed id -> class map (class ids 1..n_class-1; 0 is the pad class)
    i_class = (1 + (i_id * 2654435761 % max(corpus.n_class - 1, 1))) * valid

    # raw base scores then per-session min-max over the real items only
    mu = torch.tensor([_SCORE_MU[k % 4] for k in range(K)], dtype=torch.float64, device=dev)
    sd = torch.tensor([_SCORE_SD[k % 4] for k in range(K)], dtype=torch.float64, device=dev)
    raw = torch.randn(B, L, K, generator=gen, device=dev, dtype=torch.float64) * sd + mu
    big = torch.finfo(torch.float64).max
    lo = raw.masked_fill(~valid[:, :, None], big).amin(dim=1, keepdim=True)
    hi = raw.masked_fill(~valid[:, :, None], -big).amax(dim=1, keepdim=True)
    scores = ((raw - lo) / (hi - lo + 1e-6)) * valid[:, :, None]

    # labels: counts per behaviour, positives capped at n/2, list order shuffled
    pay = _poisson(0.26, (B,), gen, dev)
    fav = _poisson(0.94, (B,), gen, dev)
    clk = 1 + _poisson(2.8, (B,), gen, dev)
    cap = torch.clamp(n // 2, min=1)
    pay = torch.minimum(pay, cap)
    fav = torch.minimum(fav, cap - pay)
    clk = torch.minimum(clk, cap - pay - fav)
    pos = pay + fav + clk
    unl = ((n - pos).double() * spec.unlabeled_frac).long()
    neg = n - pos - unl
    c1, c2, c3, c4 = pay, pay + fav, pos, pos + neg
    slot = ar[None, :]
    ordered = torch.where(slot < c1[:, None], 3,
              torch.where(slot < c2[:, None], 2,
              torch.where(slot < c3[:, None], 1,
              torch.where(slot < c4[:, None], 0, -1))))
    # random permutation of the first n_b slots (BaseModel.py:194-196)
    key = torch.rand(B, L, generator=gen, device=dev).masked_fill(~valid, 2.0)
    perm = key.argsort(dim=1)
    ranking = torch.gather(ordered, 1, perm) * valid  # pad ranking = 0

    u_id = torch.randint(1, corpus.n_user + 1, (B,), generator=gen, device=dev)
    ctx = torch.randint(1, corpus.n_ctx, (B,), generator=gen, device=dev)
    c_id = torch.arange(B, device=dev) + 1 + 1000 * (int(seed) % 1000)

    # session-history of (context, intent vector); item-history of (item, one-hot intent)
    Hmax = corpus.history_max
    h_len = torch.randint(1, Hmax + 1, (B,), generator=gen, device=dev)
    hi_len = torch.randint(1, Hmax + 1, (B,), generator=gen, device=dev)
    H, H2 = int(h_len.max()), int(hi_len.max())
    hv = torch.arange(H, device=dev)[None, :] < h_len[:, None]
    hv2 = torch.arange(H2, device=dev)[None, :] < hi_len[:, None]
    his_ctx = torch.randint(1, corpus.n_ctx, (B, H), generator=gen, device=dev) * hv
    hidx, hval = _intent_pairs(B * H, I, spec.max_nnz, gen, dev)
    hval = hval * hv.reshape(-1, 1)
    uh = torch.rand(B, H2, generator=gen, device=dev)
    his_item = (1 + (uh * uh * corpus.n_item).long()).clamp_(1, corpus.n_item) * hv2
    hot = torch.randint(0, I, (B, H2), generator=gen, device=dev)
    intents = _intent_rows(B, I, spec.max_nnz, gen, dev)
    extra: Dict[str, object] = {}
    if layout in ("dense", "both"):
        his_item_int = torch.zeros(B, H2, I, dtype=torch.float64, device=dev)
        his_item_int.scatter_(2, hot[:, :, None], hv2[:, :, None].double())
        extra["his_intents"] = _densify(hidx, hval, I).view(B, H, I)
        extra["his_item_int"] = his_item_int
    if layout in ("compact", "both"):
        extra["his_intents_idx"] = hidx.view(B, H, -1).to(torch.int32)
        extra["his_intents_val"] = hval.view(B, H, -1).to(torch.float32)
        extra["his_item_int_idx"] = hot.view(B, H2, 1).to(torch.int32)
        extra["his_item_int_val"] = hv2.view(B, H2, 1).to(torch.float32)
    if layout not in ("dense", "compact", "both"):
        raise ValueError(layout)

    out = {
        "u_id_c": u_id, "c_id_c": c_id, "context_mh": ctx,
        "user_mh": torch.zeros(B, dtype=torch.long, device=dev),
        "c_paynum_i": pay, "c_favnum_i": fav, "c_clicknum_i": clk,
        "i_class_c": i_class, "i_id_s": i_id, "session_len": n,
        "intents": intents, "ranking": ranking,
        "his_context_mh": his_ctx,
        "position": h_len.clone(), "history_len": h_len,
        "intentloss_w": torch.full((B, I), 1.0 / I, dtype=torch.float64, device=dev),
        "his_item_id": his_item,
        "history_item_len": hi_len, "scores": scores,
        "batch_size": B, "phase": spec.phase,
    }
    out.update(extra)
    return out


def batch_to(batch: Dict[str, object], device, non_blocking: bool = False) -> Dict[str, object]:
    """Device move of every tensor value (reference utils.batch_to_gpu, utils.py:91-95)."""
    return {k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
            for k, v in batch.items()}


def shard_batch(batch: Dict[str, object], rank: int, world: int) -> Dict[str, object]:
    """Contiguous equal shard of the sessions of one global batch.

    Every rank keeps the global padded widths (L, H) so pad keys take part in the
    unmasked self-attention exactly as in a single-process run."""
    B = int(batch["batch_size"])
    if B % world:
        raise ValueError(f"global batch {B} is not divisible by world size {world}")
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    out = {k: (v[sl] if torch.is_tensor(v) else v) for k, v in batch.items()}
    out["batch_size"] = per
    return out


def eval_set(n_sessions: int, max_len: int, min_len: int, seed: int = 0,
             device: str | torch.device = "cpu") -> Tuple[torch.Tensor, torch.Tensor, Dict[str, torch.Tensor], torch.Tensor]:
    """(scores f32 [N,L], ranking i64 [N,L], pos_nums, session_len) for evaluate_method."""
    corpus = CorpusSpec(n_item=50, n_class=7, n_user=5, n_ctx=5, model_num=1, intent_num=2, history_max=1)
    spec = BatchSpec(batch_size=n_sessions, max_len=max_len, min_len=min_len, max_nnz=1)
    b = make_batch(corpus, spec, seed=seed, device=device)
    gen = torch.Generator(device=torch.device(device))
    gen.manual_seed(int(seed) + 7919)
    valid = torch.arange(max_len, device=device)[None, :] < b["session_len"][:, None]
    pred = (torch.randn(n_sessions, max_len, generator=gen, device=device) * 0.7 + 0.2) * valid
    pos = {k: b[k] for k in ("c_paynum_i", "c_favnum_i", "c_clicknum_i")}
    return pred.float(), b["ranking"], pos, b["session_len"]
