"""``BaseRunner.evaluate_method`` / ``evaluate_intents`` replacements (reference
IntEL/src/helpers/BaseRunner.py:57-150) on the segmented top-k / NDCG@k kernels.

``evaluate_method`` keeps the reference signature (lists of per-session numpy rows, numpy counts);
``evaluate_batches`` is the device-resident form used by the benchmark and the data-parallel runner:
padded [N, L] tensors in, one dict of python floats out, no per-session host work.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib

_BEHAVIORS = ("pay", "fav", "click")


def _topk_array(topk: Sequence[int]):
    arr = (C.c_int32 * len(topk))(*[int(k) for k in topk])
    return arr


def ndcg_sums(pred: torch.Tensor, ranking: torch.Tensor, session_len: torch.Tensor, pay: torch.Tensor,
              fav: torch.Tensor, click: torch.Tensor, max_len: int, topk: Sequence[int]):
    """Device sums for one shard of sessions -> (sums f64 [n_topk*7], counts f64 [4]); additive across shards."""
    lib = _lib.load()
    dev = pred.device
    N, ld = pred.shape
    sums = torch.zeros(len(topk) * 7, dtype=torch.float64, device=dev)
    counts = torch.zeros(4, dtype=torch.float64, device=dev)
    if N == 0:
        return sums, counts
    ws = torch.empty(lib.intel_ndcg_workspace_bytes(N, len(topk)), dtype=torch.uint8, device=dev)
    _lib.check(lib.intel_ndcg_topk(N, ld, _lib.ptr(pred.contiguous(), torch.float32),
                                   _lib.ptr(ranking.contiguous(), torch.int64),
                                   _lib.ptr(session_len.contiguous(), torch.int64),
                                   _lib.ptr(pay.contiguous(), torch.int64), _lib.ptr(fav.contiguous(), torch.int64),
                                   _lib.ptr(click.contiguous(), torch.int64), int(max_len), _topk_array(topk),
                                   len(topk), _lib.ptr(sums), _lib.ptr(counts), _lib.ptr(ws), ws.numel(),
                                   _lib.stream_ptr(dev)))
    return sums, counts


def metrics_from_sums(sums: np.ndarray, counts: np.ndarray, topk: Sequence[int], metrics: Sequence[str]) -> Dict[str, float]:
    """Means with the reference's key names ('{behavior}_{metric}@{k}', 'NDCG@{k}'; BaseRunner.py:104,126)."""
    res: Dict[str, float] = {}
    with np.errstate(invalid="ignore", divide="ignore"):
        for x, beh in enumerate(_BEHAVIORS):
            for t, k in enumerate(topk):
                for metric in metrics:
                    if metric == 'HR':
                        res[f'{beh}_HR@{k}'] = float(sums[t * 7 + 1 + 2 * x] / counts[1 + x])
                    elif metric == 'NDCG':
                        if k == 1:
                            continue        # NDCG@1 is the same as HR@1 (BaseRunner.py:109-110)
                        res[f'{beh}_NDCG@{k}'] = float(sums[t * 7 + 2 + 2 * x] / counts[1 + x])
                    else:
                        raise ValueError('Undefined evaluation metric: {}.'.format(metric))
        for t, k in enumerate(topk):
            res['NDCG@%d' % k] = float(sums[t * 7] / counts[0])
    return res


def evaluate_batches(pred: torch.Tensor, ranking: torch.Tensor, pos_nums: Dict[str, torch.Tensor],
                     session_len: torch.Tensor, topk: Sequence[int], metrics: Sequence[str], max_len: int = 0) -> Dict[str, float]:
    if max_len <= 0:
        max_len = max(int(session_len.max().item()), max(topk))
    sums, counts = ndcg_sums(pred, ranking, session_len, pos_nums['c_paynum_i'], pos_nums['c_favnum_i'],
                             pos_nums['c_clicknum_i'], max_len, topk)
    return metrics_from_sums(sums.cpu().numpy(), counts.cpu().numpy(), topk, metrics)


def evaluate_method(prediction_scores, ranking_lists, pos_nums, topk, metrics, session_len, show_num=False,
                    device="cuda") -> Dict[str, float]:
    """Reference signature (BaseRunner.py:57): ragged per-session rows are packed once into padded device
    tensors, everything else happens in intel_ndcg_topk."""
    n = len(prediction_scores)
    session_len = np.asarray(session_len)[:n].astype(np.int64)
    pos = {k: np.asarray(v)[:n].astype(np.int64) for k, v in pos_nums.items()}
    ld = max(len(r) for r in prediction_scores)
    pred = np.zeros((n, ld), dtype=np.float32)
    rank = np.zeros((n, ld), dtype=np.int64)
    for i in range(n):
        m = len(prediction_scores[i])
        pred[i, :m] = prediction_scores[i]
        rank[i, :m] = ranking_lists[i][:m]
    # rows shorter than their session_len behave as in the reference (BaseRunner.py:66-75): min(len, row)
    row_len = np.array([len(r) for r in prediction_scores], dtype=np.int64)
    eff_len = np.minimum(session_len, row_len)
    max_len = int(max(session_len.max(), max(topk)))
    t = lambda a: torch.from_numpy(a).to(device)
    sums, counts = ndcg_sums(t(pred), t(rank), t(eff_len), t(pos['c_paynum_i']), t(pos['c_favnum_i']),
                             t(pos['c_clicknum_i']), max_len, topk)
    return metrics_from_sums(sums.cpu().numpy(), counts.cpu().numpy(), topk, metrics)


def evaluate_intents(true_intents, predict_intents, topk=(1, 5, 10, 30), device="cuda") -> Dict[str, float]:
    """BaseRunner.evaluate_intents (BaseRunner.py:133-150)."""
    lib = _lib.load()
    ti = true_intents if torch.is_tensor(true_intents) else torch.from_numpy(np.asarray(true_intents, dtype=np.float64))
    pi = predict_intents if torch.is_tensor(predict_intents) else torch.from_numpy(np.asarray(predict_intents, dtype=np.float32))
    ti, pi = ti.to(device).double().contiguous(), pi.to(device).float().contiguous()
    N, I = pi.shape
    for k in topk:
        # the reference multiplies true_sort[:, :k] ([N, min(k, I)]) by discounts[:k] ([min(k, 40)]) (BaseRunner.py:142-144):
        # a cut-off beyond the number of intent classes is a numpy broadcast error there, so it is an error here too
        a, b = min(int(k), I), min(int(k), 40)
        if a != b and a != 1 and b != 1:
            raise ValueError(f"operands could not be broadcast together with shapes ({N},{a}) ({b},) "
                             f"(Int-NDCG@{k} needs k <= intent_num = {I} and k <= 40, as in BaseRunner.evaluate_intents)")
    sums = torch.zeros(len(topk) * 2, dtype=torch.float64, device=pi.device)
    ws = torch.empty(lib.intel_intent_topk_workspace_bytes(N, len(topk)), dtype=torch.uint8, device=pi.device)
    _lib.check(lib.intel_intent_topk(N, I, _lib.ptr(ti), _lib.ptr(pi), _topk_array(topk), len(topk), _lib.ptr(sums),
                                     _lib.ptr(ws), ws.numel(), _lib.stream_ptr(pi.device)))
    s = sums.cpu().numpy()
    res: Dict[str, float] = {}
    for t, k in enumerate(topk):
        res['Int-NDCG@%d' % k] = float(s[2 * t] / N)
        res['Int-HR@%d' % k] = float(s[2 * t + 1] / N)
    return res
