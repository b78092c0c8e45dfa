"""CPU ORACLE (test infrastructure, NOT a product path) for the IntEL hot path.

A functional restatement of the reference's per-session ensemble scoring, listwise
losses and evaluation, written against a plain ``state_dict`` so that tests can feed
the same weights to the CUDA path.  It deliberately keeps the reference's *literal*
semantics (padded keys are live in self-attention, the [B,1,L]->[B,L,L] broadcast of
the cross-attention mask, [B,L,L(,K)] loss tensors) so the collapsed/fused forms used
by the kernels are checked against the real thing, not against themselves.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module.

The arithmetic lives in third-party PyTorch / numpy (reference pins torch==1.7.1,
numpy==1.19.2 in requirements.txt; this image has torch 2.11 / numpy 2.3).  The
reference has no tests or golden vectors of its own, so parity is
pinned by ``tests/golden/*.npz``: outputs of the unmodified reference modules imported
in the build container by ``oracle/make_golden.py``; ``tests/test_oracle_golden.py``
checks every function below against them.

Reference map (all paths under IntEL/src/):
  predict_intent      models/IntEL/IntEL.py:126-155
  gru_encode          models/GeneralSeq.py:58-78      (nn.GRU equations, torch docs)
  bert_encode         models/GeneralSeq.py:80-106, modules/layers.py:62-88
  mha                 modules/layers.py:31-60
  predict_ensemble    models/IntEL/IntEL.py:158-217
  cross_att           modules/attention.py:54-63, 149-161
  list_loss           loss/Listloss.py:12-43
  bpr_loss            loss/BPRloss.py:12-56
  mse_loss            loss/MSEloss.py:12-30
  intent_loss         loss/BaseIntloss.py:30-67
  total_loss          loss/Int{List,BPR,MSE}loss.py:14-19
  evaluate_method     helpers/BaseRunner.py:57-131
  evaluate_intents    helpers/BaseRunner.py:133-150
  single_sort / borda / random_fusion
                      models/unsupervise/SingleSort.py:23-32, Borda.py:23-30, models/GeneralSeq.py:23-32
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from intel_sigir2023_b200.config import IntelConfig

Tensor = torch.Tensor
State = Dict[str, Tensor]


# --------------------------------------------------------------------------- helpers
def init_state(cfg: IntelConfig, seed: int = 0, dtype=torch.float32) -> State:
    """Random weights with PyTorch-default-like distributions for every state_dict key."""
    g = torch.Generator().manual_seed(seed)
    sd: State = {}
    for key, shape in cfg.param_shapes().items():
        if "embeddings.weight" in key and len(shape) == 2 and key.split(".")[0] in (
                "iid_embeddings", "item_embeddings", "uid_embeddings", "context_embeddings") \
                or key.endswith("p_embeddings.weight"):
            t = torch.randn(shape, generator=g)
        elif "layer_norm" in key:
            t = (1.0 + 0.1 * torch.randn(shape, generator=g)) if key.endswith("weight") \
                else 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] if len(shape) == 2 else max(shape[0], 1)
            if ".rnn." in key:
                fan_in = cfg.gru_hidden
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[key] = t.to(dtype)
    return sd


def _lin(sd: State, name: str, x: Tensor, bias: bool = True) -> Tensor:
    y = x @ sd[name + ".weight"].t()
    if bias and (name + ".bias") in sd:
        y = y + sd[name + ".bias"]
    return y


def _layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def mha(sd: State, name: str, x: Tensor, heads: int, key_mask: Optional[Tensor]) -> Tensor:
    """layers.py:31-60.  x [B,T,d]; key_mask [B,T] bool or None.  No output projection;
    the softmax is shifted by the *global* max of the score tensor (layers.py:57)."""
    B, T, d = x.shape
    dk = d // heads
    def split(t):
        return t.view(B, T, heads, dk).transpose(1, 2)
    q, k, v = (split(_lin(sd, f"{name}.{n}_linear", x)) for n in "qkv")
    s = q @ k.transpose(-1, -2) / dk ** 0.5
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], -math.inf)
    p = (s - s.max()).softmax(dim=-1)
    p = torch.where(torch.isnan(p), torch.zeros_like(p), p)
    return (p @ v).transpose(1, 2).reshape(B, T, d)


# --------------------------------------------------------------------------- encoders
def gru_encode(sd: State, name: str, seq: Tensor, lengths: Tensor, hidden: int) -> Tensor:
    """GeneralSeq.py:58-78: packed single-layer GRU, last hidden state, bias-free out
    projection.  Written as the explicit recurrence (gate order r,z,n)."""
    B, T, _ = seq.shape
    w_ih, w_hh = sd[f"{name}.rnn.weight_ih_l0"], sd[f"{name}.rnn.weight_hh_l0"]
    b_ih, b_hh = sd[f"{name}.rnn.bias_ih_l0"], sd[f"{name}.rnn.bias_hh_l0"]
    h = seq.new_zeros(B, hidden)
    gi_all = seq @ w_ih.t() + b_ih
    for t in range(T):
        gi, gh = gi_all[:, t], h @ w_hh.t() + b_hh
        i_r, i_z, i_n = gi.chunk(3, dim=1)
        h_r, h_z, h_n = gh.chunk(3, dim=1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h_new = (1 - z) * n + z * h
        live = (t < lengths)[:, None]
        h = torch.where(live, h_new, h)
    return h @ sd[f"{name}.out.weight"].t()


def bert_encode(sd: State, name: str, seq: Tensor, lengths: Tensor, layers: int, heads: int) -> Tensor:
    """GeneralSeq.py:80-106 with TransformerLayer (layers.py:62-88): learned positions
    (pad slots use position 0), key-masked post-LN blocks with d_ff = d, state at len-1."""
    B, T, d = seq.shape
    ar = torch.arange(T, device=seq.device)
    valid = ar[None, :] < lengths[:, None]
    x = seq + sd[f"{name}.p_embeddings.weight"][ar[None, :] * valid.long()]
    for l in range(layers):
        blk = f"{name}.transformer_block.{l}"
        ctx = mha(sd, f"{blk}.masked_attn_head", x, heads, valid)
        ctx = _layer_norm(ctx + x, sd[f"{blk}.layer_norm1.weight"], sd[f"{blk}.layer_norm1.bias"])
        out = _lin(sd, f"{blk}.linear2", torch.relu(_lin(sd, f"{blk}.linear1", ctx)))
        x = _layer_norm(out + ctx, sd[f"{blk}.layer_norm2.weight"], sd[f"{blk}.layer_norm2.bias"])
    x = x * valid[:, :, None].to(x.dtype)
    return x[torch.arange(B, device=seq.device), lengths - 1]


def _encode(sd: State, cfg: IntelConfig, name: str, seq: Tensor, lengths: Tensor) -> Tensor:
    if cfg.encoder == "GRU4Rec":
        return gru_encode(sd, name, seq, lengths, cfg.gru_hidden)
    if cfg.encoder == "BERT4Rec":
        return bert_encode(sd, name, seq, lengths, cfg.bert_layers, cfg.bert_heads)
    raise ValueError("Invalid sequence encoder.")


# --------------------------------------------------------------------------- model
def predict_intent(sd: State, cfg: IntelConfig, batch: Dict[str, object]) -> Tensor:
    """IntEL.py:126-155 -> softmax intent distribution [B, I]."""
    dt = sd["pred_layer.weight"].dtype
    ctx_tab, iid_tab = sd["context_embeddings.weight"], sd["iid_embeddings.weight"]
    tok1 = torch.cat([ctx_tab[batch["his_context_mh"]],
                      _lin(sd, "intent_embeddings", batch["his_intents"].to(dt))], dim=-1)
    v1 = _encode(sd, cfg, "encoder", tok1, batch["history_len"])
    tok2 = torch.cat([iid_tab[batch["his_item_id"]],
                      _lin(sd, "intent_embeddings", batch["his_item_int"].to(dt))], dim=-1)
    v2 = _encode(sd, cfg, "item_encoder", tok2, batch["history_item_len"])
    cur = torch.cat([ctx_tab[batch["context_mh"]], sd["uid_embeddings.weight"][batch["u_id_c"]]], dim=-1)
    return _lin(sd, "pred_layer", torch.cat([cur, v2, v1], dim=-1)).softmax(dim=-1)


def _self_att_stack(sd: State, p: str, h: Tensor, cfg: IntelConfig) -> Tensor:
    """IntEL.py:182-197: N iterations that reuse one set of weights; pad slots are live."""
    for _ in range(cfg.num_layers):
        res = h
        h = mha(sd, f"{p}_attn_head", h, cfg.num_heads, None)
        h = _lin(sd, f"{p}_W2", torch.relu(_lin(sd, f"{p}_W1", h)))
        h = _layer_norm(h + res, sd[f"{p}_layer_norm.weight"], sd[f"{p}_layer_norm.bias"])
    return h


def cross_att(sd: State, name: str, intent: Tensor, h: Tensor, valid2: Tensor, scale: float) -> Tensor:
    """attention.py:149-161 + 54-63.  The [B,1,L] attention row is broadcast against the
    [B,L,L] pair mask, so every valid row receives the same pooled vector."""
    q = intent[:, None, :] @ sd[f"{name}.query_layer.weight"].t()          # [B,1,a]
    k = h @ sd[f"{name}.key_layer.weight"].t()
    v = h @ sd[f"{name}.value_layer.weight"].t()
    att = (q @ k.transpose(-1, -2)) * scale                                 # [B,1,L]
    att = att - att.max(dim=-1, keepdim=True)[0]
    att = att.masked_fill(valid2 <= 0, -math.inf)                           # -> [B,L,L]
    w = att.softmax(dim=-1)
    w = torch.where(torch.isnan(w), torch.zeros_like(w), w)
    return w @ v


def predict_ensemble(sd: State, cfg: IntelConfig, batch: Dict[str, object], intent: Tensor) -> Tuple[Tensor, Tensor]:
    """IntEL.py:158-217 -> (weights [B,L,K], ens_score [B,L])."""
    dt = intent.dtype
    items, scores = batch["i_id_s"], batch["scores"].to(dt)
    B, L = items.shape
    valid = torch.arange(L, device=items.device)[None, :] < batch["session_len"][:, None]
    valid2 = valid[:, :, None] * valid[:, None, :]
    h_i = sd["iid_embeddings.weight"][items]
    if "item_embeddings.weight" in sd:
        h_i = torch.cat([h_i, sd["item_embeddings.weight"][batch["i_class_c"]]], dim=2)
    h_u = torch.relu(sd["uid_embeddings.weight"][batch["u_id_c"]])[:, None, :].expand(B, L, -1)
    h_i = _self_att_stack(sd, "i", h_i, cfg)
    h_s = _self_att_stack(sd, "s", _lin(sd, "score_embeddings", scores), cfg)
    if cfg.cross_attention:
        scale = 1.0 / math.sqrt(cfg.cross_attn_qsize)
        x_i = cross_att(sd, "intent_item_attention", intent, h_i, valid2, scale)
        x_s = cross_att(sd, "intent_score_attention", intent, h_s, valid2, scale)
    else:
        def mlp(name):
            t = torch.relu(intent @ sd[f"{name}.0.weight"].t() + sd[f"{name}.0.bias"])
            return (t @ sd[f"{name}.2.weight"].t())[:, None, :]
        x_i = h_i * mlp("intent_item_embeddings")
        x_s = h_s * mlp("intent_score_embeddings")
    h_int = torch.relu(_lin(sd, "intent_embeddings", intent))[:, None, :].expand(B, L, -1)
    weights = _lin(sd, "weight_embeddings", torch.cat([x_i, x_s, h_u, h_int], dim=-1))
    return weights, (weights * scores).sum(dim=2)


def forward(sd: State, cfg: IntelConfig, batch: Dict[str, object]) -> Dict[str, Tensor]:
    """IntEL.py:117-124."""
    intent = predict_intent(sd, cfg, batch)
    weights, ens = predict_ensemble(sd, cfg, batch, intent)
    return {"weights": weights, "ens_score": ens, "intents": intent}


# --------------------------------------------------------------------------- losses
def _pair_setup(ens: Tensor, batch: Dict[str, object]):
    L = ens.shape[1]
    valid = torch.arange(L, device=ens.device)[None, :] < batch["session_len"][:, None]
    valid2 = valid[:, :, None] * valid[:, None, :]
    rank = batch["ranking"].clamp(min=0)
    diff = ens[:, :, None] - ens[:, None, :]
    return valid, valid2, rank, diff


def list_loss(out: Dict[str, Tensor], batch: Dict[str, object], cal_diversity: int, alpha: float) -> Tensor:
    """Listloss.py:12-43 (Plackett-Luce style) incl. the diversity regulariser."""
    ens = out["ens_score"]
    valid, valid2, rank, diff = _pair_setup(ens, batch)
    pos = rank > 0
    m = (rank[:, :, None] > rank[:, None, :]) * valid2
    e = torch.exp(-diff) * m
    per_item = ((e.sum(dim=2) + 1) * pos).clamp(min=1).log()
    loss = (per_item.sum(dim=1) / pos.sum(dim=-1)).mean()
    if cal_diversity:
        x = batch["scores"]
        xd = x[:, :, None, :] - x[:, None, :, :]
        ex = torch.exp(-diff)
        up = ((ex[..., None] * (xd - diff[..., None]) * m[..., None]).sum(dim=2)) ** 2
        num = (out["weights"] * up).sum(-1)
        den = 2 * (1 + (ex * m).sum(dim=2)) ** 2
        div = -((num / den * pos).sum(dim=-1) / pos.sum(dim=-1)).mean()
        loss = (loss + div * alpha).to(loss.dtype)
    return loss


def bpr_loss(out: Dict[str, Tensor], batch: Dict[str, object], cal_diversity: int, alpha: float,
             noise: Tensor) -> Tensor:
    """BPRloss.py:12-56.  ``noise`` [B,L,L] in [0,1) replaces torch.rand_like (BPRloss.py:26)."""
    ens = out["ens_score"]
    valid, valid2, rank, diff = _pair_setup(ens, batch)
    pos = rank > 0
    g = (rank[:, :, None] - rank[:, None, :]) * valid2
    sim = (g.max() + 1 - g) * (g > 0)
    cand = ((sim == sim.max(dim=-1)[0][:, :, None]) * (g > 0)).int()
    pick = (cand + noise.float() / 10).argmax(dim=-1)
    sel = F.one_hot(pick, num_classes=ens.shape[1])
    per_item = (-torch.sigmoid(diff).log() * sel).sum(dim=-1) * pos
    loss = (per_item.sum(dim=-1) / pos.sum(dim=-1)).mean()
    if cal_diversity:
        x = batch["scores"]
        xd = x[:, :, None, :] - x[:, None, :, :]
        sg = torch.sigmoid(diff)
        dsg = sg * (1 - sg)
        zd = (dsg[..., None] * (xd - diff[..., None]) ** 2 * sel[..., None]).sum(dim=2)
        a = (zd * out["weights"]).sum(dim=-1) * pos
        div = -(a.sum(dim=-1) / pos.sum(dim=-1)).mean()
        loss = (loss + div * alpha).to(loss.dtype)
    return loss


def mse_loss(out: Dict[str, Tensor], batch: Dict[str, object], cal_diversity: int, alpha: float) -> Tensor:
    """MSEloss.py:12-30."""
    ens = out["ens_score"]
    valid, _, rank, _ = _pair_setup(ens, batch)
    loss = ((((ens - rank) ** 2) * valid).sum(dim=-1) / valid.sum(dim=-1)).mean()
    if cal_diversity:
        d = out["weights"] * (batch["scores"] - ens[:, :, None]) ** 2
        div = -((d * valid[:, :, None]).sum(dim=-1).sum(dim=-1) / valid.sum(dim=-1)).mean()
        loss = (loss + div * alpha).to(loss.dtype)
    return loss


def intent_loss(pred: Tensor, true: Tensor, kl_weight: float, kl_temp: float) -> Tuple[Tensor, Tensor, Tensor]:
    """BaseIntloss.py:30-67 -> (intent_loss, ce, kl*T^2).  ``true`` float64 [B,I]."""
    if pred.min() == 0:
        soft = pred + 1e-6
        soft = soft / soft.sum(dim=-1)[:, None]
    else:
        soft = pred
    ce = -(((true > 0) * true * soft.log() + (true == 0) * (1 - soft).log())).sum(dim=-1).mean()
    t32 = true.to(pred.dtype)
    kl = (torch.xlogy(t32, t32) - t32 * soft.log()).double().sum(dim=-1).mean() * kl_temp * kl_temp
    return ce * (1 - kl_weight) + kl * kl_weight, ce, kl


def total_loss(kind: str, out: Dict[str, Tensor], batch: Dict[str, object], *, cal_diversity: int = 0,
               diversity_alpha: float = 0.01, intent_weight: float = 0.1, ensemble_weight: float = 1.0,
               kl_weight: float = 0.5, kl_temp: float = 2.0, noise: Optional[Tensor] = None):
    """Int{List,BPR,MSE}loss.forward -> (loss, ensemble_loss, intent_loss)."""
    il, _, _ = intent_loss(out["intents"], batch["intents"], kl_weight, kl_temp)
    if kind == "list":
        el = list_loss(out, batch, cal_diversity, diversity_alpha)
    elif kind == "bpr":
        el = bpr_loss(out, batch, cal_diversity, diversity_alpha, noise)
    elif kind == "mse":
        el = mse_loss(out, batch, cal_diversity, diversity_alpha)
    else:
        raise ValueError(kind)
    return el * ensemble_weight + il * intent_weight, el, il


# --------------------------------------------------------------------------- evaluation
def evaluate_method(prediction_scores: Sequence[np.ndarray], ranking_lists: Sequence[np.ndarray],
                    pos_nums: Dict[str, np.ndarray], topk: Sequence[int], metrics: Sequence[str],
                    session_len: Sequence[int]) -> Dict[str, float]:
    """BaseRunner.py:57-131 with every argsort made *stable* - the reference uses numpy's
    default unstable sort, so results are defined only up to exact ties; the stable
    idealisation below is the tie rule the CUDA path implements."""
    n = len(prediction_scores)
    session_len = np.asarray(session_len)[:n]
    pos_nums = {k: np.asarray(v)[:n] for k, v in pos_nums.items()}
    max_len = int(max(session_len.max(), max(topk)))
    pred = np.zeros((n, max_len), dtype=np.float64)
    rank = np.full((n, max_len), -2, dtype=np.int64)
    for i in range(n):
        m = min(int(session_len[i]), len(prediction_scores[i]))
        pred[i, :m] = np.asarray(prediction_scores[i][:m], dtype=np.float32)
        rank[i, :m] = ranking_lists[i][:m]
    rows = np.arange(n)[:, None]
    order = np.argsort(rank, axis=1, kind="stable")[:, ::-1]
    rank, pred = rank[rows, order], pred[rows, order]
    rank[rank < 0] = 0
    asc = np.argsort(pred, axis=1, kind="stable")
    disc = 1.0 / np.log2(np.arange(max_len) + 2.0)
    res: Dict[str, float] = {}
    total_pos = np.sum(np.array(list(pos_nums.values())), axis=0).reshape(-1, 1)
    for btype, cnt in pos_nums.items():
        behavior = btype.split("_")[1].split("num")[0]
        all_pos = total_pos if "click" in btype else cnt.reshape(-1, 1)
        sel = np.where(all_pos[:, 0] > 0)[0]
        hit_pos = (asc < all_pos)[sel]
        ap = all_pos[sel]
        for k in topk:
            mk = min(k, max_len)
            for metric in metrics:
                key = f"{behavior}_{metric}@{k}"
                if metric == "HR":
                    res[key] = (hit_pos[:, -mk:].sum(axis=1) > 0).mean()
                elif metric == "NDCG":
                    if k == 1:
                        continue
                    dcg = (hit_pos[:, -mk:] * disc[:mk][::-1]).sum(axis=1)
                    idcg = ((np.arange(mk)[None, :] < ap) * disc[:mk]).sum(axis=1)
                    res[key] = (dcg / idcg).mean()
                else:
                    raise ValueError(f"Undefined evaluation metric: {metric}.")
    desc = asc[:, ::-1]
    gains = rank[rows, desc]
    ideal = np.sort(gains, axis=1)[:, ::-1]
    for k in topk:
        with np.errstate(invalid="ignore", divide="ignore"):
            res[f"NDCG@{k}"] = ((gains[:, :k] * disc[:k]).sum(axis=1) / (ideal[:, :k] * disc[:k]).sum(axis=1)).mean()
    return res


def evaluate_intents(true_intents: np.ndarray, predict_intents: np.ndarray, topk: Sequence[int]) -> Dict[str, float]:
    """BaseRunner.py:133-150 (stable sorts, see evaluate_method)."""
    true_intents, predict_intents = np.asarray(true_intents), np.asarray(predict_intents)
    label = np.argmax(true_intents, axis=1).reshape(-1, 1)
    asc = np.argsort(predict_intents, axis=1, kind="stable")
    desc = asc[:, ::-1]
    rows = np.arange(len(predict_intents))[:, None]
    got = true_intents[rows, desc]
    ideal = np.sort(true_intents, axis=1)[:, ::-1]
    disc = 1.0 / np.log2(np.arange(40) + 2.0)
    res: Dict[str, float] = {}
    for k in topk:
        res[f"Int-NDCG@{k}"] = ((got[:, :k] * disc[:k]).sum(axis=1) / (ideal[:, :k] * disc[:k]).sum(axis=1)).mean()
        res[f"Int-HR@{k}"] = ((asc == label)[:, -k:].sum(axis=-1) > 0).mean()
    return res


# --------------------------------------------------------------------------- fixed-weight baselines
def single_sort(batch: Dict[str, object], column: int) -> Dict[str, Tensor]:
    x = batch["scores"].float()
    return {"weights": torch.zeros_like(x), "ens_score": x[:, :, column]}


def borda(batch: Dict[str, object]) -> Dict[str, Tensor]:
    """Borda.py:23-30: rank of every item inside each basic list over the *padded* length
    (stable sort here), mean over the K lists."""
    x = batch["scores"].float()
    rank = torch.argsort(torch.argsort(x, dim=1, stable=True), dim=1, stable=True)
    w = torch.ones_like(x) / x.size(2)
    return {"weights": w, "ens_score": (w * rank).sum(dim=2)}


def awelv(sd: State, batch: Dict[str, object]) -> Dict[str, Tensor]:
    """models/supervise/aWELv.py:28-39: per-user softmax over <h_u, h_m>, broadcast over the list."""
    scores = batch["scores"].float()
    h_u = sd["uid_embeddings.weight"][batch["u_id_c"]]                      # [B, h]
    logits = torch.stack([(h_u * sd["model_embeddings.weight"][m][None, :]).sum(dim=1)
                          for m in range(scores.size(2))], dim=1)           # [B, K]
    w = logits.unsqueeze(1).repeat(1, scores.size(1), 1).softmax(dim=-1)
    return {"weights": w, "ens_score": (w * scores).sum(dim=2)}


def awelv_int(sd: State, cfg: IntelConfig, batch: Dict[str, object]) -> Dict[str, Tensor]:
    """models/supervise/aWELv_Int.py:97-113: the intent predictor of IntEL (same predict_intent, :66-95), then
    per-session softmax over <[h_u || intent_embeddings(intent)], h_m>, broadcast over the list."""
    scores = batch["scores"].float()
    intent = predict_intent(sd, cfg, batch)
    h = torch.cat([sd["uid_embeddings.weight"][batch["u_id_c"]], _lin(sd, "intent_embeddings", intent)], dim=1)
    w = (h @ sd["model_embeddings.weight"].t()).softmax(dim=1).unsqueeze(1).repeat(1, scores.size(1), 1)
    return {"weights": w, "ens_score": (w * scores).sum(dim=2), "intents": intent}


def awelv_intel(sd: State, cfg: IntelConfig, batch: Dict[str, object]) -> Dict[str, Tensor]:
    """models/supervise/aWELv_IntEL.py:113-201: IntEL's intent predictor and self-attention stacks, the gate form of the
    intent conditioning (h * MLP(intent)), then ONE weight vector per session from the unmasked mean over the list slots
    of the gated streams, softmax applied twice (:197-198)."""
    dt = sd["pred_layer.weight"].dtype
    intent = predict_intent(sd, cfg, batch)
    items, scores = batch["i_id_s"], batch["scores"].to(dt)
    h_i = sd["iid_embeddings.weight"][items]
    if "item_embeddings.weight" in sd:
        h_i = torch.cat([h_i, sd["item_embeddings.weight"][batch["i_class_c"]]], dim=2)
    h_u = torch.relu(sd["uid_embeddings.weight"][batch["u_id_c"]])
    h_i = _self_att_stack(sd, "i", h_i, cfg)
    h_s = _self_att_stack(sd, "s", _lin(sd, "score_embeddings", scores), cfg)

    def mlp(name):
        t = torch.relu(intent @ sd[f"{name}.0.weight"].t() + sd[f"{name}.0.bias"])
        return (t @ sd[f"{name}.2.weight"].t())[:, None, :]
    x_i = (h_i * mlp("intent_item_embeddings")).mean(dim=1)
    x_s = (h_s * mlp("intent_score_embeddings")).mean(dim=1)
    h_int = torch.relu(_lin(sd, "intent_embeddings", intent))
    w = _lin(sd, "weight_embeddings", torch.cat([x_i, x_s, h_u, h_int], dim=-1)).softmax(dim=-1)
    w = w.unsqueeze(1).repeat(1, items.size(1), 1).softmax(dim=-1)
    return {"weights": w, "ens_score": (w * scores).sum(dim=2), "intents": intent}


def lambdarank_scorer(sd: State, batch: Dict[str, object]) -> Dict[str, Tensor]:
    """models/supervise/LambdaRank.py:39-47: MLP over [E_iid || category id || scores] per slot, softmax over the padded list."""
    scores = batch["scores"].float()
    h = torch.cat([sd["iid_embeddings.weight"][batch["i_id_s"]], batch["i_class_c"].unsqueeze(2).float(), scores], dim=2)
    n_lin = sum(1 for k in sd if k.startswith("mlp.Linear-") and k.endswith(".weight"))
    for i in range(n_lin):
        h = h @ sd[f"mlp.Linear-{i}.weight"].t() + sd[f"mlp.Linear-{i}.bias"]
        if i < n_lin - 1:
            h = torch.relu(h)
    return {"weights": torch.zeros_like(scores), "ens_score": h.squeeze(-1).softmax(dim=-1)}


def compute_lambda(true_scores: Tensor, temp_scores: Tensor, session_len: Tensor) -> Tensor:
    """helpers/LambdaRankRunner.py:315-344 (compute_lambda_new) per session with explicit loops over the valid pairs:
    Delta_ij = |g_i d_j + g_j d_i - g_i d_i - g_j d_j| / IDCG with g = 2^t - 1 and d_j = 1/log2(j+2) of the list slot j,
    Rho_ij = 1/(1+exp(s_i-s_j)); Lambda_i = sum_{t_i>t_j} Delta Rho_ij - sum_{t_i<t_j} Delta Rho_ji.  IDCG = 0 -> NaN row."""
    B, L = true_scores.shape
    t = true_scores.clamp(min=0).numpy()
    s = temp_scores.detach().numpy().astype(np.float32)
    d = (1.0 / np.log2(np.arange(L, dtype=np.float32) + np.float32(2.0))).astype(np.float32)
    out = np.zeros((B, L), dtype=np.float32)
    for b in range(B):
        n = min(int(session_len[b]), L)
        g = (2.0 ** t[b] - 1).astype(np.float32)
        ideal = np.sort(g)[::-1]
        idcg = np.float32((ideal[:n].astype(np.float64) * d[:n]).sum())
        if not idcg > 0:
            out[b] = np.nan
            continue
        for i in range(n):
            acc = 0.0
            for j in range(n):
                if t[b, i] == t[b, j]:
                    continue
                delta = np.float32(abs(np.float32(np.float32(np.float32(g[i] * d[j]) + np.float32(g[j] * d[i]))
                                                  - np.float32(g[i] * d[i])) - np.float32(g[j] * d[j]))) / idcg
                if t[b, i] > t[b, j]:
                    acc += float(delta) / (1.0 + np.exp(np.float64(s[b, i]) - np.float64(s[b, j])))
                else:
                    acc -= float(delta) / (1.0 + np.exp(np.float64(s[b, j]) - np.float64(s[b, i])))
            out[b, i] = acc
    return torch.from_numpy(out)


def random_fusion(batch: Dict[str, object], raw_weights: Tensor) -> Dict[str, Tensor]:
    """GeneralSeq.py:23-32 with the uniform draw passed in."""
    x = batch["scores"].float()
    w = F.softmax(raw_weights, dim=2)
    return {"weights": w, "ens_score": (w * x).sum(dim=2)}
