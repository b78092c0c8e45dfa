"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Runs only in the build container (needs /root/reference).  The reference has no tests
or golden vectors, so these files are what pins the oracle: every
array below is produced by importing IntEL/src/{models,loss,helpers} from the read-only
reference tree with three py3.12/numpy-2 import shims and calling

    IntEL.forward                      (models/IntEL/IntEL.py:117-124)
    Int{List,BPR,MSE}loss.forward      (loss/*.py)   + autograd for parameter gradients
    BaseRunner.evaluate_method         (helpers/BaseRunner.py:57-131)
    BaseRunner.evaluate_intents        (helpers/BaseRunner.py:133-150)
    SingleSort/Borda.forward           (models/unsupervise/*.py)
    aWELv / aWELv_Int / aWELv_IntEL.forward (models/supervise/*.py)
    LambdaRankRunner.compute_lambda_new (helpers/LambdaRankRunner.py:315-344)

on seeded synthetic batches from intel_sigir2023_b200.synthetic.  BPR's torch.rand_like
is monkey-patched to return a saved noise tensor so its negative choice is replayable.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_SRC = "/root/reference/IntEL/src"

from intel_sigir2023_b200.config import IntelConfig            # noqa: E402
from intel_sigir2023_b200 import synthetic                      # noqa: E402


def import_reference():
    """py3.12 / numpy-2 shims, then import the reference packages from the read-only tree."""
    np.object = object
    np.float = float
    np.int = int
    imp = types.ModuleType("imp")
    imp.reload = importlib.reload
    sys.modules["imp"] = imp
    sys.path.insert(0, REF_SRC)
    from models.IntEL import IntEL as ref_intel
    from models.unsupervise import SingleSort as ref_single, Borda as ref_borda
    from models.supervise import aWELv as ref_awelv, aWELv_Int as ref_awelv_int, aWELv_IntEL as ref_awelv_intel
    from loss import IntListloss, IntBPRloss, IntMSEloss, Listloss
    from helpers import BaseRunner
    return dict(aWELv=ref_awelv.aWELv, aWELv_Int=ref_awelv_int.aWELv_Int, aWELv_IntEL=ref_awelv_intel.aWELv_IntEL, Listloss=Listloss.Listloss, IntEL=ref_intel.IntEL, SingleSort=ref_single.SingleSort, Borda=ref_borda.Borda,
                list=IntListloss.IntListloss, bpr=IntBPRloss.IntBPRloss, mse=IntMSEloss.IntMSEloss,
                BaseRunner=BaseRunner.BaseRunner)


class _Corpus:
    def __init__(self, cfg: IntelConfig):
        self.itemfnum = [cfg.class_rows]
        self.contextfnum = [cfg.ctx_rows]
        self.zero_int = np.zeros(cfg.intent_num)
        self.max_uid = cfg.user_rows - 1
        self.max_iid = cfg.item_rows - 1


def _args(cfg: IntelConfig, **loss_kw):
    a = argparse.Namespace(**{k: v for k, v in cfg.to_dict().items()})
    a.device = torch.device("cpu")
    a.model_path = "/tmp/intel_golden/model.pt"
    a.buffer = 1
    a.cal_diversity = 1
    a.diversity_alpha = 0.05
    a.intent_weight = 0.1
    a.ensemble_weight = 1.0
    a.kl_temp = 2.0
    a.kl_weight = 0.5
    for k, v in loss_kw.items():
        setattr(a, k, v)
    return a


CASES = {
    # name: (config overrides, corpus sizes, batch spec, seed)
    "default_bert": (dict(), dict(n_item=60, n_class=9, n_user=11, n_ctx=13, model_num=3, intent_num=24),
                     dict(batch_size=6, max_len=12, min_len=1), 0),
    "script_pl_gru": (dict(encoder="GRU4Rec", context_emb_size=32, intent_emb_size=32, cross_attn_qsize=64,
                           num_heads=2, num_layers=2),
                      dict(n_item=80, n_class=7, n_user=9, n_ctx=17, model_num=3, intent_num=30, history_max=6),
                      dict(batch_size=5, max_len=17, min_len=3), 1),
    "script_bpr_gru_k4": (dict(encoder="GRU4Rec", context_emb_size=64, intent_emb_size=32, cross_attn_qsize=32,
                               num_heads=2, num_layers=2),
                          dict(n_item=50, n_class=5, n_user=7, n_ctx=11, model_num=4, intent_num=20, history_max=5),
                          dict(batch_size=4, max_len=9, min_len=2), 2),
    "direct_att_bert": (dict(cross_attention=0, num_heads=2, num_layers=1, cross_attn_qsize=16),
                        dict(n_item=40, n_class=6, n_user=8, n_ctx=9, model_num=3, intent_num=18, history_max=7),
                        dict(batch_size=5, max_len=10, min_len=2), 3),
    "wide_intent_bert_full": (dict(num_heads=2, num_layers=2),
                              dict(n_item=90, n_class=12, n_user=10, n_ctx=12, model_num=4, intent_num=132,
                                   history_max=20),
                              dict(batch_size=4, max_len=33, min_len=33), 4),
}


def make_case(ref, name: str):
    over, csz, bsz, seed = CASES[name]
    corpus = synthetic.CorpusSpec(**csz)
    cfg = IntelConfig(item_rows=corpus.item_rows, class_rows=corpus.n_class, user_rows=corpus.user_rows,
                      ctx_rows=corpus.n_ctx, intent_num=corpus.intent_num, model_num=corpus.model_num,
                      history_max=corpus.history_max, **over)
    batch = synthetic.make_batch(corpus, synthetic.BatchSpec(**bsz), seed=seed)
    torch.manual_seed(100 + seed)
    model = ref["IntEL"](_args(cfg), _Corpus(cfg))
    # PyTorch-default LayerNorm init is (1,0); perturb so its gradients are exercised
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "layer_norm" in n:
                p.add_(0.1 * torch.randn_like(p))
    model.eval()
    assert set(model.state_dict().keys()) == set(cfg.param_shapes().keys()), \
        set(model.state_dict().keys()) ^ set(cfg.param_shapes().keys())
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == cfg.param_shapes()[k], (k, v.shape, cfg.param_shapes()[k])

    out = {"cfg": np.frombuffer(json.dumps(cfg.to_dict()).encode(), dtype=np.uint8)}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out["batch." + k] = v.numpy()
    for k, v in model.state_dict().items():
        out["state." + k] = v.detach().numpy().copy()

    B, L = batch["i_id_s"].shape
    noise = torch.rand(B, L, L, generator=torch.Generator().manual_seed(7 + seed))
    out["bpr_noise"] = noise.numpy()
    real_rand_like = torch.rand_like
    for kind in ("list", "bpr", "mse"):
        crit = ref[kind](_args(cfg))
        model.zero_grad()
        res = model(dict(batch))
        if kind == "bpr":
            torch.rand_like = lambda t, dtype=None, **kw: noise.to(dtype or t.dtype)
        try:
            loss, ens_l, int_l = crit(res, batch)
        finally:
            torch.rand_like = real_rand_like
        loss.backward()
        if kind == "list":
            for k in ("weights", "ens_score", "intents"):
                out["out." + k] = res[k].detach().numpy().copy()
        out[f"loss.{kind}"] = np.array([loss.item(), ens_l.item(), int_l.item()], dtype=np.float64)
        for n, p in model.named_parameters():
            out[f"grad.{kind}.{n}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", f"model_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB  loss(list/bpr/mse)=",
          out["loss.list"][0], out["loss.bpr"][0], out["loss.mse"][0])


def make_eval(ref):
    """evaluate_method / evaluate_intents / fixed-weight baselines on tie-free synthetic sets."""
    runner = ref["BaseRunner"]
    out = {}
    # set A: general label mix; set B: no 'pay' items so the reference's first-favnum-columns
    # quirk (BaseRunner.py:88-99) does not cut through a tie group -> fav_* is well defined
    for tag, n, L, lo, seed in (("A", 64, 23, 4, 0), ("B", 48, 9, 2, 1), ("C", 40, 130, 60, 2)):
        pred, ranking, pos, slen = synthetic.eval_set(n, L, lo, seed=seed)
        if tag == "B":
            ranking = torch.where(ranking == 3, torch.full_like(ranking, 1), ranking)
            pos["c_clicknum_i"] = pos["c_clicknum_i"] + pos["c_paynum_i"]
            pos["c_paynum_i"] = torch.zeros_like(pos["c_paynum_i"])
        # the reference is fed per-batch padded rows (BaseRunner.py:338-339)
        scores = [pred[i].numpy() for i in range(n)]
        ranks = [ranking[i].numpy() for i in range(n)]
        posd = {k: v.numpy().copy() for k, v in pos.items()}
        res = runner.evaluate_method(scores, ranks, {k: v.copy() for k, v in posd.items()},
                                     [3, 1, 5, 10], ["NDCG", "HR"], slen.numpy().tolist())
        out[f"{tag}.pred"], out[f"{tag}.ranking"], out[f"{tag}.session_len"] = pred.numpy(), ranking.numpy(), slen.numpy()
        for k, v in posd.items():
            out[f"{tag}.pos.{k}"] = v
        for k, v in res.items():
            out[f"{tag}.metric.{k}"] = np.float64(v)
    g = torch.Generator().manual_seed(5)
    true_int = synthetic._intent_rows(37, 50, 6, g, torch.device("cpu")).numpy()
    pred_int = torch.softmax(torch.randn(37, 50, generator=g), dim=-1).numpy()
    res = runner.evaluate_intents(runner, true_int, pred_int, topk=[1, 3, 5, 10, 30])
    out["I.true"], out["I.pred"] = true_int, pred_int
    for k, v in res.items():
        out[f"I.metric.{k}"] = np.float64(v)

    corpus = synthetic.CorpusSpec(n_item=30, n_class=5, n_user=5, n_ctx=5, model_num=3, intent_num=6, history_max=2)
    batch = synthetic.make_batch(corpus, synthetic.BatchSpec(batch_size=7, max_len=11, min_len=3), seed=9)
    a = argparse.Namespace(device=torch.device("cpu"), model_path="", buffer=1, choose_list="pCVR")
    cfg = IntelConfig(item_rows=31, class_rows=5, user_rows=6, ctx_rows=5, intent_num=6)
    out["F.scores"] = batch["scores"].numpy()
    out["F.session_len"] = batch["session_len"].numpy()
    out["F.single_pCVR"] = ref["SingleSort"](a, _Corpus(cfg)).forward(batch)["ens_score"].numpy()
    out["F.borda"] = ref["Borda"](a, _Corpus(cfg)).forward(batch)["ens_score"].numpy()
    path = os.path.join(ROOT, "tests", "golden", "eval.npz")
    np.savez_compressed(path, **out)
    print("eval:", os.path.getsize(path) // 1024, "KiB", {k: float(v) for k, v in out.items() if k.startswith("A.metric.NDCG")})


def make_awelv(ref):
    """aWELv.forward (models/supervise/aWELv.py:28-39) + Listloss with the diversity term (script/baselines.sh:33)."""
    corpus = synthetic.CorpusSpec(n_item=90, n_class=7, n_user=23, n_ctx=11, model_num=3, intent_num=20)
    cfg = IntelConfig(item_rows=corpus.item_rows, class_rows=corpus.n_class, user_rows=corpus.user_rows,
                      ctx_rows=corpus.n_ctx, intent_num=corpus.intent_num, model_num=corpus.model_num,
                      history_max=corpus.history_max)
    batch = synthetic.make_batch(corpus, synthetic.BatchSpec(batch_size=13, max_len=17, min_len=2), seed=5)
    args = _args(cfg, hidden_size=32, diversity_alpha=0.05)
    torch.manual_seed(77)
    model = ref["aWELv"](args, _Corpus(cfg))
    out = {"user_rows": np.array([cfg.user_rows]), "model_num": np.array([cfg.model_num]), "hidden_size": np.array([32])}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out["batch." + k] = v.numpy()
    for k, v in model.state_dict().items():
        out["state." + k] = v.detach().numpy().copy()
    res = model(dict(batch))
    loss = ref["Listloss"](args)(res, batch)
    loss = loss[0] if isinstance(loss, (tuple, list)) else loss
    loss.backward()
    out["out.weights"] = res["weights"].detach().numpy().copy()
    out["out.ens_score"] = res["ens_score"].detach().numpy().copy()
    out["loss.list"] = np.array([loss.item()], dtype=np.float64)
    for n, p in model.named_parameters():
        out["grad.list." + n] = p.grad.numpy().copy()
    path = os.path.join(ROOT, "tests", "golden", "awelv.npz")
    np.savez_compressed(path, **out)
    print(f"awelv: {os.path.getsize(path) / 1024:.0f} KiB  loss(list)=", out["loss.list"][0])


AWELV_INT_CASES = {
    # the script's flags (script/baselines.sh:40: GRU4Rec, context 32, intent 32, user 16) and the bare defaults (BERT4Rec)
    "gru": (dict(encoder="GRU4Rec", context_emb_size=32, intent_emb_size=32, u_emb_size=16),
            dict(n_item=70, n_class=7, n_user=19, n_ctx=11, model_num=3, intent_num=27, history_max=6),
            dict(batch_size=9, max_len=14, min_len=2), 6),
    "bert": (dict(encoder="BERT4Rec", u_emb_size=16),
             dict(n_item=50, n_class=5, n_user=12, n_ctx=9, model_num=4, intent_num=20, history_max=8),
             dict(batch_size=7, max_len=11, min_len=1), 7),
}


AWELV_INTEL_CASES = {
    # the script's flags (script/baselines.sh:47: GRU4Rec, u_emb 16, 2 heads, 2 layers) and the bare defaults (BERT4Rec, 1/1)
    "gru": (dict(encoder="GRU4Rec", context_emb_size=32, intent_emb_size=32, cross_attn_qsize=64, num_heads=2, num_layers=2,
                 u_emb_size=16, cross_attention=0),
            dict(n_item=70, n_class=7, n_user=19, n_ctx=11, model_num=3, intent_num=27, history_max=6),
            dict(batch_size=7, max_len=15, min_len=2), 8),
    "bert": (dict(encoder="BERT4Rec", cross_attention=0),
             dict(n_item=50, n_class=5, n_user=12, n_ctx=9, model_num=4, intent_num=20, history_max=8),
             dict(batch_size=6, max_len=11, min_len=1), 9),
}


def make_awelv_intel(ref):
    """aWELv_IntEL.forward (models/supervise/aWELv_IntEL.py:113-201) + IntListloss."""
    make_awelv_int(ref, model_key="aWELv_IntEL", cases=AWELV_INTEL_CASES, prefix="awelv_intel")


def make_awelv_int(ref, model_key="aWELv_Int", cases=None, prefix="awelv_int"):
    """aWELv_Int.forward (models/supervise/aWELv_Int.py:66-113) + IntListloss (script/baselines.sh:40)."""
    for name, (over, csz, bsz, seed) in (cases or AWELV_INT_CASES).items():
        corpus = synthetic.CorpusSpec(**csz)
        cfg = IntelConfig(item_rows=corpus.item_rows, class_rows=corpus.n_class, user_rows=corpus.user_rows,
                          ctx_rows=corpus.n_ctx, intent_num=corpus.intent_num, model_num=corpus.model_num,
                          history_max=corpus.history_max, **over)
        batch = synthetic.make_batch(corpus, synthetic.BatchSpec(**bsz), seed=seed)
        args = _args(cfg, user_emb_size=cfg.u_emb_size)
        torch.manual_seed(200 + seed)
        model = ref[model_key](args, _Corpus(cfg))
        with torch.no_grad():
            for n, p in model.named_parameters():
                if "layer_norm" in n:
                    p.add_(0.1 * torch.randn_like(p))
        model.eval()
        out = {"cfg": np.frombuffer(json.dumps(cfg.to_dict()).encode(), dtype=np.uint8)}
        for k, v in batch.items():
            if torch.is_tensor(v):
                out["batch." + k] = v.numpy()
        for k, v in model.state_dict().items():
            out["state." + k] = v.detach().numpy().copy()
        res = model(dict(batch))
        loss, ens_l, int_l = ref["list"](args)(res, batch)
        loss.backward()
        for k in ("weights", "ens_score", "intents"):
            out["out." + k] = res[k].detach().numpy().copy()
        out["loss.list"] = np.array([loss.item(), ens_l.item(), int_l.item()], dtype=np.float64)
        for n, p in model.named_parameters():
            out["grad.list." + n] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
        path = os.path.join(ROOT, "tests", "golden", f"{prefix}_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{prefix}_{name}: {os.path.getsize(path) / 1024:.0f} KiB  loss(list)=", out["loss.list"])


def make_lambdarank_model(ref):
    """LambdaRank.forward (models/supervise/LambdaRank.py:39-47) and one training signal of LambdaRankRunner.fit (:240-259):
    lambdas = compute_lambda_new(clamp(ranking, 0), ens.detach(), session_len); ens.backward(lambdas)."""
    from models.supervise import LambdaRank as ref_lr
    from helpers import LambdaRankRunner
    for name, hidden, K, seed in (("h32", "32", 3, 12), ("h24_8", "24,8", 4, 13)):
        corpus = synthetic.CorpusSpec(n_item=60, n_class=9, n_user=7, n_ctx=5, model_num=K, intent_num=6, history_max=2)
        cfg = IntelConfig(item_rows=corpus.item_rows, class_rows=corpus.n_class, user_rows=corpus.user_rows,
                          ctx_rows=corpus.n_ctx, intent_num=corpus.intent_num, model_num=K, history_max=2)
        batch = synthetic.make_batch(corpus, synthetic.BatchSpec(batch_size=8, max_len=13, min_len=2), seed=seed)
        args = _args(cfg, hidden_size=hidden, i_emb_size=32)
        torch.manual_seed(300 + seed)
        model = ref_lr.LambdaRank(args, _Corpus(cfg))
        model.eval()
        out = {"hidden_size": np.frombuffer(hidden.encode(), dtype=np.uint8), "model_num": np.array([K]),
               "item_rows": np.array([cfg.item_rows])}
        for k, v in batch.items():
            if torch.is_tensor(v):
                out["batch." + k] = v.numpy()
        for k, v in model.state_dict().items():
            out["state." + k] = v.detach().numpy().copy()
        res = model(dict(batch))
        ens = res["ens_score"]
        lam = LambdaRankRunner.LambdaRankRunner.compute_lambda_new(None, torch.clamp(batch["ranking"], min=0), ens.detach(),
                                                                   batch["session_len"])
        lam = torch.nan_to_num(lam, nan=0.0)        # sessions without positives: the reference stops on NaN (:247-255)
        model.zero_grad()
        ens.backward(lam)
        out["out.ens_score"], out["out.weights"], out["lambdas"] = ens.detach().numpy().copy(), res["weights"].numpy().copy(), lam.numpy()
        for n, p in model.named_parameters():
            out["grad." + n] = p.grad.numpy().copy()
        path = os.path.join(ROOT, "tests", "golden", f"lambdarank_model_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"lambdarank_model_{name}: {os.path.getsize(path) // 1024} KiB", float(np.abs(out["grad.iid_embeddings.weight"]).max()))


def make_lambdarank(ref):
    """LambdaRankRunner.compute_lambda_new (helpers/LambdaRankRunner.py:315-344) on ragged label sets: softmaxed scores
    as LambdaRank.forward emits them (set S), wide raw scores (set W), and a set with sessions without positives (set Z:
    IDCG = 0 -> NaN rows in the reference)."""
    from helpers import LambdaRankRunner
    fn = LambdaRankRunner.LambdaRankRunner.compute_lambda_new
    out = {}
    for tag, n, L, lo, seed in (("S", 24, 19, 2, 0), ("W", 16, 50, 50, 1), ("Z", 12, 9, 1, 2), ("L", 6, 130, 40, 3)):
        pred, ranking, pos, slen = synthetic.eval_set(n, L, lo, seed=seed)
        scores = pred.float()
        if tag == "S":
            scores = scores.softmax(dim=-1)
        if tag == "W":
            scores = scores * 6
        if tag == "Z":
            ranking[::3] = torch.where(ranking[::3] > 0, torch.zeros_like(ranking[::3]), ranking[::3])
        true_scores = torch.clamp(ranking, min=0)
        lam = fn(None, true_scores, scores, slen)
        out[f"{tag}.ranking"], out[f"{tag}.scores"], out[f"{tag}.session_len"] = ranking.numpy(), scores.numpy(), slen.numpy()
        out[f"{tag}.lambdas"] = lam.numpy()
    path = os.path.join(ROOT, "tests", "golden", "lambdarank.npz")
    np.savez_compressed(path, **out)
    print("lambdarank:", os.path.getsize(path) // 1024, "KiB", {k: float(np.nanmax(np.abs(v))) for k, v in out.items() if k.endswith("lambdas")},
          "nan rows in Z:", int(np.isnan(out["Z.lambdas"]).all(axis=1).sum()))


if __name__ == "__main__":
    ref = import_reference()
    torch.set_num_threads(4)
    if "--awelv-only" in sys.argv:
        make_awelv(ref)
        sys.exit(0)
    if "--lambdarank-only" in sys.argv:
        make_lambdarank(ref)
        make_lambdarank_model(ref)
        sys.exit(0)
    if "--awelv-intel-only" in sys.argv:
        make_awelv_intel(ref)
        sys.exit(0)
    if "--awelv-int-only" in sys.argv:
        make_awelv_int(ref)
        sys.exit(0)
    for name in CASES:
        make_case(ref, name)
    make_eval(ref)
    make_awelv(ref)
    make_awelv_int(ref)
    make_awelv_intel(ref)
    make_lambdarank(ref)
    make_lambdarank_model(ref)
