"""Size-independent properties at BASELINE.json's full per-step sizes (the oracle cannot run these in seconds):
shard additivity of the evaluation sums, batch-split invariance of the forward pass, linearity of the fusion,
a directional finite-difference check of the fused loss gradients, permutation invariance of NDCG."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(cfg_kw, corpus_kw, seed=0):
    from intel_sigir2023_b200 import synthetic
    from intel_sigir2023_b200.config import IntelConfig
    from intel_sigir2023_b200.IntEL import IntEL
    corpus = synthetic.CorpusSpec(**corpus_kw)
    cfg = IntelConfig(item_rows=corpus.item_rows, class_rows=corpus.n_class, user_rows=corpus.user_rows,
                      ctx_rows=corpus.n_ctx, intent_num=corpus.intent_num, model_num=corpus.model_num,
                      history_max=corpus.history_max, **cfg_kw)
    torch.manual_seed(seed)
    model = IntEL(argparse.Namespace(device=torch.device(DEV), model_path="", buffer=1), cfg=cfg).to(DEV)
    return corpus, cfg, model


C2 = dict(n_item=100_000, n_class=357, n_user=10_000, n_ctx=931, model_num=4, intent_num=1071, history_max=20)
PL = dict(encoder="GRU4Rec", context_emb_size=32, intent_emb_size=32, cross_attn_qsize=64, num_heads=2, num_layers=2)


def test_eval_sums_are_additive_over_shards_1m_sessions():
    """configs[1] size: 1M sessions x 50 candidates; sum over 4 shards == one pass; shuffling sessions changes nothing."""
    from intel_sigir2023_b200 import evaluate, synthetic
    N, L = 1_000_000, 50
    pred, ranking, pos, slen = synthetic.eval_set(N, L, 25, seed=3, device=DEV)
    topk = [3, 1, 5, 10]
    args = lambda sl: (pred[sl], ranking[sl], slen[sl], pos["c_paynum_i"][sl], pos["c_favnum_i"][sl], pos["c_clicknum_i"][sl], L, topk)
    full_s, full_c = evaluate.ndcg_sums(*args(slice(None)))
    parts = [evaluate.ndcg_sums(*args(slice(i * N // 4, (i + 1) * N // 4))) for i in range(4)]
    ps, pc = sum(p[0] for p in parts), sum(p[1] for p in parts)
    assert torch.equal(full_c, pc)
    assert torch.allclose(full_s, ps, rtol=1e-12, atol=0)
    perm = torch.randperm(N, device=DEV)
    sh_s, sh_c = evaluate.ndcg_sums(pred[perm], ranking[perm], slen[perm], pos["c_paynum_i"][perm], pos["c_favnum_i"][perm],
                                    pos["c_clicknum_i"][perm], L, topk)
    assert torch.equal(full_c, sh_c) and torch.allclose(full_s, sh_s, rtol=1e-12, atol=0)
    res = evaluate.metrics_from_sums(full_s.cpu().numpy(), full_c.cpu().numpy(), topk, ["NDCG", "HR"])
    assert 0.0 < res["NDCG@3"] < 1.0 and res["click_HR@10"] <= 1.0
    # a perfect scorer (score = gain) has NDCG@k = 1 for every k
    perfect = ranking.clamp(min=0).float() + 0.01 * torch.rand(N, L, device=DEV)
    valid = torch.arange(L, device=DEV)[None, :] < slen[:, None]
    s2, c2 = evaluate.ndcg_sums(perfect * valid, ranking, slen, pos["c_paynum_i"], pos["c_favnum_i"], pos["c_clicknum_i"], L, topk)
    r2 = evaluate.metrics_from_sums(s2.cpu().numpy(), c2.cpu().numpy(), topk, ["NDCG", "HR"])
    assert all(abs(r2[f"NDCG@{k}"] - 1.0) < 1e-12 for k in topk)


def test_eval_long_lists_config4_shape():
    """configs[3] shape: 200 candidates (a 100k-session slice); additivity + perfect scorer."""
    from intel_sigir2023_b200 import evaluate, synthetic
    N, L = 100_000, 200
    pred, ranking, pos, slen = synthetic.eval_set(N, L, 100, seed=5, device=DEV)
    topk = [3, 1, 5, 10]
    a, ac = evaluate.ndcg_sums(pred, ranking, slen, pos["c_paynum_i"], pos["c_favnum_i"], pos["c_clicknum_i"], L, topk)
    h = N // 2
    b1 = evaluate.ndcg_sums(pred[:h], ranking[:h], slen[:h], pos["c_paynum_i"][:h], pos["c_favnum_i"][:h], pos["c_clicknum_i"][:h], L, topk)
    b2 = evaluate.ndcg_sums(pred[h:], ranking[h:], slen[h:], pos["c_paynum_i"][h:], pos["c_favnum_i"][h:], pos["c_clicknum_i"][h:], L, topk)
    assert torch.allclose(a, b1[0] + b2[0], rtol=1e-12, atol=0) and torch.equal(ac, b1[1] + b2[1])


def test_forward_is_invariant_to_batch_split_full_step():
    """one bench-sized step (4096 sessions x 50 x K=4, I=1071): the two halves give the same rows as the whole."""
    from intel_sigir2023_b200 import synthetic
    corpus, cfg, model = _model(PL, C2)
    model.eval()
    batch = synthetic.make_batch(corpus, synthetic.BatchSpec(batch_size=4096, max_len=50, min_len=20), seed=1, device=DEV)
    with torch.no_grad():
        whole = model(batch)
        halves = [model(synthetic.shard_batch(batch, r, 2)) for r in range(2)]
    for k in ("weights", "ens_score", "intents"):
        cat = torch.cat([h[k] for h in halves], dim=0)
        assert torch.allclose(whole[k], cat, rtol=1e-6, atol=1e-7), k
    assert torch.isfinite(whole["ens_score"]).all()
    # pad slots carry score 0 -> ens 0; intents are distributions
    valid = torch.arange(50, device=DEV)[None, :] < batch["session_len"][:, None]
    assert (whole["ens_score"][~valid] == 0).all()
    assert torch.allclose(whole["intents"].sum(-1), torch.ones(4096, device=DEV), atol=1e-5)


def test_fusion_is_linear_in_the_weights():
    from intel_sigir2023_b200 import baselines
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.rand(4096, 50, 4, generator=g, device=DEV, dtype=torch.float64)
    w1 = torch.randn(4096, 50, 4, generator=g, device=DEV)
    w2 = torch.randn(4096, 50, 4, generator=g, device=DEV)
    lhs = baselines.fuse(2.0 * w1 + w2, x)
    rhs = 2.0 * baselines.fuse(w1, x) + baselines.fuse(w2, x)
    assert torch.allclose(lhs, rhs, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("kind", ["list", "bpr", "mse"])
def test_loss_gradient_matches_finite_difference_full_batch(kind):
    """directional derivative of the fused loss kernels at B=4096, L=50, K=4 (float32 central difference)."""
    from intel_sigir2023_b200 import losses, synthetic
    corpus = synthetic.CorpusSpec(**C2)
    batch = synthetic.make_batch(corpus, synthetic.BatchSpec(batch_size=4096, max_len=50, min_len=20), seed=2, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(1)
    ens = torch.randn(4096, 50, generator=g, device=DEV)
    w = 0.3 * torch.randn(4096, 50, 4, generator=g, device=DEV)
    cls = {"list": losses.Listloss, "bpr": losses.BPRloss, "mse": losses.MSEloss}[kind]
    crit = cls(argparse.Namespace(cal_diversity=1, diversity_alpha=0.05, intent_weight=0.1, ensemble_weight=1.0, kl_temp=2.0, kl_weight=0.5))
    crit.bpr_noise = torch.rand(4096, 50, 50, generator=g, device=DEV)

    def f(e, ww):
        return crit({"ens_score": e, "weights": ww}, batch)[0]
    e0, w0 = ens.clone().requires_grad_(True), w.clone().requires_grad_(True)
    f(e0, w0).backward()
    de, dw = torch.randn_like(ens), torch.randn_like(w)
    analytic = (e0.grad.double() * de.double()).sum() + (w0.grad.double() * dw.double()).sum()
    eps = 1e-2
    num = (f(ens + eps * de, w + eps * dw).double() - f(ens - eps * de, w - eps * dw).double()) / (2 * eps)
    assert abs(analytic.item() - num.item()) <= 2e-2 * abs(num.item()) + 1e-4, (analytic.item(), num.item())


def test_train_step_holds_no_reference_cycles_on_device_memory():
    """A ctx -> output -> grad_fn -> ctx cycle keeps a step's tensors alive until the cyclic GC runs; the caching
    allocator then has to cudaMalloc fresh segments inside the step (tens of ms of host stall).  With the GC off,
    steady-state steps must not grow the allocator at all."""
    import gc
    from intel_sigir2023_b200 import losses, synthetic
    corpus, cfg, model = _model(PL, C2)
    crit = losses.IntListloss(argparse.Namespace(cal_diversity=1, diversity_alpha=1e-6, intent_weight=0.001,
                                                 ensemble_weight=1.0, kl_weight=1.0, kl_temp=2.0))
    batch = synthetic.make_batch(corpus, synthetic.BatchSpec(batch_size=512, max_len=50, min_len=50), seed=1, device=DEV)

    def step():
        for p in model.parameters():
            p.grad = None
        out = model(batch)
        loss, _, _ = crit(out, batch)
        loss.backward()

    gc.collect()
    gc.disable()
    try:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        seg0 = torch.cuda.memory_stats()["segment.all.allocated"]
        live0 = torch.cuda.memory_allocated()
        for _ in range(12):
            step()
        torch.cuda.synchronize()
        assert torch.cuda.memory_stats()["segment.all.allocated"] == seg0
        assert torch.cuda.memory_allocated() == live0
    finally:
        gc.enable()


def test_device_prefetcher_yields_the_same_batches_in_order():
    """loader.DevicePrefetcher (H2D of batch i+1 overlapped with step i) hands over exactly the tensors a synchronous
    utils.batch_to_gpu would, in order, and training on them gives the same loss."""
    from intel_sigir2023_b200 import loader, losses, synthetic
    corpus, cfg, model = _model(PL, C2)
    crit = losses.IntListloss(argparse.Namespace(cal_diversity=1, diversity_alpha=1e-6, intent_weight=0.001,
                                                 ensemble_weight=1.0, kl_weight=1.0, kl_temp=2.0))
    host = [loader.pin_batch(synthetic.make_batch(corpus, synthetic.BatchSpec(batch_size=256, max_len=50, min_len=3), seed=s))
            for s in range(5)]
    direct = []
    for hb in host:
        b = synthetic.batch_to(hb, DEV)
        direct.append(float(crit(model(b), b)[0]))
    staged = []
    for hb, b in zip(host, loader.DevicePrefetcher(iter(host), DEV)):
        for k, v in hb.items():
            if torch.is_tensor(v):
                assert torch.equal(b[k].cpu(), v), k
        staged.append(float(crit(model(b), b)[0]))
    assert staged == direct


def test_host_packed_history_gives_the_same_step():
    """DevicePrefetcher(pack_history=True) ships only the non-zeros of the dense float64 history tensors; the model's
    outputs and the loss match the dense path (fp32 summation order of the sparse rows aside)."""
    from intel_sigir2023_b200 import loader, losses, synthetic
    corpus, cfg, model = _model(PL, C2)
    crit = losses.IntListloss(argparse.Namespace(cal_diversity=1, diversity_alpha=1e-6, intent_weight=0.001,
                                                 ensemble_weight=1.0, kl_weight=1.0, kl_temp=2.0))
    host = [loader.pin_batch(synthetic.make_batch(corpus, synthetic.BatchSpec(batch_size=192, max_len=50, min_len=3), seed=s))
            for s in range(4)]
    dense = [(model(b), b) for b in (synthetic.batch_to(hb, DEV) for hb in host)]
    # a tiny initial capacity forces the grow-and-retry path
    for (out_d, bd), bp in zip(dense, loader.DevicePrefetcher(iter(host), DEV, pack_history=True, pack_nz=2)):
        assert "his_intents" not in bp and bp["his_intents_idx"].dtype == torch.int32
        out_p = model(bp)
        for k in ("intents", "weights", "ens_score"):
            a, c = out_d[k].detach().cpu().numpy(), out_p[k].detach().cpu().numpy()
            assert np.abs(a - c).max() <= 2e-6 * max(1.0, np.abs(a).max()), k
        assert abs(float(crit(out_d, bd)[0]) - float(crit(out_p, bp)[0])) < 1e-5


@pytest.mark.parametrize("encoder", ["BERT4Rec", "GRU4Rec"])
def test_full_batch_step_tcgen05_vs_mma_sync(encoder):
    """At B = 4096 most nn.Linear passes are large enough for the tcgen05 / TMEM GEMM (bias, residual, relu and mask
    epilogues, split-K weight gradients).  The same step with the large GEMMs forced onto the mma.sync kernels must
    give the same outputs, loss and gradients."""
    from intel_sigir2023_b200 import _lib, losses, synthetic
    kw = dict(PL, encoder=encoder)
    corpus, cfg, model = _model(kw, C2)
    crit = losses.IntListloss(argparse.Namespace(cal_diversity=1, diversity_alpha=1e-6, intent_weight=0.001,
                                                 ensemble_weight=1.0, kl_weight=1.0, kl_temp=2.0))
    batch = synthetic.make_batch(corpus, synthetic.BatchSpec(batch_size=4096, max_len=50, min_len=5), seed=2, device=DEV)
    res = []
    for on in (1, 0):
        _lib.check(_lib.load().intel_debug_use_tcgen05_gemm(on))
        try:
            for p in model.parameters():
                p.grad = None
            out = model(batch)
            loss = crit(out, batch)[0]
            loss.backward()
            res.append((out, float(loss), {k: p.grad.clone() for k, p in model.named_parameters()}))
        finally:
            _lib.check(_lib.load().intel_debug_use_tcgen05_gemm(1))
    (oa, la, ga), (ob, lb, gb) = res
    assert abs(la - lb) <= 1e-6 * abs(lb)
    for k in ("intents", "weights", "ens_score"):
        a, b = oa[k].detach().cpu().numpy(), ob[k].detach().cpu().numpy()
        assert np.abs(a - b).max() <= 5e-6 * max(1.0, np.abs(b).max()), k
    gmax = max(float(g.abs().max()) for g in gb.values())
    for k, g in gb.items():
        d = float((ga[k] - g).abs().max())
        assert d <= 2e-4 * float(g.abs().max()) + 3e-6 * gmax, (k, d)


@pytest.mark.parametrize("kind,encoder", [("list", "GRU4Rec"), ("bpr", "GRU4Rec"), ("mse", "BERT4Rec")])
def test_c2_shapes_against_the_oracle(kind, encoder):
    """BASELINE.json configs[1] shapes (L = 50, K = 4, I = 1071, H = 20, the script's model flags) at B = 1024 - large
    enough for the tcgen05 GEMMs and for every CTA-level path of the fused kernels - against the CPU oracle: outputs,
    loss and all parameter gradients."""
    import parity_checks as P
    kw = dict(PL, encoder=encoder)
    P.check_against_oracle(DEV, seed=11, B=1024, L=50, min_len=4, kind=kind,
                           corpus_kw=dict(n_item=5000, n_class=357, n_user=2000, n_ctx=931, model_num=4, intent_num=1071,
                                          history_max=20), **kw)
