"""LambdaRank's per-item lambdas on the B200 path: `compute_lambda_new` of
helpers/LambdaRankRunner.py:315-344, the O(L^2) pairwise pass of every LambdaRank training step (:246), as one kernel
(`intel_lambdarank_lambdas`, one warp per session) instead of ten [B,L,L] temporaries."""
from __future__ import annotations

import torch

from . import _lib


def compute_lambda_new(true_scores: torch.Tensor, temp_scores: torch.Tensor, session_len: torch.Tensor) -> torch.Tensor:
    """Same arguments as the reference method (without `self`): `true_scores` int64 [B,L] = clamp(batch['ranking'], min=0)
    (the raw ranking works too, the kernel clamps), `temp_scores` float32 [B,L] = the detached `ens_score`,
    `session_len` int64 [B] -> Lambda float32 [B,L], to be fed to `predicted_scores.backward(Lambda)` (:259)."""
    lib = _lib.load()
    B, L = true_scores.shape
    ranking = true_scores.contiguous()
    scores = temp_scores.detach().contiguous()
    lens = session_len.contiguous()
    lambdas = torch.empty(B, L, dtype=torch.float32, device=scores.device)
    _lib.check(lib.intel_lambdarank_lambdas(B, L, _lib.ptr(ranking, torch.int64), _lib.ptr(scores, torch.float32),
                                            _lib.ptr(lens, torch.int64), _lib.ptr(lambdas), _lib.stream_ptr(scores.device)))
    return lambdas


class _ScorerFn(torch.autograd.Function):
    """(iid table, W0, b0, W1, b1, ...) -> ens_score [B,L]: gather + MLP + softmax over the (padded) list through
    intel_gather_fwd / intel_linear_*_ex / intel_softmax_rows_*; pre-activations are kept for the backward pass."""

    @staticmethod
    def forward(ctx, batch, table, *wb):
        lib = _lib.load()
        items, scores = batch["i_id_s"], batch["scores"]
        B, L, K = scores.shape
        R, d = B * L, table.shape[1]
        F = d + 1 + K
        dev = table.device
        st = _lib.stream_ptr(dev)
        X = torch.empty(R, F, dtype=torch.float32, device=dev)
        idx = items.reshape(-1).contiguous()
        _lib.check(lib.intel_gather_fwd(R, d, _lib.ptr(table), _lib.ptr(idx, torch.int64), _lib.ptr(X), F, 0, st))
        X[:, d] = batch["i_class_c"].reshape(-1)                 # the category id itself is a feature (LambdaRank.py:43)
        X[:, d + 1:] = scores.reshape(R, K)
        acts = [X]
        for li in range(len(wb) // 2):
            W, bias = wb[2 * li], wb[2 * li + 1]
            A = acts[-1]
            Z = torch.empty(R, W.shape[0], dtype=torch.float32, device=dev)
            _lib.check(lib.intel_linear_fwd_ex(R, W.shape[0], W.shape[1], _lib.ptr(A), A.shape[1], _lib.ptr(W), _lib.ptr(bias),
                                               _lib.ptr(Z), W.shape[0], 1 if li > 0 else 0, st))
            acts.append(Z)
        logits = acts.pop().view(B, L)
        ens = torch.empty(B, L, dtype=torch.float32, device=dev)
        _lib.check(lib.intel_softmax_rows_fwd(B, L, _lib.ptr(logits), _lib.ptr(ens), st))
        ctx.acts, ctx.idx, ctx.params = acts, idx, (table,) + tuple(wb)
        ctx.save_for_backward(ens)
        return ens

    @staticmethod
    def backward(ctx, d_ens):
        lib = _lib.load()
        (ens,) = ctx.saved_tensors
        table, wb = ctx.params[0], ctx.params[1:]
        acts, idx = ctx.acts, ctx.idx
        B, L = ens.shape
        R, d = B * L, table.shape[1]
        dev = table.device
        st = _lib.stream_ptr(dev)
        dZ = torch.empty(B, L, dtype=torch.float32, device=dev)
        _lib.check(lib.intel_softmax_rows_bwd(B, L, _lib.ptr(ens), _lib.ptr(d_ens.contiguous(), torch.float32), _lib.ptr(dZ), st))
        dZ = dZ.view(R, 1)
        grads = [None] * len(wb)
        for li in reversed(range(len(wb) // 2)):
            W = wb[2 * li]
            A = acts[li]                                            # input of layer li: X, or the pre-activation of layer li-1
            gW, gb = torch.zeros_like(W), torch.zeros_like(wb[2 * li + 1])
            _lib.check(lib.intel_linear_dw_ex(R, W.shape[0], W.shape[1], _lib.ptr(dZ), _lib.ptr(A), A.shape[1], _lib.ptr(gW),
                                              _lib.ptr(gb), 1 if li > 0 else 0, st))
            grads[2 * li], grads[2 * li + 1] = gW, gb
            dA = torch.empty_like(A)
            _lib.check(lib.intel_linear_dx_ex(R, W.shape[0], W.shape[1], _lib.ptr(dZ), _lib.ptr(W), _lib.ptr(dA), A.shape[1],
                                              _lib.ptr(A) if li > 0 else None, A.shape[1], st))
            dZ = dA
        g_table = torch.zeros_like(table)
        _lib.check(lib.intel_scatter_add_bwd(R, d, _lib.ptr(dZ), dZ.shape[1], _lib.ptr(idx), _lib.ptr(g_table), st))
        ctx.acts = ctx.idx = ctx.params = None
        return (None, g_table) + tuple(grads)


class LambdaRank(torch.nn.Module):
    """The LambdaRank scorer (models/supervise/LambdaRank.py:15-47; script/baselines.sh): an MLP over
    [E_iid[item] || category id || the K basic scores] per list slot, softmax over the padded list.  Same parameter names
    (`iid_embeddings.weight`, `mlp.Linear-<i>.{weight,bias}`), flags (`--hidden_size` comma list, `--i_emb_size`) and output
    dict (`weights` is all zeros, as in the reference); trained by `compute_lambda_new` + `ens_score.backward(lambdas)`."""
    reader, runner = "BaseReader", "LambdaRankRunner"
    extra_log_args: list = []

    @staticmethod
    def parse_model_args(parser):
        parser.add_argument('--hidden_size', type=str, default='32')
        parser.add_argument('--i_emb_size', type=int, default=32, help='Embedding size for item id.')
        parser.add_argument('--model_path', type=str, default='', help='Model save path.')
        parser.add_argument('--buffer', type=int, default=1, help='Whether to buffer feed dicts for dev/test')
        parser.add_argument('--model_num', type=int, default=2, help='Number of base models.')
        return parser

    def __init__(self, args, corpus=None, item_num: int = None):
        super().__init__()
        nn = torch.nn
        self.device = getattr(args, "device", None)
        self.item_num = int(item_num if item_num is not None else corpus.max_iid + 1)
        self.iid_embeddings = nn.Embedding(self.item_num, args.i_emb_size)
        sizes = [args.model_num + args.i_emb_size + 1] + [int(x) for x in str(args.hidden_size).split(',')]
        self.hidden_sizes = sizes
        self.mlp = nn.Sequential()
        for i in range(len(sizes) - 1):
            self.mlp.add_module('Linear-%d' % i, nn.Linear(sizes[i], sizes[i + 1]))
            self.mlp.add_module('Activate-%d' % i, nn.ReLU())
        self.mlp.add_module('Linear-%d' % (len(sizes) - 1), nn.Linear(sizes[-1], 1))

    def customize_parameters(self, define_dict=None) -> list:
        from .optim import customize_parameters
        return customize_parameters(self)

    def forward(self, data):
        wb = [p for n, p in self.mlp.named_parameters()]
        ens = _ScorerFn.apply(data, self.iid_embeddings.weight, *wb)
        return {"weights": torch.zeros(data["scores"].shape, dtype=torch.float32, device=ens.device), "ens_score": ens}
