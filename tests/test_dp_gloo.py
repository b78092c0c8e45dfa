"""World-size-2 data-parallel path on CPU (gloo): session sharding + gradient all-reduce + sharded
evaluation must reproduce the single-process result.  The compute runs through the product's host layer on
the CUDA emulator build of the kernels (tests/emu), so the N>1 host logic is what is covered here."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "emu"))


def _setup_emu():
    from build_emu import build_emu
    from intel_sigir2023_b200 import _lib
    _lib._lib = None
    _lib.load(build_emu())
    _lib._allow_host_tensors = True


def _grads_and_metrics(rank, world, case, overlap=True):
    import parity_checks as P
    from conftest import load_model_case
    from intel_sigir2023_b200 import dp, losses, synthetic
    cfg, batch, state, gold = load_model_case(case)
    B = batch["batch_size"]
    keep = B - (B % 2)                      # equal shards
    batch = {k: (v[:keep] if torch.is_tensor(v) else v) for k, v in batch.items()}
    batch["batch_size"] = keep
    shard = synthetic.shard_batch(batch, rank, world)
    model = P.make_model(cfg, state, "cpu").train()
    crit = losses.IntListloss(P.loss_args())
    # overlap: the reducer hooks into the backward pass and exchanges everything but the score stream's gradients while
    # the score stack is still being differentiated; the late reducer (old call order) takes the one-collective path
    reducer = dp.GradReducer(model, world, overlap=overlap) if overlap else None
    out = model(shard)
    loss, _, _ = crit(out, shard)
    loss.backward()
    if overlap and world > 1:
        assert reducer._pending is not None and 0 < model._flat_late < model._flat_grad.numel()
    (reducer or dp.GradReducer(model, world, overlap=False)).allreduce()
    pos = {k: shard[k] for k in ("c_paynum_i", "c_favnum_i", "c_clicknum_i")}
    res = dp.evaluate_sharded(out["ens_score"].detach(), shard["ranking"], pos, shard["session_len"], [3, 1, 5], ["NDCG", "HR"])
    return {n: p.grad.clone() for n, p in model.named_parameters()}, res


def _worker(rank, world, port, case, q, overlap=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    _setup_emu()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grads, res = _grads_and_metrics(rank, world, case, overlap)
    if rank == 0:
        q.put(({k: v.numpy() for k, v in grads.items()}, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,overlap", [("script_bpr_gru_k4", True), ("default_bert", True), ("script_bpr_gru_k4", False)])
def test_two_rank_grads_match_single_process(case, overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    grads2, res2 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    _setup_emu()
    try:
        grads1, res1 = _grads_and_metrics(0, 1, case)
    finally:
        from intel_sigir2023_b200 import _lib
        _lib._lib, _lib._allow_host_tensors = None, False
    gmax = max(float(v.abs().max()) for v in grads1.values())
    for k, g in grads1.items():
        err = np.abs(g.numpy() - grads2[k]).max()
        assert err <= 1e-4 * float(g.abs().max()) + 2e-6 * gmax, (k, err)
    for k, v in res1.items():
        assert (np.isnan(v) and np.isnan(res2[k])) or abs(v - res2[k]) < 1e-9, (k, v, res2[k])


def test_overlap_refuses_gradient_accumulation():
    """The early all-reduce works on the flat gradient buffer of the backward pass: if the .grad tensors do not alias it
    (gradients accumulated into existing .grad), allreduce() must say so instead of exchanging the wrong memory."""
    from intel_sigir2023_b200 import dp

    class Work:
        waited = False
        def wait(self):
            Work.waited = True

    lin = torch.nn.Linear(3, 2)
    lin._early_reduce = None
    red = dp.GradReducer.__new__(dp.GradReducer)
    red.model, red.params, red.world, red.small_numel = lin, list(lin.parameters()), 2, 1 << 16
    red.avg_op, red.post_scale, red.overlap = None, 1.0, True
    flat = torch.zeros(16)
    for p in lin.parameters():
        p.grad = torch.zeros_like(p)                    # own storage: not views of `flat`
    red._pending = (Work(), flat, 8)
    with pytest.raises(RuntimeError, match="do not alias"):
        red.allreduce()
    assert Work.waited and red._pending is None
