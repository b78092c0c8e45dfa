"""One warm train step + one eval step (+ one device-built batch) of the C2 workload for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:<kernels> -s <skip> -c <n> -o gpurun_out/x python profiles/tools/prof_step.py"""
import argparse, sys, os, torch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from intel_sigir2023_b200 import synthetic, losses, evaluate, corpus as corpus_mod
from intel_sigir2023_b200.IntEL import IntEL
sys.argv = [sys.argv[0]]
a = bench.parse()
corpus, cfg, loss_kind, loss_args = bench.make_cfg(a)
dev = torch.device("cuda")
torch.manual_seed(0)
model = IntEL(argparse.Namespace(device=dev, model_path="", buffer=1), cfg=cfg).to(dev)
crit = losses.IntListloss(loss_args)
batch = synthetic.make_batch(corpus, synthetic.BatchSpec(batch_size=a.batch, max_len=a.list_len, min_len=a.list_len), seed=0, device=dev)
for i in range(int(os.environ.get("STEPS", "2"))):
    for p in model.parameters(): p.grad = None
    out = model(batch); loss, _, _ = crit(out, batch); loss.backward()
if os.environ.get("EVAL", "1") == "1":
    model.eval()
    with torch.no_grad():
        out = model(batch)
    evaluate.ndcg_sums(out["ens_score"], batch["ranking"], batch["session_len"], batch["c_paynum_i"], batch["c_favnum_i"],
                       batch["c_clicknum_i"], max(a.list_len, 10), [3, 1, 5, 10])
    cols, shared, nz = corpus_mod.synthetic_columns(4 * a.batch, a.list_len, a.n_item, 357, a.n_user, a.n_ctx, a.model_num, a.intent_num, seed=3)
    dc = corpus_mod.DeviceCorpus(cols, shared, a.model_num, a.intent_num, 20, nz, dev)
    b2 = dc.batch(np.arange(a.batch))
torch.cuda.synchronize()
