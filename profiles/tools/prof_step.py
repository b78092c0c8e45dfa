import argparse, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from intel_sigir2023_b200 import synthetic, losses, _lib
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
torch.cuda.synchronize()
