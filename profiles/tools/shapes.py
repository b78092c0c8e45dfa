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
def step():
    for p in model.parameters(): p.grad = None
    out = model(batch); loss, _, _ = crit(out, batch); loss.backward()
for i in range(3): step()
torch.cuda.synchronize()
_lib.profile(2)
N = 5
for i in range(N): step()
rep = _lib.profile_report()
_lib.profile(0)
tot = sum(v["ms"] for v in rep.values())
print("total %.3f ms/step" % (tot / N))
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    ms = v["ms"] / N
    print("%-60s n=%4.1f  %7.3f ms  %5.1f%%  %7.1f GB/s %8.1f GF/s" % (k, v["launches"] / N, ms, 100 * v["ms"] / tot,
          v["bytes"] / v["ms"] / 1e6 if v["ms"] else 0, v["flops"] / v["ms"] / 1e6 if v["ms"] else 0))
