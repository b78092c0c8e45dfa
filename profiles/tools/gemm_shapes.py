import argparse, sys, os, torch, json
sys.path.insert(0, "/root/repo")
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
_lib.profile(2)
for i in range(3): step()
prof = _lib.profile_report()
_lib.profile(False)
tot = sum(v["ms"] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:60]:
    print(f"{v['ms']/3:8.3f} ms  n={v['launches']/3:4.0f}  {v['bytes']/max(v['ms'],1e-9)/1e6:8.0f} GB/s  {k}")
print("total", tot/3)
