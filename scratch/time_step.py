import argparse, sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from intel_sigir2023_b200 import synthetic, losses, _lib
from intel_sigir2023_b200.IntEL import IntEL
a = bench.parse()
corpus, cfg, loss_kind, loss_args = bench.make_cfg(a)
dev = torch.device("cuda")
torch.manual_seed(0)
model = IntEL(argparse.Namespace(device=dev, model_path="", buffer=1), cfg=cfg).to(dev)
crit = losses.IntListloss(loss_args)
batch = synthetic.make_batch(corpus, synthetic.BatchSpec(batch_size=a.batch, max_len=a.list_len, min_len=a.list_len), seed=1, device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
import gc
if os.environ.get('NOGC'): gc.disable()
if os.environ.get('FREEZE'): gc.collect(); gc.freeze()
for it in range(16):
    for p in model.parameters(): p.grad = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e = [ev() for _ in range(4)]
    e[0].record()
    out = model(batch)
    e[1].record()
    loss, _, _ = crit(out, batch)
    e[2].record()
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    e[3].record()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"it{it}: fwd {e[0].elapsed_time(e[1]):.2f} loss {e[1].elapsed_time(e[2]):.2f} bwd {e[2].elapsed_time(e[3]):.2f} ms | host: fwd+loss enqueue {1e3*(t1-t0):.2f} bwd enqueue {1e3*(t2-t1):.2f} total wall {1e3*(t3-t0):.2f}")
