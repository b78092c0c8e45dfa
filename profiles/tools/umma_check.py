import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from intel_sigir2023_b200 import _lib
from conftest import rel_err
lib = _lib.load()
dev = "cuda"
g = torch.Generator().manual_seed(2)
def run(M, N, K, what):
    st = _lib.stream_ptr(torch.device(dev))
    if what == "fwd":
        X = torch.randn(M, K, generator=g).to(dev); W = torch.randn(N, K, generator=g).to(dev); b = torch.randn(N, generator=g).to(dev)
        Y = torch.empty(M, N, device=dev)
        _lib.check(lib.intel_linear_fwd(M, N, K, _lib.ptr(X), _lib.ptr(W), _lib.ptr(b), _lib.ptr(Y), st))
        ref = X.double().cpu() @ W.double().cpu().t() + b.double().cpu()
        return rel_err(Y.double().cpu().numpy(), ref.numpy())
    if what == "dx":
        dY = torch.randn(M, N, generator=g).to(dev); W = torch.randn(N, K, generator=g).to(dev); U = torch.randn(M, K, generator=g).to(dev)
        dX = torch.empty(M, K, device=dev)
        _lib.check(lib.intel_linear_dx(M, N, K, _lib.ptr(dY), _lib.ptr(W), _lib.ptr(dX), _lib.ptr(U), st))
        ref = (dY.double().cpu() @ W.double().cpu()) * (U.cpu() > 0)
        return rel_err(dX.double().cpu().numpy(), ref.numpy())
    if what == "dw":
        dY = torch.randn(M, N, generator=g).to(dev); X = torch.randn(M, K, generator=g).to(dev)
        dW = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
        _lib.check(lib.intel_linear_dw(M, N, K, _lib.ptr(dY), _lib.ptr(X), _lib.ptr(dW), _lib.ptr(db), st))
        ref = dY.double().cpu().t() @ X.double().cpu()
        return rel_err(dW.double().cpu().numpy(), ref.numpy())
import time
def bench(M, N, K, what, iters=10):
    run(M, N, K, what); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); Y = torch.empty(M, N, device=dev); dY = torch.randn(M, N, device=dev)
    dX = torch.empty(M, K, device=dev); dW = torch.zeros(N, K, device=dev)
    st = _lib.stream_ptr(torch.device(dev))
    e0.record()
    for _ in range(iters):
        if what == "fwd": lib.intel_linear_fwd(M, N, K, _lib.ptr(X), _lib.ptr(W), None, _lib.ptr(Y), st)
        elif what == "dx": lib.intel_linear_dx(M, N, K, _lib.ptr(dY), _lib.ptr(W), _lib.ptr(dX), None, st)
        else: lib.intel_linear_dw(M, N, K, _lib.ptr(dY), _lib.ptr(X), _lib.ptr(dW), None, st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (M, N, K) in [(4096, 128, 256), (4096, 384, 64), (5000, 48, 384), (8192, 100, 176), (20000, 384, 48), (4099, 72, 1000), (4096, 1071, 32), (4096, 32, 1071), (4096, 1071, 176), (81920, 384, 64), (86016, 384, 128)]:
    for what in ("fwd", "dx", "dw"):
        errs = []
        for on in (1, 0):
            lib.intel_debug_use_tcgen05_gemm(on)
            errs.append(run(M, N, K, what))
        torch.cuda.synchronize()
        ts = []
        for on in (1, 0):
            lib.intel_debug_use_tcgen05_gemm(on)
            ts.append(bench(M, N, K, what))
        print((M, N, K), what, "umma err %.3e   mma.sync err %.3e   | umma %.3f ms  mma.sync %.3f ms" % (errs[0], errs[1], ts[0], ts[1]), flush=True)
lib.intel_debug_use_tcgen05_gemm(1)
