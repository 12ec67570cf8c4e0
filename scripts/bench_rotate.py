"""Side benchmark / profiling target: RotatE epilogue kernels at the FB15k-237 shape (D = 1000, 64 slots)."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from rnnlogic_b200 import _lib
from rnnlogic_b200.rotate import _RotateFn, _MiniDriver
N, R, D, S, gamma = 14541, 474, 1000, 64, 9.0
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
rr = (gamma + 2.0) / D
eemb = ((torch.rand(N, 2 * D, generator=g) * 2 - 1) * rr).to(dev).requires_grad_()
remb = ((torch.rand(R, D, generator=g) * 2 - 1) * rr).to(dev).requires_grad_()
heads = torch.randint(R, (S,), generator=g).to(torch.int32).to(dev)
lane_h = torch.randint(N, (S * 32,), generator=g).to(torch.int32).to(dev)
drv = _MiniDriver(N, dev, heads, lane_h)
G = torch.randn(S, N, 32, device=dev) * 1e-3
def step():
    out = _RotateFn.apply(eemb, remb, gamma, drv, drv)
    out.backward(G)
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); out = _RotateFn.apply(eemb, remb, gamma, drv, drv); e1.record(); out.backward(G); e2.record(); torch.cuda.synchronize()
print("rotate fwd ms %.3f  bwd ms %.3f  (S=%d slots, N=%d, D=%d)" % (e0.elapsed_time(e1), e1.elapsed_time(e2), S, N, D))
