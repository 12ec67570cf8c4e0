"""Rule-encoder LSTM (hidden 16, 3 layers, T=4): rl_rnn.cu vs cuDNN, forward + backward."""
import sys, torch
sys.path.insert(0, '/root/repo')
from rnnlogic_b200.predictors import _LstmEncodeFn
torch.backends.cudnn.allow_tf32 = False
H, L, T, V = 16, 3, 4, 474
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
emb = torch.nn.Embedding(V + 1, H, padding_idx=V).cuda()
rnn = torch.nn.LSTM(H, H, L, batch_first=True).cuda()
lens = torch.randint(2, T + 1, (n,), device="cuda")
tok = torch.randint(0, V, (n, T), device="cuda")
tok[torch.arange(T, device="cuda")[None, :] >= lens[:, None]] = V
proj = torch.randn(n, H, device="cuda")
ws = [getattr(rnn, "%s_l%d" % (nm, l)) for l in range(L) for nm in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
def run(fused):
    x = emb(tok)
    if fused:
        out = _LstmEncodeFn.apply(x, lens.to(torch.int32), L, *ws)
    else:
        o, _ = rnn(x)
        out = torch.gather(o, 1, (lens - 1).view(-1, 1, 1).expand(-1, -1, H)).squeeze(1)
    (out * proj).sum().backward()
def t(fn, k=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print("n=%d  rl_rnn.cu %.3f ms   cuDNN %.3f ms" % (n, t(lambda: run(True)), t(lambda: run(False))))
