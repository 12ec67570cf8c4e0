"""Autograd glue of the RotatE epilogue kernels (rl_rotate.cu)."""
import torch

from . import _lib
from .engine import _stream

LANES = _lib.LANES


class _RotateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eemb, remb, gamma, sk, sl):
        D = remb.shape[1]
        e = eemb.detach().contiguous().float()
        r = remb.detach().contiguous().float()
        P = torch.empty(sl.S, 2 * D, LANES, dtype=torch.float32, device=e.device)
        out = torch.empty(sl.S, sk.N, LANES, dtype=torch.float32, device=e.device)
        _lib.check(_lib.lib().rl_rotate_scores(sk.dg.ref(), sl.ref(), D, float(gamma), e.data_ptr(), r.data_ptr(),
                                               P.data_ptr(), out.data_ptr(), _stream()), "rl_rotate_scores")
        ctx.sk, ctx.sl, ctx.gamma, ctx.D = sk, sl, float(gamma), D
        ctx.save_for_backward(e, r, P)
        return out

    @staticmethod
    def backward(ctx, G):
        e, r, P = ctx.saved_tensors
        sk, sl = ctx.sk, ctx.sl
        G = torch.nan_to_num(G.contiguous().float(), nan=0.0)
        dP = torch.zeros_like(P)
        de = torch.zeros_like(e)
        dr = torch.zeros_like(r)
        _lib.check(_lib.lib().rl_rotate_backward(sk.dg.ref(), sl.ref(), ctx.D, ctx.gamma, e.data_ptr(), r.data_ptr(),
                                                 P.data_ptr(), G.data_ptr(), dP.data_ptr(), de.data_ptr(), dr.data_ptr(),
                                                 _stream()), "rl_rotate_backward")
        return de, dr, None, None, None


def rotate_slot_scores(mod, sk, sl):
    if mod.eemb.shape[0] != sk.N:
        raise ValueError("RotatE table has %d entities, graph has %d" % (mod.eemb.shape[0], sk.N))
    return _RotateFn.apply(mod.eemb, mod.remb, mod.gamma, sk, sl)


class _MiniDriver:
    """Just enough of ScoreKernels / Slots for the RotatE kernels (they only read num_entities,
    slot_head and lane_h): lets RotatE.forward run without a KnowledgeGraph."""

    def __init__(self, N, device, heads, lane_h):
        self.N, self.device = N, device
        self.struct_g = _lib.RlGraph()
        self.struct_g.num_entities = N
        self.S = int(heads.shape[0])
        self.heads_t, self.lane_t = heads, lane_h
        self.struct_s = _lib.RlSlots(self.S, heads.data_ptr(), lane_h.data_ptr(), None, None, None, None, None, None)

    class _Ref:
        def __init__(self, st):
            self.st = st

        def ref(self):
            import ctypes
            return ctypes.byref(self.st)

    @property
    def dg(self):
        return self._Ref(self.struct_g)

    def ref(self):
        import ctypes
        return ctypes.byref(self.struct_s)


def rotate_dense_scores(mod, all_h, all_r):
    """RotatE.forward(all_h, all_r) -> fp32[B,N] (embedding.py:64-70); any relation per query."""
    _lib.require_cuda(all_h, "all_h")
    dev = all_h.device
    B, N = int(all_h.shape[0]), int(mod.eemb.shape[0])
    # one slot per distinct relation chunk of <= 32 queries
    r_host = all_r.detach().cpu().tolist()
    order = sorted(range(B), key=lambda i: r_host[i])
    heads, lanes, where = [], [], []
    i = 0
    while i < B:
        j = i
        while j < B and j - i < LANES and r_host[order[j]] == r_host[order[i]]:
            j += 1
        heads.append(r_host[order[i]])
        lanes.append(order[i:j] + [-1] * (LANES - (j - i)))
        where += [(len(heads) - 1, k) for k in range(j - i)]
        i = j
    idx = torch.tensor(lanes, dtype=torch.long, device=dev)
    lane_h = torch.where(idx >= 0, all_h[idx.clamp(min=0)], torch.full_like(idx, -1)).to(torch.int32).contiguous()
    heads_t = torch.tensor(heads, dtype=torch.int32, device=dev)
    drv = _MiniDriver(N, dev, heads_t, lane_h.view(-1))
    out = _RotateFn.apply(mod.eemb, mod.remb, mod.gamma, drv, drv)        # [S][N][32]
    slot = torch.tensor([w[0] for w in where], device=dev)
    lane = torch.tensor([w[1] for w in where], device=dev)
    rows = out[slot, :, lane]                                             # queries in sorted order
    inv = torch.empty(B, dtype=torch.long, device=dev)
    inv[torch.tensor(order, device=dev)] = torch.arange(B, device=dev)
    return rows[inv]
