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


def rotate_dense_scores(mod, all_h, all_r):
    """RotatE.forward(all_h, all_r) -> fp32[B,N]; one relation per call like the predictors."""
    raise NotImplementedError("use PredictorPlus(entity_feature='RotatE'); the stand-alone dense entry point "
                              "needs a graph-bound driver")
