"""Achieved parity margins of the fused PredictorPlus / Predictor train step against the reference's golden outputs
(tests/golden/*.npz): per dataset and model variant the relative error of the loss and, per parameter, the largest
gradient error as a fraction of the gradient's scale and the share of elements outside 1e-3 of the scale.
usage (GPU box): python scripts/parity_margins.py > profiles/r2_parity_margins.txt"""
import sys, os, tempfile, pathlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests import _golden as G
from tests.test_gpu_plus import make_kg, make_model

tmp = pathlib.Path(tempfile.mkdtemp())
print("variant: type/aggregator/entity_feature | loss rel.err (worst batch) | worst parameter: max|err|/scale, share of elements > 1e-3*scale")
for name in G.DATASETS:
    fx = G.load(name)
    kg = make_kg(fx)
    for tag in G.plus_tags(fx):
        cfg = G.plus_cfg(fx, tag)
        m, _ = make_model(fx, kg, tag, tmp)
        worst_loss, worst = 0.0, ("-", 0.0, 0.0)
        for j in range(3):
            if "%s_tb%d_loss" % (tag, j) not in fx:
                continue
            tri, _, _ = G.train_batch_inputs(fx, j)
            m.zero_grad()
            loss, _ = m.fused_train_step([[tuple(x) for x in tri.tolist()]], 0.2)
            want = float(fx["%s_tb%d_loss" % (tag, j)])
            worst_loss = max(worst_loss, abs(loss[0].item() - want) / max(abs(want), 1e-12))
            for pn, par in m.named_parameters():
                key = "%s_tb%d_g_%s" % (tag, j, pn)
                if key in fx and par.grad is not None:
                    w = fx[key]
                    scale = max(1e-6, float(np.abs(w).max()))          # gradients that are zero in exact arithmetic (b2 without an entity feature) stay at rounding level
                    err = np.abs(par.grad.cpu().numpy() - w)
                    if err.max() / scale > worst[1]:
                        worst = (pn, float(err.max() / scale), float((err > 1e-3 * scale).mean()))
        print("%-8s %-6s %s/%s/%s | loss %.2e | %s: %.2e, %.1f%%" % (name, tag, cfg["type"], cfg["aggregator"], cfg["entity_feature"],
                                                               worst_loss, worst[0], worst[1], 100 * worst[2]))
