"""End-to-end: the B200 TrainerPredictor reproduces a short run of the reference trainer
(20 Adam steps of PredictorPlus(emb,sum,bias) + filtered valid MRR; tests/golden/make_golden.py (f)),
and the reference-format dataset directory / rule file / YAML config load unchanged."""
import os
import random

import numpy as np
import pytest
import torch

from tests import _golden as G

gpu = pytest.mark.gpu


def write_dataset_dir(path, fx):
    N, R = int(fx["N"]), int(fx["R"])
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "entities.dict"), "w") as f:
        f.write("".join("%d\te%d\n" % (i, i) for i in range(N)))
    with open(os.path.join(path, "relations.dict"), "w") as f:
        f.write("".join("%d\tr%d\n" % (i, i) for i in range(R)))
    for split in ("train", "valid", "test"):
        with open(os.path.join(path, split + ".txt"), "w") as f:
            f.write("".join("e%d\tr%d\te%d\n" % tuple(x) for x in fx[split].tolist()))


@gpu
@pytest.mark.parametrize("name", ["umls", "syn"])
def test_short_training_run_matches_reference(name, tmp_path):
    from rnnlogic_b200.data import KnowledgeGraph, TrainDataset, ValidDataset, TestDataset
    from rnnlogic_b200.predictors import PredictorPlus
    from rnnlogic_b200.trainer import TrainerPredictor
    from rnnlogic_b200.utils import set_seed
    fx = G.load(name)
    d = str(tmp_path / name)
    write_dataset_dir(d, fx)
    rule_file = os.path.join(d, "rules.txt")
    with open(rule_file, "w") as f:
        f.write("".join(" ".join(str(v) for v in r) + "\n" for r in G.rules_of(fx)))
    bs = 32 if name != "syn" else 50
    set_seed(1)
    graph = KnowledgeGraph(d)                                   # text files, like the reference
    train_set, valid_set, test_set = TrainDataset(graph, bs), ValidDataset(graph, bs), TestDataset(graph, bs)
    model = PredictorPlus(graph, type="emb", num_layers=3, hidden_dim=16, entity_feature="bias", aggregator="sum")
    model.set_rules(rule_file)                                  # rule FILE path, ints only
    sd0 = {k[len("train_sd0_"):]: torch.from_numpy(v.copy()) for k, v in fx.items() if k.startswith("train_sd0_")}
    model.load_state_dict(sd0, strict=True)
    optim = torch.optim.Adam(model.parameters(), lr=0.005, weight_decay=0)
    solver = TrainerPredictor(model, train_set, valid_set, test_set, optim, gpus=[0])
    solver.train(batch_per_epoch=20, smoothing=0.2, print_every=1000)
    sd1 = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    for k, v in sd1.items():
        want = fx["train_sd1_" + k]
        moved = np.abs(want - fx["train_sd0_" + k]).max()
        err = np.abs(v - want).max()
        # parameters move by ~lr per step; after 20 steps the two runs agree to a small fraction of that
        assert err <= max(2e-2 * moved, 1e-6), (k, err, moved)
    mrr = solver.evaluate("valid", expectation=True)
    np.testing.assert_allclose(mrr, float(fx["train_mrr_valid"]), rtol=2e-3)
    # checkpoint round trip with the reference's state_dict keys (trainer.py:273-289)
    ck = str(tmp_path / (name + ".pt"))
    solver.save(ck)
    state = torch.load(ck, map_location="cpu")
    assert set(state["model"].keys()) == set(sd0.keys())
    solver.load(ck)


def test_reference_yaml_config_loads(tmp_path):
    from rnnlogic_b200.utils import load_config, save_config
    cfg_text = """gpus: [0]
save_path: out
load_path: null
seed: 1
num_iters: 1
data:
  data_path: ../data/x
  rule_file: ../data/x/rules.txt
  batch_size: 32
predictor:
  model:
    type: lstm
    num_layers: 3
    hidden_dim: 16
    entity_feature: bias
    aggregator: sum
    embedding_path: null
  optimizer:
    lr: 0.005
    weight_decay: 0
  train:
    smoothing: 0.2
    batch_per_epoch: 1000000
    print_every: 1000
  eval:
    expectation: True
"""
    p = tmp_path / "c.yaml"
    p.write_text(cfg_text)
    cfg = load_config(str(p))[0]
    assert cfg.predictor.model.type == "lstm" and cfg.data.batch_size == 32 and cfg.load_path is None
    assert dict(cfg.predictor.optimizer) == {"lr": 0.005, "weight_decay": 0}
    save_config(cfg, str(tmp_path))
    assert os.path.exists(tmp_path / "config.yaml")


@gpu
@pytest.mark.parametrize("per_step", [1, 3])
def test_pipelined_train_loop_equals_synchronous(per_step, tmp_path):
    """TrainerPredictor.train for Predictor(bias): the software-pipelined loop (next step grounded before the
    gradient exchange, losses read one step late) ends with the same parameters as the step-by-step loop."""
    from rnnlogic_b200.data import KnowledgeGraph, TrainDataset, ValidDataset, TestDataset
    from rnnlogic_b200.predictors import Predictor
    from rnnlogic_b200.trainer import TrainerPredictor
    from rnnlogic_b200.utils import set_seed
    fx = G.load("umls")
    d = str(tmp_path / "umls")
    write_dataset_dir(d, fx)
    out = []
    for pipelined in (True, False):
        set_seed(3)
        graph = KnowledgeGraph(d)
        sets = TrainDataset(graph, 32), ValidDataset(graph, 32), TestDataset(graph, 32)
        model = Predictor(graph, entity_feature="bias")
        model.set_rules(G.rules_of(fx))
        optim = torch.optim.Adam(model.parameters(), lr=0.01)
        solver = TrainerPredictor(model, *sets, optim, gpus=[0])
        solver.pipelined, solver.slots_per_step = pipelined, per_step
        solver.train(batch_per_epoch=14, smoothing=0.2, print_every=5)
        out.append({k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()})
    for k in out[0]:
        assert np.abs(out[0][k]).max() > 0
        np.testing.assert_allclose(out[0][k], out[1][k], rtol=1e-4, atol=1e-6, err_msg=k)


@gpu
@pytest.mark.parametrize("kind", ["predictor", "plus_emb", "plus_lstm"])
def test_graph_schedule_equals_eager(kind, tmp_path):
    """The reference's schedule (one batch per optimizer step) on the per-head CUDA graphs -- grounding of step k+1
    enqueued behind the scoring of step k -- ends with the same parameters as the eager step-by-step loop, and a new
    rule set (set_rules) drops the captured graphs."""
    from rnnlogic_b200.data import KnowledgeGraph, TrainDataset, ValidDataset, TestDataset
    from rnnlogic_b200.predictors import Predictor, PredictorPlus
    from rnnlogic_b200.trainer import TrainerPredictor
    from rnnlogic_b200.utils import set_seed
    fx = G.load("umls")
    d = str(tmp_path / "umls")
    write_dataset_dir(d, fx)
    out = []
    for graphs in (True, False):
        set_seed(5)
        graph = KnowledgeGraph(d)
        sets = TrainDataset(graph, 32), ValidDataset(graph, 32), TestDataset(graph, 32)
        if kind == "predictor":
            model = Predictor(graph, entity_feature="bias")
        else:
            model = PredictorPlus(graph, type=kind.split("_")[1], num_layers=2, hidden_dim=16, entity_feature="bias", aggregator="sum")
        model.set_rules(G.rules_of(fx))
        optim = torch.optim.Adam(model.parameters(), lr=0.01)
        solver = TrainerPredictor(model, *sets, optim, gpus=[0])
        solver.use_graphs, solver.pipelined = graphs, False
        solver.train(batch_per_epoch=16, smoothing=0.2, print_every=5)
        if graphs:
            assert len(model.__dict__.get("_graph_steps", {})) > 0 and not model.__dict__.get("_graphs_broken", False)
        out.append({k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()})
        if graphs:
            model.set_rules(G.rules_of(fx)[:-1])
            assert "_graph_steps" not in model.__dict__
    for k in out[0]:
        scale = max(1e-6, float(np.abs(out[1][k]).max()))
        assert np.abs(out[0][k] - out[1][k]).max() <= 2e-4 * scale + 1e-6, k
