#!/bin/bash
# Drop-in proof: the reference's UNMODIFIED src/run_predictorplus.py runs against the B200 modules (PYTHONPATH=compat)
# and reaches the same valid / test MRR as the reference itself (CPU, same seed, same rule file, same YAML keys).
#
#   scripts/dropin_check.sh stage     (build container: needs /root/reference; writes dropin/_stage, git-ignored,
#                                      travels to the GPU box with gpurun)
#   scripts/dropin_check.sh ref       (build container, CPU: the reference run -> dropin/_stage/ref_<ds>.log)
#   scripts/dropin_check.sh gpu       (GPU box: the same script + configs against compat/ -> gpurun_out/dropin_<ds>.log)
#   scripts/dropin_check.sh compare   (anywhere: profiles/r2_dropin.txt)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
STAGE="$ROOT/dropin/_stage"
REF=/root/reference
DATASETS="umls kinship"

case "$1" in
stage)
    rm -rf "$STAGE"; mkdir -p "$STAGE"/{src,config,standins,miner}
    cp "$REF/src/run_predictorplus.py" "$STAGE/src/"                       # byte-identical copy of the reference script
    sha256sum "$REF/src/run_predictorplus.py" "$STAGE/src/run_predictorplus.py" > "$STAGE/script.sha256"
    g++ -O3 -w -o "$STAGE/miner/rnnlogic" "$REF/miner/rnnlogic.cpp" "$REF/miner/main.cpp" -lpthread
    for ds in $DATASETS; do
        mkdir -p "$STAGE/data/$ds"
        for f in entities.dict relations.dict train.txt valid.txt test.txt; do cp "$REF/data/$ds/$f" "$STAGE/data/$ds/"; done
        # README.md:49 miner flags; the predictor wants ints only, so the H column is dropped
        (cd "$STAGE/miner" && ./rnnlogic -data-path "$STAGE/data/$ds" -max-length 3 -threads 8 -lr 0.01 -wd 0.0005 -temp 100 \
            -iterations 1 -top-n 0 -top-k 0 -top-n-out 0 -output-file "$STAGE/data/$ds/mined_rules.txt" > "$STAGE/miner/$ds.log" 2>&1)
        awk '{NF=NF-1; print}' "$STAGE/data/$ds/mined_rules.txt" | sort -u | head -n 12000 > "$STAGE/data/$ds/rules.txt"
        for arm in ref gpu; do
            gp="null"; [ "$arm" = gpu ] && gp="[0]"
            cat > "$STAGE/config/${ds}_${arm}.yaml" <<YAML
gpus: $gp
save_path: ../out_${ds}_${arm}
load_path: null
seed: 1
num_iters: 1

data:
  data_path: ../data/$ds
  rule_file: ../data/$ds/rules.txt
  batch_size: 32

predictor:
  model:
    type: emb
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
    print_every: 20
  eval:
    expectation: True
YAML
        done
    done
    # stand-ins for the reference's two third-party imports that are not installed here (test infrastructure)
    cat > "$STAGE/standins/torch_scatter.py" <<'PY'
import torch
def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0 and reduce == "sum"
    return torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add_(0, index, src)
scatter_add = scatter
scatter_min = scatter_max = scatter_mean = None
PY
    cat > "$STAGE/standins/easydict.py" <<'PY'
class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v
    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(EasyDict(x) if isinstance(x, dict) else x for x in v)
        super().__setitem__(k, v)
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)
    __setattr__ = __setitem__
PY
    wc -l "$STAGE"/data/*/rules.txt; cat "$STAGE/script.sha256"
    ;;
ref)
    for ds in $DATASETS; do
        (cd "$STAGE/src" && PYTHONPATH="$STAGE/standins:$REF/src" python run_predictorplus.py --config "../config/${ds}_ref.yaml" \
            > "$STAGE/ref_${ds}.log" 2>&1) || { tail -5 "$STAGE/ref_${ds}.log"; exit 1; }
        grep -E "MRR|Hit1 " "$STAGE/ref_${ds}.log" | tail -4
    done
    ;;
gpu)
    mkdir -p "$ROOT/gpurun_out"
    for ds in $DATASETS; do
        (cd "$STAGE/src" && PYTHONPATH="$ROOT/compat:$ROOT" python run_predictorplus.py --config "../config/${ds}_gpu.yaml" \
            > "$ROOT/gpurun_out/dropin_${ds}.log" 2>&1) || { tail -15 "$ROOT/gpurun_out/dropin_${ds}.log"; exit 1; }
        grep -E "MRR|Hit1 " "$ROOT/gpurun_out/dropin_${ds}.log" | tail -4
    done
    ;;
compare)
    python - "$STAGE" "$ROOT" <<'PY'
import re, sys, os
stage, root = sys.argv[1], sys.argv[2]
out = ["drop-in check: the reference's unmodified src/run_predictorplus.py (sha256 below) with PYTHONPATH=compat on one B200",
       "against the reference itself on CPU (same script, same YAML keys, same seed, same mined rule file).", ""]
out.append(open(os.path.join(stage, "script.sha256")).read().strip())
def metrics(path):
    vals = {}
    for ln in open(path):
        m = re.search(r"(Hit1|Hit3|Hit10|MR|MRR)\s*:\s*([0-9.]+)", ln)
        if m:
            vals.setdefault(m.group(1), []).append(float(m.group(2)))
    return vals
for ds in ("umls", "kinship"):
    a, b = metrics(os.path.join(stage, "ref_%s.log" % ds)), metrics(os.path.join(root, "gpurun_out", "dropin_%s.log" % ds))
    out.append("")
    out.append("%s (PredictorPlus emb/sum/bias, H=16, 1 epoch of train, then filtered valid and test metrics):" % ds)
    for k in ("Hit1", "Hit3", "Hit10", "MR", "MRR"):
        for split, i in (("valid", 0), ("test", 1)):
            ra, rb = a[k][i], b[k][i]
            out.append("  %-5s %-5s reference-CPU %.6f   B200 %.6f   rel.diff %.2e" % (split, k, ra, rb, abs(ra - rb) / max(abs(ra), 1e-12)))
    losses = lambda p: [float(m.group(1)) for m in (re.search(r" \d+ \d+ ([0-9.]+) [0-9.]+$", ln.strip()) for ln in open(p)) if m]
    la, lb = losses(os.path.join(stage, "ref_%s.log" % ds)), losses(os.path.join(root, "gpurun_out", "dropin_%s.log" % ds))
    out.append("  train loss (mean over print_every=20 batches): reference %s | B200 %s" % (" ".join("%.5f" % x for x in la[:6]), " ".join("%.5f" % x for x in lb[:6])))
open(os.path.join(root, "profiles", "r2_dropin.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
PY
    ;;
*) echo "usage: $0 stage|ref|gpu|compare"; exit 2;;
esac
