# A/B helper: bench.py under different library builds / launch-shape overrides (run on the GPU box)
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))"; }
run base
RL_CPW=32 run cpw32
RL_SCORE_WARPS=2 run score_warps2
