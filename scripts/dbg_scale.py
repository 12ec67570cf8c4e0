import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from tests.test_gpu_scale import workload, batches_of, DEV
from rnnlogic_b200.engine import Grounder
from rnnlogic_b200 import _lib
kg, okg, rules, cr, train = workload("fb15k237")
sparse = Grounder(kg, cr, DEV)
dense = Grounder(kg, cr, DEV, force_dense=True)
for bi, b in enumerate(batches_of(train, kg.relation_size, 4, 1)):
    q = int(b[0, 1]); ids = cr.head_rules[q]
    print('batch', bi, 'q', q, 'B', len(b), 'rules', len(ids), 'nodes', cr.head_nodes[q], 'rows', cr.head_rows[q], 'chunks', cr.head_chunks[q], 'lvl', cr.level_chunks[q], cr.level_nodes[q], flush=True)
    if not ids: continue
    etr = torch.from_numpy(kg.edge_index_of(b)).to(DEV); h = torch.from_numpy(b[:, 0]).to(DEV)
    for name, gr in (('sparse', sparse), ('dense', dense)):
        sl = gr.make_slots([q], [len(b)], h, None, etr)
        torch.cuda.synchronize(); print('  slots ok', name, flush=True)
        L = _lib.lib()
        gr._run(sl, 32)
        torch.cuda.synchronize(); print('  run ok', name, int(sl.overflow.item()), flush=True)
        c = gr.rule_counts(sl, ids[:3]); torch.cuda.synchronize(); print('  counts ok', name, int(c.sum()), flush=True)
