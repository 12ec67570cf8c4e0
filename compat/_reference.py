"""Out-of-scope pieces (rule generator side: RuleDataset, Iterator, TrainerGenerator) are NOT
re-implemented here; when the reference's own src/ directory is also on sys.path (after compat/), they
are taken from there so that run_rnnlogic.py still imports everything it needs."""
import importlib.util
import os
import sys

_cache = {}


def reference_attr(module_name, attr):
    here = os.path.dirname(os.path.abspath(__file__))
    if module_name not in _cache:
        for p in sys.path:
            cand = os.path.join(p or ".", module_name + ".py")
            if os.path.isfile(cand) and os.path.abspath(os.path.dirname(cand)) != here:
                spec = importlib.util.spec_from_file_location("_reference_" + module_name, cand)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _cache[module_name] = mod
                break
        else:
            raise AttributeError("%s.%s is outside the B200 hot path; put the reference's src/ on PYTHONPATH "
                                 "after compat/ to use the reference's own implementation" % (module_name, attr))
    return getattr(_cache[module_name], attr)
