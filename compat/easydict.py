"""Stand-in for the third-party ``easydict`` package (``from easydict import EasyDict`` in the reference's
run scripts, src/run_predictorplus.py:8, src/run_rnnlogic.py:8, and in src/utils.py:8): the attribute-access
dict the B200 path's own ``utils.load_config`` returns.  When the real package is installed it wins only if it
precedes compat/ on sys.path; both behave the same for the YAML configs."""
from rnnlogic_b200.utils import EasyDict  # noqa: F401

__all__ = ["EasyDict"]
