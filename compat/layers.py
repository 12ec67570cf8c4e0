"""Drop-in module: `from layers import ...` resolves to the B200 path's modules."""
from rnnlogic_b200.layers import *  # noqa: F401,F403
from rnnlogic_b200.layers import MLP, FuncToNode, FuncToNodeSum  # noqa: F401,E402
