"""Drop-in module: `from predictors import ...` in the reference's run scripts resolves to the B200 path."""
from rnnlogic_b200.predictors import *  # noqa: F401,F403
