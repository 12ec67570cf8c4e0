"""Drop-in module: `from trainer import ...` in the reference's run scripts resolves to the B200 path."""
from rnnlogic_b200.trainer import *  # noqa: F401,F403
