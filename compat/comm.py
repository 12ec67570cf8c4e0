"""Drop-in module: `from comm import ...` in the reference's run scripts resolves to the B200 path."""
from rnnlogic_b200.comm import *  # noqa: F401,F403
from rnnlogic_b200.comm import get_rank, get_world_size, get_group, init_process_group, synchronize, get_cpu_count  # noqa: F401,E402
