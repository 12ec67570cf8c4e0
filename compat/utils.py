"""Drop-in module: `from utils import ...` in the reference's run scripts resolves to the B200 path."""
from rnnlogic_b200.utils import *  # noqa: F401,F403
from rnnlogic_b200.utils import load_config, save_config, set_seed, set_logger  # noqa: F401,E402
import logging, os, sys, json, random, argparse  # noqa: F401,E401,E402  (the reference's trainer does `from utils import *`)
