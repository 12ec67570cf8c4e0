"""Drop-in module: `from data import ...` in the reference's run scripts resolves to the B200 path."""
from rnnlogic_b200.data import *  # noqa: F401,F403


def __getattr__(name):          # RuleDataset / Iterator (rule generator side) stay the reference's
    from _reference import reference_attr
    return reference_attr("data", name)
