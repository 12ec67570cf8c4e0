"""Drop-in module: `from trainer import ...` in the reference's run scripts resolves to the B200 path."""
from rnnlogic_b200.trainer import *  # noqa: F401,F403


def __getattr__(name):          # TrainerGenerator (rule generator side) stays the reference's
    from _reference import reference_attr
    return reference_attr("trainer", name)
