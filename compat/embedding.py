"""Drop-in module: `from embedding import RotatE` resolves to the B200 path's module."""
from rnnlogic_b200.embedding import RotatE  # noqa: F401
