"""`from generators import Generator` (src/run_rnnlogic.py:16): the rule generator is outside the B200 hot path
and stays the reference's -- resolved from the reference's own src/ when it follows compat/ on sys.path."""


def __getattr__(name):
    from _reference import reference_attr
    return reference_attr("generators", name)
