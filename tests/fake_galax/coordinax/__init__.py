"""coordinax stand-in: the one vector type the potential API returns (register_funcs.py:134-155)."""
from . import vecs  # noqa: F401
