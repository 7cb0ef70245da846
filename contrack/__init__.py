"""Drop-in import name of the reference package (contrack/__init__.py:26 there): ``from contrack import contrack``."""
from contrack_b200 import contrack          # noqa: F401
