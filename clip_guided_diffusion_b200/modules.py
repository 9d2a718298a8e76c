"""``from cgd.modules import MakeCutouts`` compatibility (cgd/modules.py:5-66)."""
from .guidance import MakeCutouts  # noqa: F401
