"""``from cgd.clip_util import MakeCutouts`` compatibility: the reference re-exports it there (cgd/clip_util.py:13).  CLIP weight loading
and the text tower live in ``cgd.py`` (`_load_clip_sd`, `_encode_text`): they run once per run, outside the hot path."""
from .guidance import MakeCutouts  # noqa: F401
