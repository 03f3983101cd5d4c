"""``import evs as faiss`` -- the five-symbol faiss surface oldapp.py uses (oldapp.py:9, :87-88, :98, :117,
:2005, :2112), served by the B200 engine.  Everything lives in ``evo-ssearch_b200/``."""
from evo_ssearch_b200 import *  # noqa: F401,F403
from evo_ssearch_b200 import __all__  # noqa: F401
