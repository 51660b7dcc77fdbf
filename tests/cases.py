"""Re-export of the synthetic case generators (aquagpusph_b200/cases.py)."""
from aquagpusph_b200.cases import *  # noqa: F401,F403
