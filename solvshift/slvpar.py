"""Drop-in alias: ``solvshift.slvpar`` of the reference maps onto ``slv_b200.slvpar``."""
from slv_b200.slvpar import *          # noqa: F401,F403
from slv_b200.slvpar import __all__    # noqa: F401
