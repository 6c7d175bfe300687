"""Drop-in alias: ``solvshift.solefp`` of the reference maps onto ``slv_b200.solefp``."""
from slv_b200.solefp import *          # noqa: F401,F403
from slv_b200.solefp import __all__    # noqa: F401
