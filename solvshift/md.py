"""Drop-in alias: ``solvshift.md`` of the reference maps onto ``slv_b200.md``."""
from slv_b200.md import *          # noqa: F401,F403
from slv_b200.md import __all__    # noqa: F401
