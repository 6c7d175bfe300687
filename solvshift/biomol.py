"""Drop-in name of the reference module solvshift/biomol.py: the implementation is slv_b200.biomol."""
from slv_b200.biomol import *  # noqa: F401,F403
from slv_b200.biomol import BiomoleculeFragmentation, Topology, parse_tc  # noqa: F401
