"""Unit constants of the SolEFP path.

The reference takes these from the un-vendored ``libbbg.units.UNITS`` (every class
on the path inherits it: solvshift/solefp.py:15, solvshift/slvpar.py:24).  The two
that enter the hot path were recovered from the reference's golden vector
(examples/1_small-clusters/f: printed Angstrom coordinates and cm^-1 shifts);
they are the CODATA-1986 values.
"""

BohrToAngstrom = 0.529177249
AngstromToBohr = 1.0 / BohrToAngstrom
HartreePerHbarToCmRec = 219474.63067
HartreeToKcalPerMole = 627.509451


class UNITS(object):
    """Mixin carrying the constants as attributes (libbbg.units.UNITS surface)."""
    BohrToAngstrom = BohrToAngstrom
    AngstromToBohr = AngstromToBohr
    HartreePerHbarToCmRec = HartreePerHbarToCmRec
    HartreeToKcalPerMole = HartreeToKcalPerMole
