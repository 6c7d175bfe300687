"""slv_b200 -- B200-native SolEFP / SolCAMM frequency-shift engine (drop-in for the per-frame
evaluation path of globulion/slv: solvshift.solefp.EFP, solvshift.slvpar.Frag, solvshift.md.MDInput)."""
__version__ = '0.1'
