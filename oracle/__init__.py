"""CPU oracle for the SolEFP frequency-shift path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``slv_b200/`` (the product) may import this package.  It is used by
``tests/``, by ``__graft_entry__.smoke()`` as the checker, and by ``bench.py`` for the
``cpu_baseline`` / ``--impl reference`` legs.

Contents
  solefp_oracle.py  NumPy restatement of the reference's per-frame algorithm
                    (solvshift/solefp.py:1011-1629 and the natives it calls).
  f77c.py           F77-subset -> C transpiler: builds oracle/_ref/libslvref.so from the
                    reference's own Fortran sources where they lie under /root/reference
                    (no Fortran compiler exists in this image).  Used to validate the
                    restatement of solpol2.f / shftex.f / efprot.f.

Parity pinning:  ele_mea, ele_ea, cor_mea, cor_ea, dis_mea, dis_mea_iso and
Kabsch/rotation are pinned by the reference's golden data (examples/1_small-clusters/f, dat).
"""
