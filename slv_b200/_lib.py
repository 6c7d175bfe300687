"""ctypes binding of slv_b200/csrc/libslvgpu.so (the C ABI of include/slvgpu.h).

The library is the only arithmetic path of the product: importing this module raises if the
shared object is missing, and every call raises on a non-zero return code (no CPU fallback).
"""
import ctypes
import os

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SLVGPU_LIB') or os.path.join(_HERE, 'csrc', 'libslvgpu.so')   # SLVGPU_LIB: developer override (kernel variants)

NOUT = 24
NMETA = 32
NSOL = 16
FLAG = dict(elect=1, ea=2, corr=4, pol=8, disp=16, rep=32, glob=64)
OUT = dict(ele_mea=0, ele_ea=1, cor_mea=2, cor_ea=3, pol_mea=4, rep_mea=5, dis_mea=6, dis_mea_iso=7,
           rms_c=8, rms_smax=9, rms_ave=10, n_elect=11, n_pol=12, n_rep=13, pol_iters=14, pol_resid=15, pol_add=16, pol_ncent=17, pol_nsite=18, pol_pairs=19,
           env_ele_mea=20, env_ele_ea=21, env_cor_mea=22, env_cor_ea=23)
M = dict(NATOMS=0, NDMA=1, NPOL=2, NMOS=3, NBASIS=4, NSUPL=5, IO_SUPL=6, IO_REORD=7, O_POS=8, O_RDMA=9, O_MOM=10,
         O_RPOL=11, O_POL=12, O_POLINV=13, O_DPOLI=14, O_LMOC=15, O_FOCK=16, O_VECL=17, O_ZNUC=18, IO_BFC=19,
         IO_BFL=20, IO_BFNP=21, O_BFEXP=22, O_BFCOEF=23, REMOVABLE=24, NSUBSH=25, IO_SUBSH=26, IO_SHPAIR=27, IO_SHITEM=28, O_PPTAB=29, IO_BFBLK=30, NBFBLK=31)
S = dict(SMEA=0, SEA=1, CMEA=2, CEA=3, DPOLI1=4, LMOC1=5, DPOL1=6, LVEC=7, DMA1=8, FOCK1=9, VECL1=10)
ST_NONFINITE, ST_POL_NOCONV, ST_SINGULAR, ST_OVERFLOW = 1, 2, 4, 8


class PlanDesc(ctypes.Structure):
    _fields_ = [('ntypes', ctypes.c_int32), ('ninst', ctypes.c_int32), ('natoms_total', ctypes.c_int32),
                ('flags', ctypes.c_int32), ('ccut', ctypes.c_double), ('pcut', ctypes.c_double),
                ('ecut', ctypes.c_double), ('type_meta', ctypes.c_void_p), ('sol_meta', ctypes.c_void_p),
                ('itab', ctypes.c_void_p), ('n_itab', ctypes.c_int64), ('dtab', ctypes.c_void_p),
                ('n_dtab', ctypes.c_int64), ('inst_type', ctypes.c_void_p), ('inst_atom0', ctypes.c_void_p),
                ('max_chunk', ctypes.c_int32), ('pol_maxit', ctypes.c_int32), ('pol_tol', ctypes.c_double),
                ('pol_cent_cap', ctypes.c_int32), ('want_detail', ctypes.c_int32)]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError("slv_b200: CUDA library %s is missing -- build it with "
                          "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    lib.slvgpu_plan_create.argtypes = [ctypes.POINTER(PlanDesc), ctypes.POINTER(vp)]
    lib.slvgpu_plan_destroy.argtypes = [vp]; lib.slvgpu_plan_destroy.restype = None
    lib.slvgpu_eval_frames.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.slvgpu_eval_frames_host.argtypes = [vp, vp, i64, vp, vp]
    lib.slvgpu_eval_frames_host_f32.argtypes = [vp, vp, i64, i32, vp, dbl, vp, vp]
    lib.slvgpu_last_launches.argtypes = [vp]; lib.slvgpu_last_launches.restype = i64
    lib.slvgpu_set_timing.argtypes = [vp, i32]
    lib.slvgpu_get_timing.argtypes = [vp, vp, vp]
    lib.slvgpu_get_pol_detail.argtypes = [vp, i64, vp, vp, vp, vp, i32]
    lib.slvgpu_kabsch.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.slvgpu_rotate_dma.argtypes = [vp, vp, vp, i64, vp]
    lib.slvgpu_rotate_pol.argtypes = [vp, i64, vp]
    lib.slvgpu_rotate_vec.argtypes = [vp, i64, vp, vp]
    lib.slvgpu_rotate_vecl.argtypes = [vp, i64, i32, vp, vp]
    lib.slvgpu_tcf.argtypes = [vp, vp, i64, i64, i32, vp]
    lib.slvgpu_expres.argtypes = [vp, i64, dbl, i64, vp, vp]
    lib.slvgpu_trig_transform.argtypes = [vp, i64, dbl, dbl, vp, i64, i32, vp]
    lib.slvgpu_hist.argtypes = [vp, i64, dbl, dbl, i32, vp]
    lib.slvgpu_xtc_scan.argtypes = [ctypes.c_char_p, vp, vp]
    lib.slvgpu_xtc_read.argtypes = [ctypes.c_char_p, i64, i64, vp, vp, vp, vp]
    lib.slvgpu_last_error.restype = ctypes.c_char_p
    lib.slvgpu_version.restype = ctypes.c_char_p
    return lib


lib = _load()


class SlvGpuError(RuntimeError):
    pass


def check(rc):
    if rc < 0:
        raise SlvGpuError('slvgpu error %d: %s' % (rc, lib.slvgpu_last_error().decode()))
    return rc


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def kabsch(ref_pos, xyz, suplist=None):
    """rot, transl, rms of Frag.sup's superimposition (slvpar.py:641-679), computed on the GPU."""
    ref = numpy.ascontiguousarray(ref_pos, numpy.float64)
    tgt = numpy.ascontiguousarray(xyz, numpy.float64)
    if len(tgt) != 1 and suplist is not None:
        ref = numpy.ascontiguousarray(ref[suplist]); tgt = numpy.ascontiguousarray(tgt[suplist])
    rot = numpy.zeros((3, 3)); tr = numpy.zeros(3); rms = ctypes.c_double(0.0)
    check(lib.slvgpu_kabsch(_ptr(ref), _ptr(tgt), len(ref), _ptr(rot), _ptr(tr), ctypes.byref(rms)))
    return rot, tr, rms.value


def frag_transform(frag, rot, transl):
    """Rotate/translate every tensor of a Frag in place (body of Frag.sup, slvpar.py:680-733)."""
    p = frag._p
    R = numpy.ascontiguousarray(rot, numpy.float64)

    def vec(key, shift):
        if key in p:
            a = numpy.ascontiguousarray(p[key], numpy.float64).copy()
            check(lib.slvgpu_rotate_vec(_ptr(a), a.size // 3, _ptr(R), _ptr(transl) if shift else None))
            p[key] = a
    for k in ('pos', 'lmoc', 'rpol', 'rdma'):
        vec(k, True)
    for k in ('lmoc1', 'lvec'):
        vec(k, False)
    for sfx in ('', '1', '2'):
        if 'dmad' + sfx in p:
            d = numpy.ascontiguousarray(p['dmad' + sfx]).copy(); q = numpy.ascontiguousarray(p['dmaq' + sfx]).copy()
            o = numpy.ascontiguousarray(p['dmao' + sfx]).copy()
            check(lib.slvgpu_rotate_dma(_ptr(d), _ptr(q), _ptr(o), d.size // 3, _ptr(R)))
            p['dmad' + sfx], p['dmaq' + sfx], p['dmao' + sfx] = d, q, o
    for k in ('dpol', 'dpol1', 'dpoli', 'dpoli1'):
        if k in p:
            a = numpy.ascontiguousarray(p[k]).copy()
            check(lib.slvgpu_rotate_pol(_ptr(a), a.size // 9, _ptr(R)))
            p[k] = a
    if 'vecl' in p:
        typ = numpy.ascontiguousarray(frag.get_bfs().get_bfst().sum(axis=1), numpy.int32)
        for k in ('vecl', 'vecc', 'vecl1', 'vecc1'):
            if k in p:
                a = numpy.ascontiguousarray(p[k]).copy()
                check(lib.slvgpu_rotate_vecl(_ptr(a), a.size // a.shape[-1], a.shape[-1], _ptr(typ), _ptr(R)))
                p[k] = a
