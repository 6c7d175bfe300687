#!/usr/bin/env python
"""bench.py -- SolEFP frequency-shift throughput on synthetic MD trajectories (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C] [--only]

Headline workload = BASELINE.json configs[2] ("config 3"): MeSCN in a 50/50 water/DMSO mix of 400 fragments, FULL
SolEFP (electrostatics MEA+EA+corrections, polarisation, exchange repulsion, dispersion: all four kernel groups),
10 000 synthetic frames per GPU per step (SURVEY.md 8(d) generator), mode 4, ccut 40 / pcut 17 / ecut 12 Bohr --
the configuration BASELINE.json quotes at 1/2/4/8 GPUs.  One step = one pass of the hot path over the whole frame
batch.  Frames shard across ranks with no data-path collective ("scaling": "weak": every rank gets the same number
of frames); one gather of the per-frame rows at the end of each step.  Unless --only is given, the same JSON line
nests complete sub-records for config 2 (NMA-d7 + 500 waters, elect+pol) and config 5 (NMA + 4000 waters, all terms)
under "configs"; --config 4 runs the protein-like shape (about 2000 mixed EFP fragments, remove_clashes).

value      = frames/s, whole job, coordinates already resident in HBM (CUDA events on the launch stream,
             barrier + synchronize on both sides, max over ranks)
e2e        = frames/s through EFP.eval_frames with HOST (pinned) coordinates: H2D of the coordinates and
             D2H of the result rows inside the timed region
roofline   = dominant kernel (k_pol) against the FP64 DFMA peak measured by tools/fp64_peak.cu; roofline_elect and
             roofline_rep are the same figure for the electrostatics and exchange-repulsion kernels
cpu_baseline / --impl reference = the oracle port (NumPy + BLAS restatement with the reference's cost structure: LU
             inverse + 2 DGEMM per mode, per-fragment integrals) farmed frame by frame over os.cpu_count() worker
             processes with one BLAS thread each -- the pp.Server(ncpus) model of util/slv_md-run:96-97, 155-169.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CUTS = (40.0, 17.0, 12.0)
HEADLINE = 3

# BASELINE.json configs that fit one GPU; 3 is the configuration the metric is quoted on at 1/2/4/8 GPUs.
WORKLOADS = {
    2: dict(name='NMA-d7 + %(n)d waters, SolEFP elect(MEA+EA+corr)+pol', solute='nma-d7', solvents=['water'], nsolv=500, mode=8,
            flags=dict(elect=True, corr=True, pol=True), frames=10000, spacing=5.87, seed=2),
    3: dict(name='MeSCN in 50/50 water/DMSO (%(n)d fragments), full SolEFP elect+pol+rep+disp', solute='mescn',
            solvents=['water', 'dmso'], nsolv=400, mode=4, flags=dict(elect=True, corr=True, pol=True, rep=True, disp=True),
            frames=10000, spacing=9.5, seed=3, jit_lattice=0.3),
    4: dict(name='protein-like through the biomol front-end: SCN probe residue in a 160-residue chain + waters (%(n)d EFP fragments '
                 'incl. charged side chains, 8 near-zone CONH point charges), full SolEFP, remove_clashes',
            biomol=True, nres=160, nsolv=2000, mode=4, flags=dict(elect=True, corr=True, pol=True, rep=True, disp=True), frames=1000,
            seed=4, cuts=(45.0, 16.0, 13.0), remove_clashes=True, unique=50),
    5: dict(name='NMA + %(n)d waters, full SolEFP elect+pol+rep+disp', solute='nma', solvents=['water'], nsolv=4000, mode=8,
            flags=dict(elect=True, corr=True, pol=True, rep=True, disp=True), frames=2000, spacing=5.87, seed=5),
}
# Algorithmic FP64 flop counts (frozen): FMA = 2, the reciprocal square root = 1.
FLOPS_PER_T_PAIR = 49.0       # k_pol sweep: one (centre, centre) dipole-tensor apply to two vectors: R 3, r^2 5, rsqrt 1,
                              # powers 4, two dots 10, two scalings 2, six accumulations x 4
FLOPS_PER_FIELD_PAIR = 123.0  # k_pol fields: one (centre, site) multipole field through octupoles (slv::site_field_acc)
FLOPS_PER_ELECT_PAIR = 248.0  # k_frame_elect: one (solute site, solvent site) pair through R^-4: slv::pair_potentials 154
                              # (R 3, r^2 5, rsqrt + powers 5, RR 6, RRR 10, mu.R 5, Theta.R 15, R.Theta.R 5, Omega:RRR 21,
                              # potential 7, field 23, field gradient 34, second gradient 11, charge-only gradient 4) + the
                              # contraction with the two SolCAMM moment sets and the two correction vectors 94
SWEEP_BODY_TFLOPS = 20.5      # algorithmic TFLOP/s of the k_pol sweep loop run alone (profiles/sweep_peak_r01.json)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--config', type=int, default=HEADLINE, choices=sorted(WORKLOADS), help='BASELINE.json config number')
    ap.add_argument('--only', action='store_true', help='no nested config 2 / config 5 sub-records')
    ap.add_argument('--frames', type=int, default=0, help='frames per GPU per step (0 = the config default)')
    ap.add_argument('--fragments', '--waters', dest='waters', type=int, default=0, help='solvent fragments (0 = the config default)')
    ap.add_argument('--unique', type=int, default=0, help='distinct generated frames, tiled up to --frames (0 = config default, 250)')
    ap.add_argument('--cpu-frames', type=int, default=0, help='frames of the cpu_baseline sample (0 = one per worker process)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    return ap.parse_args()


def workload_string(cfg, nw, nframes):
    wl = WORKLOADS[cfg]; cuts = wl.get('cuts', CUTS)
    return (wl['name'] % dict(n=nw)) + ', mode %d, ccut/pcut/ecut %g/%g/%g Bohr, %d frames per GPU per step' % (
        wl['mode'], cuts[0], cuts[1], cuts[2], nframes)


BIOMOL_SEQ = ['ASP', 'SER', 'GLY', 'LYS', 'ASN', 'ALA', 'GLU', 'TYR', 'PHE', 'THR', 'LEU', 'VAL', 'ILE', 'GLN', 'ARG', 'MET', 'HIE', 'CYS']


def make_biomol(wl, nfrag, nunique, nframes, rank):
    """Config 4: a synthetic protein-like system pushed through the real front-end (slv_b200/biomol.py: .tc tables, amide
    zones, $Frag input, near-zone charges).  Returns the workload dictionary of make_workload."""
    import numpy as np
    from slv_b200 import biomol, synth_biomol
    from slv_b200.units import AngstromToBohr
    tcdir = os.path.join(ROOT, 'slv_b200', 'data', 'tc')
    tc = biomol.parse_tc(os.path.join(tcdir, 'gmx.tc')); ptc = biomol.parse_tc(os.path.join(tcdir, 'probe.tc'))
    rng = np.random.default_rng(wl['seed'])
    seq = [BIOMOL_SEQ[k] for k in rng.integers(0, len(BIOMOL_SEQ), wl['nres'])]
    nprot = wl['nres'] + sum(len(tc[r]) for r in seq)                 # links + side-chain fragments (about)
    top, xyz = synth_biomol.make_system(tc, ptc, seq, rng, nwater=max(0, nfrag - nprot))
    cuts = wl['cuts']
    bf = biomol.BiomoleculeFragmentation(tc, ptc, wl['mode'], ccut=cuts[0], pcut=cuts[1], ecut=cuts[2], report=os.devnull,
                                         include_ions=False, include_polar=False)
    (ind, nmol, bsm, supl, reord), idx = bf.build(top, xyz)
    nu = min(nunique, nframes)
    rec = synth_biomol.jiggle(xyz, top, nu, np.random.default_rng(wl['seed'] + 1000 * rank), amp=0.02).astype(np.float32)
    rec = np.ascontiguousarray(np.concatenate([rec] * ((nframes + nu - 1) // nu))[:nframes])
    return dict(bsm=bsm, ind=ind, nmol=nmol, supl=supl, reord=reord, rec32=rec, idx=np.asarray(idx, np.int32), scale=AngstromToBohr,
                nfrag=len(ind) - 1, efp=bf._efp)


def make_workload(cfg, nw, nunique, nframes, rank):
    """Synthetic trajectory of a workload as float32 RECORDS in Angstrom (what DCD / XTC files hold) plus the frame slice:
    dict(bsm, ind, nmol, supl, reord, rec32 [nframes, natoms_src, 3], idx or None, scale)."""
    import numpy as np
    from slv_b200.units import AngstromToBohr
    wl = WORKLOADS[cfg]
    if wl.get('biomol'):
        return make_biomol(wl, nw, nunique, nframes, rank)
    bsm, ind, nmol, X = make_frames(wl, nw, nunique, nframes, rank)
    nt = len(bsm)
    rec = np.ascontiguousarray((X / AngstromToBohr).astype(np.float32))
    return dict(bsm=bsm, ind=ind, nmol=nmol, supl=[None] * nt, reord=[None] * nt, rec32=rec, idx=None, scale=AngstromToBohr,
                nfrag=len(ind) - 1, efp=None)


def bohr_frames(w):
    """The float64 Bohr coordinates the device-resident leg sees: exactly what the float32 entry computes on the device."""
    import numpy as np
    r = w['rec32'] if w['idx'] is None else w['rec32'][:, w['idx'], :]
    return np.ascontiguousarray(r.astype(np.float64) * w['scale'])


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and therefore the pinned staging buffers it allocates next) to the CPUs of the NUMA node the GPU
    hangs off: eight ranks pushing coordinates through one socket's memory controllers was the 8-GPU e2e limit."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), 'pci_bus_id') else None
        if bus is None:
            out = subprocess.run(['nvidia-smi', '-i', str(local_rank), '--query-gpu=pci.bus_id', '--format=csv,noheader'],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            bus = out
        bus = bus.lower()
        if bus.startswith('00000000:'):
            bus = bus[4:]
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus |= set(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def make_frames(wl, nw, nunique, nframes, rank):
    """Synthetic trajectory of the workload: returns (bsm list, ind, nmol, coords[nframes, natoms, 3])."""
    import numpy as np
    from slv_b200 import synth
    from slv_b200.slvpar import Frag
    sol = Frag(wl['solute']); solv = [Frag(n) for n in wl['solvents']]
    types = np.random.default_rng(wl['seed']).integers(0, len(solv), nw) if len(solv) > 1 else np.zeros(nw, int)
    nu = min(nunique, nframes)
    X = synth.make_trajectory(sol.get_pos(), [f.get_pos() for f in solv], types, nu, seed=wl['seed'] + 1000 * rank,
                              spacing=wl['spacing'], jit_lattice=wl.get('jit_lattice', 0.6), rmin=wl.get('rmin', 3.0))
    reps = (nframes + nu - 1) // nu
    X = np.concatenate([X] * reps)[:nframes]
    ind = [0] + [1 + int(t) for t in types]
    nmol = [sol.get_natoms()] + [solv[t].get_natoms() for t in types]
    return [sol] + solv, ind, nmol, np.ascontiguousarray(X)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.samples = []; self.reasons = set(); self.stop_flag = False; self.max_mhz = None

    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self.stop_flag:
            try:
                o = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                self.samples.append(float(o[0])); self.max_mhz = float(o[1])
                for n, v in zip(names, o[2:]):
                    if v.strip().lower().startswith('active'):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
_W = {}


def _cpu_worker_init(cfg, nw):
    """Worker process of the CPU farm: one BLAS/OpenMP thread, its own copy of the parameters (the reference pickles the
    whole EFP object into every pp job, util/slv_md-run:155-169)."""
    for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMEXPR_NUM_THREADS'):
        os.environ[k] = '1'
    try:
        from threadpoolctl import threadpool_limits
        _W['limit'] = threadpool_limits(1)
    except Exception:
        pass
    import warnings
    warnings.simplefilter('ignore')
    wl = WORKLOADS[cfg]
    w = make_workload(cfg, nw, 1, 1, 0)
    pars = [b.get() for b in w['bsm']]
    nenv = sum(1 for t in w['ind'] if pars[t].get('env'))
    _W.update(wl=wl, pars=pars, ind=w['ind'], nmol=w['nmol'], supl=w['supl'], reord=w['reord'], nenv=nenv)
    if wl.get('remove_clashes'):
        from slv_b200.solefp import EFP
        names = set(EFP._RCL_BASE + ['Acetamide'] + EFP._RCL_POLAR + EFP._RCL_IONS + ['Methyl Sulphate (VI) -1 anion'])
        _W['removable'] = set(t for t, p in enumerate(pars) if t != 0 and p.get('name') in names)


def _cpu_worker_frame(xyz):
    """One frame (gathered float64 Bohr coordinates) through the oracle: EFP.set + eval (+ eval_dma for the environment)."""
    import numpy as np
    from oracle import solefp_oracle as O
    wl = _W['wl']; fl = wl['flags']; ne = _W['nenv']
    ind, nmol = _W['ind'], _W['nmol']
    t0 = time.time()
    n_efp = sum(nmol[:len(nmol) - ne])
    O.eval_frame(xyz[:n_efp], ind[:len(ind) - ne], nmol[:len(nmol) - ne], _W['pars'], _W['supl'], _W['reord'], wl['mode'],
                 *wl.get('cuts', CUTS), elect=True, ea=True, corr=True, pol=fl.get('pol', False), disp=fl.get('disp', False),
                 rep=fl.get('rep', False), literal_gemm=True, removable=_W.get('removable'))
    if ne:
        q = np.array([_W['pars'][t]['dmac'][0] for t in ind[len(ind) - ne:]])
        ro = _W['reord'][ind[0]]
        pos_c = O.reorder_xyz(xyz[:nmol[0]], ro if ro is not None else np.arange(nmol[0]))
        O.eval_dma_frame(pos_c, _W['pars'][ind[0]], _W['supl'][ind[0]], wl['mode'], np.concatenate([xyz[n_efp:], q[:, None]], axis=1))
    return time.time() - t0


class CpuFarm(object):
    """Frames farmed over worker processes (spawned, so that each imports NumPy with a single BLAS thread whatever the
    parent's environment says -- torchrun exports OMP_NUM_THREADS=1 to the ranks, and a fork would inherit the parent's
    already-sized thread pools)."""

    def __init__(self, cfg, nw, nproc=None):
        import multiprocessing as mp
        try:
            avail = len(os.sched_getaffinity(0))
        except AttributeError:
            avail = os.cpu_count() or 1
        self.nproc = nproc or max(1, avail)
        env_keep = dict((k, os.environ.get(k)) for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'))
        for k in env_keep:
            os.environ[k] = '1'                      # inherited by the spawned workers
        try:
            self.pool = mp.get_context('spawn').Pool(self.nproc, initializer=_cpu_worker_init, initargs=(cfg, nw))
            self.pool.map(_noop, range(self.nproc))   # workers up and initialised before anything is timed
        finally:
            for k, v in env_keep.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    def run(self, frames):
        """Evaluate the frames; returns (wall seconds, per-frame worker seconds)."""
        t0 = time.time()
        per = self.pool.map(_cpu_worker_frame, list(frames), chunksize=1)
        return time.time() - t0, per

    def close(self):
        self.pool.close(); self.pool.join()


def _noop(_):
    return os.getpid()


def cpu_sample_record(cfg, nw, X, nframes_sample):
    """cpu_baseline object for the main JSON line: a bounded sample of the workload's frames on all host cores."""
    farm = CpuFarm(cfg, nw)
    n = nframes_sample or farm.nproc
    wall, per = farm.run([X[f % len(X)] for f in range(n)])
    farm.close()
    return dict(value=n / wall, unit='frames/s', cores=farm.nproc, kind='port',
                sample='%d frames of the same workload farmed over %d single-threaded worker processes in %.1f s (%.1f s per frame '
                       'per core); NumPy/BLAS oracle with the reference cost structure (LU inverse + 2 DGEMM per mode, '
                       'per-fragment AO integrals)' % (n, farm.nproc, wall, sum(per) / len(per)))


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = a.config; wl = WORKLOADS[cfg]; nw = a.waters or wl['nsolv']; nframes = a.frames or wl['frames']
    farm = CpuFarm(cfg, nw)
    per_step = a.cpu_frames or farm.nproc
    nu = min(4 * per_step, 64)
    w = make_workload(cfg, nw, nu, nu, 0)
    X = bohr_frames(w); nfrag = w['nfrag']
    k = 0
    for _ in range(a.warmup):
        farm.run([X[(k + f) % nu] for f in range(per_step)]); k += per_step
    t0 = time.time()
    per_all = []
    for _ in range(a.steps):
        _, per = farm.run([X[(k + f) % nu] for f in range(per_step)]); k += per_step
        per_all += per
    dt = time.time() - t0
    farm.close()
    fps = a.steps * per_step / dt
    sample = ('%d frames per step (one per worker process) of the same synthetic workload, farmed over %d single-threaded '
              'worker processes like pp.Server(ncpus) in util/slv_md-run; %.1f s per frame per core; NumPy/BLAS restatement with '
              'the reference cost structure (LU inverse + 2 DGEMM per mode, per-fragment AO integrals)'
              % (per_step, farm.nproc, sum(per_all) / max(1, len(per_all))))
    line = dict(metric='md_frames_per_sec', value=fps, unit='frames/s', n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=dt / a.steps * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
                data='synthetic', impl='reference',
                config=dict(workload=workload_string(cfg, nfrag, nframes), baseline_config=cfg),
                cpu_baseline=dict(value=fps, unit='frames/s', cores=farm.nproc, kind='port', sample=sample),
                e2e=dict(value=fps, unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def fp64_peak():
    peak, src = 34.0, 'fallback 34.0 TFLOP/s (148 SM x 64 DFMA x 1.965 GHz x 0.91)'
    for name in ('fp64_peak_r02.json', 'fp64_peak_r01.json'):
        pk = os.path.join(ROOT, 'profiles', name)
        if os.path.isfile(pk):
            peak = json.load(open(pk))['fp64_dfma_tflops']
            src = 'tools/fp64_peak.cu DFMA micro-benchmark on this pool (profiles/%s); MEASURED_PEAKS.json carries no FP64 figure' % name
            break
    return peak, src


def ncu_traffic(kernel, cfg):
    """DRAM bytes per frame of one kernel from the committed ncu --set full capture (profiles/ncu_traffic.json, written by
    tools/ncu_traffic.py from the .ncu-rep of the same bench command), or None."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if not os.path.isfile(p):
        return None, None
    t = json.load(open(p)).get('config%d' % cfg, {}).get(kernel)
    if not t:
        return None, None
    return t['dram_bytes_per_frame'], t['source']


def run_config(cfg, a, steps, warmup, world, rank, dev, with_cpu, sampler=None):
    """One complete measurement of a workload: device-resident value, e2e through the public call, per-kernel times,
    rooflines.  Returns the record (rank 0) or None."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from slv_b200 import _lib
    from slv_b200.solefp import EFP
    wl = WORKLOADS[cfg]
    cuts = wl.get('cuts', CUTS)
    main = cfg == a.config
    nw = (a.waters if main else 0) or wl['nsolv']; nframes = (a.frames if main else 0) or wl['frames']
    nunique = (a.unique if main else 0) or wl.get('unique', 250)
    w = make_workload(cfg, nw, nunique, nframes, rank)
    bsm, ind, nmol = w['bsm'], w['ind'], w['nmol']
    X = bohr_frames(w)                                  # float64 Bohr, gathered: the device-resident leg
    nw = w['nfrag']
    efp = w['efp'] or EFP(freq=True, mode=wl['mode'], ccut=cuts[0], pcut=cuts[1], ecut=cuts[2], cunit=True, **wl['flags'])
    efp.max_chunk = min(nframes, 10240)
    efp.set(X[0], ind, nmol, bsm, w['supl'], w['reord'])
    if wl.get('remove_clashes'):
        efp._removable = efp._removable_types('remove_by_name', False, False)
    d_coords = torch.from_numpy(X).to(dev)
    # the e2e leg starts from float32 trajectory records in pinned host memory (what DCD / XTC files hold); the frame
    # slice and the conversion to Bohr happen on the device inside the public call
    h_coords = torch.from_numpy(w['rec32']).pin_memory()
    e2e_kw = dict(scale=w['scale'], idx=w['idx'])
    nf = d_coords.shape[0]
    rows = torch.empty((nf, _lib.NOUT), dtype=torch.float64, device=dev)
    status = torch.empty((nf,), dtype=torch.int32, device=dev)
    gathered = [torch.empty_like(rows) for _ in range(world)] if world > 1 and rank == 0 else None
    plan = efp._get_plan()

    def step_device():
        plan.eval_device(d_coords, rows, status)
        if world > 1:
            dist.gather(rows, gathered, dst=0)      # the path's one collective: per-frame rows to rank 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        step_device()
    barrier()
    launches_per_step = plan.last_launches()
    # ---- device-resident timing
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    plan.set_timing(True)
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    ktimes = plan.get_timing()
    plan.set_timing(False)
    # ---- end-to-end: host pinned coordinates -> rows on the host
    h_rows = torch.empty((nf, _lib.NOUT), dtype=torch.float64).pin_memory()
    h_status = torch.empty((nf,), dtype=torch.int32).pin_memory()
    for _ in range(2):
        efp.eval_frames(h_coords, out=h_rows, status=h_status, **e2e_kw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        efp.eval_frames(h_coords, out=h_rows, status=h_status, **e2e_kw)     # the public call: pinned host records in, host rows out
    barrier()
    e2e_s = time.perf_counter() - t0

    kn = ('k_frame_elect', 'k_disp', 'k_pol', 'k_rep')
    t = torch.tensor([ms, e2e_s] + [ktimes[k][0] for k in kn], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tv = [float(v) for v in t.cpu()]
    ms, e2e_s = tv[0], tv[1]
    kms = dict((k, tv[2 + i] / steps) for i, k in enumerate(kn))
    r = rows.cpu().numpy()
    st = int(status.max().item())
    same = bool((h_rows.numpy() == r).all())          # the host entry computes the same rows bit for bit
    del d_coords, h_coords, rows, status
    efp._plan = None; del plan
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    total_frames = nf * world
    fps = total_frames * steps / (ms * 1e-3)
    O = _lib.OUT
    npol = r[:, O['n_pol']]; nel = r[:, O['n_elect']]; nrep = r[:, O['n_rep']]; its = r[:, O['pol_iters']]
    ncent = r[:, O['pol_ncent']]; nsite = r[:, O['pol_nsite']]; tpairs = r[:, O['pol_pairs']]
    peak, peak_src = fp64_peak()
    step_ms = ms / steps
    rec = dict(
        value=fps, unit='frames/s', ms_per_step=step_ms,
        e2e=dict(value=total_frames * steps / e2e_s, unit='frames/s', h2d_bytes_per_step=int(w['rec32'].nbytes),
                 input='float32 trajectory records in Angstrom (pinned host memory); frame slice and conversion to Bohr on the device',
                 d2h_bytes_per_step=int(nf * (_lib.NOUT * 8 + 4)), rows_identical_to_device_path=same),
        config=dict(workload=workload_string(cfg, nw, nf), baseline_config=cfg),
        detail=dict(frames_per_gpu=nf, l2='inputs larger than L2 (%.0f MB of coordinates per step)' % (X.nbytes / 1e6),
                    distinct_frames='%d generated geometries tiled to %d frames: iteration counts and ragged list sizes sample '
                                    '%d configurations, no work is skipped' % (min(nunique, nf), nf, min(nunique, nf)),
                    parallelism='frames sharded over %d rank(s), one gather of the result rows' % world,
                    mean_fragments_in_ccut=float(nel.mean()), mean_fragments_in_pcut=float(npol.mean()),
                    mean_fragments_in_ecut=float(nrep.mean()), mean_pol_iters=float(its.mean()), status_max=st,
                    kernel_ms_per_step=kms, sum_kernel_ms=sum(kms.values()),
                    overlap='ms_per_step / sum of kernel ms = %.3f' % (step_ms / max(1e-9, sum(kms.values())))),
        gpu_launches=int(launches_per_step * steps))
    # ---- rooflines: algorithmic flops of a kernel / its own event-timed duration
    pol_flops = float((its * tpairs * FLOPS_PER_T_PAIR + ncent * nsite * FLOPS_PER_FIELD_PAIR).sum())
    if kms['k_pol'] > 0:
        ach = pol_flops / (kms['k_pol'] * 1e-3) / 1e12
        tr, trs = ncu_traffic('k_pol', cfg)
        rec['roofline'] = dict(bound='fp64', kernel='k_pol', achieved=ach, peak=peak, unit='TFLOP/s', frac=ach / peak,
                               traffic=(tr * nf if tr else None), traffic_source=trs, peak_source=peak_src,
                               ceiling=dict(achieved_by='tools/sweep_peak.cu: the sweep body alone (resident tile, no solver phases), two 256-thread CTAs per SM',
                                            alg_tflops=SWEEP_BODY_TFLOPS, frac_of_ceiling=ach / SWEEP_BODY_TFLOPS),
                               flops_per_step=pol_flops, ms_per_step=kms['k_pol'], share_of_step=kms['k_pol'] / step_ms)
    ndma_c = int(bsm[0].get()['ndma'])
    ndma_s = float(np.mean([bsm[i].get()['ndma'] for i in ind[1:] if not bsm[i].get().get('env')]))
    elect_pairs = float((ndma_c * ndma_s * nel).sum())
    ea = elect_pairs * FLOPS_PER_ELECT_PAIR / (kms['k_frame_elect'] * 1e-3) / 1e12
    rec['roofline_elect'] = dict(bound='fp64', kernel='k_frame_elect', achieved=ea, peak=peak, unit='TFLOP/s', frac=ea / peak,
                                 flops_per_pair=FLOPS_PER_ELECT_PAIR, pairs_per_step=elect_pairs, ms_per_step=kms['k_frame_elect'],
                                 hbm_gbs=(X.nbytes + nf * _lib.NOUT * 8) / (kms['k_frame_elect'] * 1e-3) / 1e9,
                                 note='superimposition (Kabsch) and moment rotation of every instance are inside this kernel and '
                                      'not in the flop count; HBM traffic is the coordinates once: far below the HBM roofline')
    rec['pair_interactions_per_sec'] = elect_pairs * world * steps / (ms * 1e-3)
    rec['pol_tensor_pairs_per_sec'] = float((its * tpairs).sum()) * world * steps / (ms * 1e-3)
    if kms['k_rep'] > 0:
        # half transform X = V C_B^T (4 integral kinds) and LMO integrals C_A X (6 products) per (solute, fragment) pair
        pa = bsm[ind[0]].get(); nbA, nmA = int(pa['nbasis']), int(pa['nmos'])
        import collections
        cnt = collections.Counter(t for t in ind[1:] if not bsm[t].get().get('env'))
        flop = lambda b: 8.0 * nbA * b['nbasis'] * b['nmos'] + 12.0 * nmA * b['nmos'] * nbA
        mean_flop = sum(flop(bsm[t].get()) * c for t, c in cnt.items()) / float(sum(cnt.values()))
        rep_pairs = float(nrep.sum())
        ra = rep_pairs * mean_flop / (kms['k_rep'] * 1e-3) / 1e12
        rec['roofline_rep'] = dict(bound='fp64', kernel='k_rep_* (integrals + transform + CALCFI)', achieved=ra, peak=peak, unit='TFLOP/s',
                                   frac=ra / peak, fragment_pairs_per_step=rep_pairs, fragment_pairs_per_sec=rep_pairs * world * steps / (ms * 1e-3),
                                   transform_flops_per_pair=mean_flop, ms_per_step=kms['k_rep'],
                                   note='flops count only the AO->LMO transform (8 nbsA nbsB nmosB + 12 nmosA nmosB nbsA per pair, type-averaged); '
                                        'the Obara-Saika integrals and CALCFI are not counted')
    if with_cpu:
        rec['cpu_baseline'] = cpu_sample_record(cfg, nw, X, a.cpu_frames)
    return rec


def main():
    a = parse()
    if a.impl == 'reference':
        return run_reference(a)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    sampler = ClockSampler(local); sampler.start()
    rec = run_config(a.config, a, a.steps, a.warmup, world, rank, dev, with_cpu=False)
    sub = {}
    if not a.only and a.config == HEADLINE:
        for c in (2, 5):
            s = run_config(c, a, max(2, a.steps // 2), 3, world, rank, dev, with_cpu=False)
            if s is not None:
                sub['config%d' % c] = s
    sampler.stop_flag = True; sampler.join(timeout=2)
    if rank == 0:
        wl = WORKLOADS[a.config]
        line = dict(metric='md_frames_per_sec', value=rec.pop('value'), unit=rec.pop('unit'), n_gpus=world, steps=a.steps,
                    warmup=max(a.warmup, 3), ms_per_step=rec.pop('ms_per_step'), higher_is_better=True, scaling='weak',
                    vs_baseline=None, dtype='f64', data='synthetic')
        line.update(rec)
        line['clocks'] = sampler.summary()
        line['host'] = dict(numa_node_of_gpu=numa, cpus=os.cpu_count())
        if sub:
            line['configs'] = sub
        if not a.no_cpu and world == 1:          # rank 0 at N=1 only
            nw = a.waters or wl['nsolv']
            if numa is not None:
                os.sched_setaffinity(0, range(os.cpu_count()))     # the CPU farm uses every core again
            X = bohr_frames(make_workload(a.config, nw, 64, 64, 0))
            line['cpu_baseline'] = cpu_sample_record(a.config, nw, X, a.cpu_frames)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
