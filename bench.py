#!/usr/bin/env python
"""bench.py -- SolEFP frequency-shift throughput on synthetic MD trajectories (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--frames F] [--waters NW]

Workload (BASELINE.json configs[1]): NMA-d7 in 500 waters, SolEFP Coulomb (MEA+EA+corrections) +
polarisation, 10 000 synthetic frames per GPU per step (SURVEY.md 8(d) generator, seed 2), mode 8,
ccut 40 / pcut 17 / ecut 12 Bohr.  One step = one pass of the hot path over the whole frame batch.
Frames shard across ranks with no data-path collective ("scaling": "weak": every rank gets the same
number of frames); one gather of the per-frame rows at the end of each step.

value      = frames/s, whole job, coordinates already resident in HBM (CUDA events on the launch stream,
             barrier + synchronize on both sides, max over ranks)
e2e        = frames/s through EFP.eval_frames with HOST (pinned) coordinates: H2D of the coordinates and
             D2H of the result rows inside the timed region
roofline   = dominant kernel (k_pol) against the FP64 DFMA peak measured by tools/fp64_peak.cu
cpu_baseline = oracle port (NumPy + BLAS, reference cost structure: LU inverse + 2 DGEMM per mode) on a
             bounded sample of the same frames, rank 0 only.
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

MODE = 8
CUTS = (40.0, 17.0, 12.0)

# BASELINE.json configs that fit one GPU; 2 is the configuration the metric is quoted on.
WORKLOADS = {
    2: dict(name='NMA-d7 + %(n)d waters, SolEFP elect(MEA+EA+corr)+pol', solute='nma-d7', solvents=['water'], nsolv=500, mode=8,
            flags=dict(elect=True, corr=True, pol=True), frames=10000, spacing=5.87, seed=2),
    3: dict(name='MeSCN in 50/50 water/DMSO (%(n)d fragments), full SolEFP elect+pol+rep+disp', solute='mescn',
            solvents=['water', 'dmso'], nsolv=400, mode=4, flags=dict(elect=True, corr=True, pol=True, rep=True, disp=True),
            frames=10000, spacing=9.5, seed=3, jit_lattice=0.3),
    5: dict(name='NMA + %(n)d waters, full SolEFP elect+pol+rep+disp', solute='nma', solvents=['water'], nsolv=4000, mode=8,
            flags=dict(elect=True, corr=True, pol=True, rep=True, disp=True), frames=2000, spacing=5.87, seed=5),
}
FLOPS_PER_T_PAIR = 49.0      # one (centre, centre) dipole-tensor apply to two vectors: R 3, r^2 5, rsqrt 1, powers 4,
                             # two dots 10, two scalings 2, six accumulations x 4 (DESIGN.md "k_pol")
FLOPS_PER_FIELD_PAIR = 123.0  # one (centre, site) multipole field through octupoles (slv::site_field_acc, counted)
POL_DRAM_BYTES_PER_FRAME = (86.834688e6 + 346.631424e6) / 1184.0   # ncu capture of k_pol, profiles/k_pol_r01h_ncu.txt
SWEEP_BODY_TFLOPS = 20.5     # algorithmic TFLOP/s of the k_pol sweep loop run alone (profiles/sweep_peak_r01.json, two rows/thread)
FLOPS_PER_ELECT_PAIR = 300.0  # one (solute site, solvent site) pair through R^-4 with two moment sets + corrections


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--config', type=int, default=2, choices=sorted(WORKLOADS), help='BASELINE.json config number')
    ap.add_argument('--frames', type=int, default=0, help='frames per GPU per step (0 = the config default)')
    ap.add_argument('--waters', type=int, default=0, help='solvent fragments (0 = the config default)')
    ap.add_argument('--unique', type=int, default=250, help='distinct generated frames (tiled up to --frames)')
    ap.add_argument('--cpu-sample', type=int, default=3, help='frames of the cpu_baseline sample')
    return ap.parse_args()


def make_frames(wl, nw, nunique, nframes, rank):
    """Synthetic trajectory of the workload: returns (bsm list, ind, nmol, coords[nframes, natoms, 3])."""
    import numpy as np
    from slv_b200 import synth
    from slv_b200.slvpar import Frag
    sol = Frag(wl['solute']); solv = [Frag(n) for n in wl['solvents']]
    types = np.random.default_rng(wl['seed']).integers(0, len(solv), nw) if len(solv) > 1 else np.zeros(nw, int)
    nu = min(nunique, nframes)
    X = synth.make_trajectory(sol.get_pos(), [f.get_pos() for f in solv], types, nu, seed=wl['seed'] + 1000 * rank,
                              spacing=wl['spacing'], jit_lattice=wl.get('jit_lattice', 0.6))
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


def cpu_reference_frames(wl, bsm, ind, nmol, X, nsample):
    """Oracle port timed on the host cores: returns (frames/s, seconds, nframes)."""
    from oracle import solefp_oracle as O
    pars = [b.get() for b in bsm]
    nt = len(bsm)
    fl = wl['flags']
    t0 = time.time()
    for f in range(nsample):
        O.eval_frame(X[f % len(X)], ind, nmol, pars, [None] * nt, [None] * nt, wl['mode'], *CUTS,
                     elect=True, ea=True, corr=True, pol=fl.get('pol', False), disp=fl.get('disp', False),
                     rep=fl.get('rep', False), literal_gemm=True)
    dt = time.time() - t0
    return nsample / dt, dt, nsample


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = WORKLOADS[a.config]; nw = a.waters or wl['nsolv']
    bsm, ind, nmol, X = make_frames(wl, nw, 8, 8, 0)
    per_step = 1
    for _ in range(a.warmup):
        cpu_reference_frames(wl, bsm, ind, nmol, X, per_step)
    t0 = time.time()
    for s in range(a.steps):
        cpu_reference_frames(wl, bsm, ind, nmol, X[s % len(X):], per_step)
    dt = time.time() - t0
    fps = a.steps * per_step / dt
    cores = os.cpu_count()
    line = dict(metric='md_frames_per_sec', value=fps, unit='frames/s', n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=dt / a.steps * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
                data='synthetic', impl='reference',
                config=dict(workload=(wl['name'] % dict(n=nw)) + ', mode %d, ccut/pcut/ecut 40/17/12 Bohr' % wl['mode'],
                            baseline_config=a.config, frames_per_step=per_step),
                cpu_baseline=dict(value=fps, unit='frames/s', cores=cores, kind='port',
                                  sample='%d frame(s) per step of the same synthetic workload; NumPy/BLAS restatement with the '
                                         'reference cost structure (LU inverse + 2 DGEMM per mode)' % per_step),
                e2e=dict(value=fps, unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    a = parse()
    if a.impl == 'reference':
        return run_reference(a)
    import numpy as np
    import torch
    import torch.distributed as dist
    from slv_b200 import _lib
    from slv_b200.solefp import EFP

    world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    wl = WORKLOADS[a.config]
    nw = a.waters or wl['nsolv']; nframes = a.frames or wl['frames']
    bsm, ind, nmol, X = make_frames(wl, nw, a.unique, nframes, rank)
    nt = len(bsm)
    efp = EFP(freq=True, mode=wl['mode'], ccut=CUTS[0], pcut=CUTS[1], ecut=CUTS[2], cunit=True, **wl['flags'])
    efp.max_chunk = min(nframes, 10240)
    d_coords = torch.from_numpy(X).to(dev)
    h_coords = torch.from_numpy(X).pin_memory()
    nf = d_coords.shape[0]
    rows = torch.empty((nf, _lib.NOUT), dtype=torch.float64, device=dev)
    status = torch.empty((nf,), dtype=torch.int32, device=dev)
    gathered = [torch.empty_like(rows) for _ in range(world)] if world > 1 and rank == 0 else None
    plan = None

    def step_device():
        nonlocal plan
        r, s = efp.eval_frames(d_coords, ind, nmol, bsm, [None] * nt, [None] * nt) if plan is None else plan.eval_device(d_coords, rows, status)
        if plan is None:
            plan = efp._get_plan()
            rows.copy_(r); status.copy_(s)
        if world > 1:
            dist.gather(rows, gathered, dst=0)      # the path's one collective: per-frame rows to rank 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step_device()
    barrier()
    launches_per_step = plan.last_launches()
    sampler = ClockSampler(local); sampler.start()
    # ---- device-resident timing
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    plan.set_timing(True)
    e0.record()
    for _ in range(a.steps):
        step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    ktimes = plan.get_timing()
    plan.set_timing(False)
    ms_pol = ktimes['k_pol'][0]; ms_nopol = ktimes['k_frame_elect'][0]
    # ---- end-to-end: host pinned coordinates -> rows on the host
    h_rows = torch.empty((nf, _lib.NOUT), dtype=torch.float64).pin_memory()
    h_status = torch.empty((nf,), dtype=torch.int32).pin_memory()
    for _ in range(2):
        efp.eval_frames(h_coords, out=h_rows, status=h_status)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        efp.eval_frames(h_coords, out=h_rows, status=h_status)     # the public call: pinned host in, host rows out
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True; sampler.join(timeout=2)

    t = torch.tensor([ms, ms_nopol, e2e_s, ms_pol], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_nopol, e2e_s, ms_pol = [float(v) for v in t.cpu()]
    r = rows.cpu().numpy()
    st = int(status.max().item())
    if rank == 0:
        total_frames = nf * world
        fps = total_frames * a.steps / (ms * 1e-3)
        npol = r[:, _lib.OUT['n_pol']]; nel = r[:, _lib.OUT['n_elect']]; its = r[:, _lib.OUT['pol_iters']]
        ncent = r[:, _lib.OUT['pol_ncent']]; nsite = r[:, _lib.OUT['pol_nsite']]; tpairs = r[:, _lib.OUT['pol_pairs']]
        pol_flops = float((its * tpairs * FLOPS_PER_T_PAIR + ncent * nsite * FLOPS_PER_FIELD_PAIR).sum())
        pol_ms = max(ms_pol, 1e-9) / a.steps
        ndma_c = int(bsm[0].get()['ndma'])
        ndma_s = float(np.mean([bsm[i].get()['ndma'] for i in ind[1:]]))
        elect_pairs = float((ndma_c * ndma_s * nel).sum())
        peak = 34.0
        pk = os.path.join(ROOT, 'profiles', 'fp64_peak_r01.json')
        if os.path.isfile(pk):
            peak = json.load(open(pk))['fp64_dfma_tflops']
        achieved = pol_flops / (pol_ms * 1e-3) / 1e12
        hbm = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
        cpu_fps, cpu_dt, cpu_n = cpu_reference_frames(wl, bsm, ind, nmol, X, a.cpu_sample)
        line = dict(
            metric='md_frames_per_sec', value=fps, unit='frames/s', n_gpus=world, steps=a.steps, warmup=max(a.warmup, 3),
            ms_per_step=ms / a.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
            config=dict(workload=(wl['name'] % dict(n=nw)) + ', mode %d, ccut/pcut/ecut 40/17/12 Bohr, %d frames per GPU per step' % (wl['mode'], nf),
                        baseline_config=a.config, frames_per_gpu=nf, l2='inputs larger than L2 (%.0f MB of coordinates per step)' % (X.nbytes / 1e6),
                        parallelism='frames sharded over %d rank(s), one gather of the result rows' % world,
                        mean_fragments_in_pcut=float(npol.mean()), mean_pol_iters=float(its.mean()), status_max=st,
                        kernel_ms_per_step=dict((k, v[0] / a.steps) for k, v in ktimes.items())),
            pair_interactions_per_sec=elect_pairs * world * a.steps / (ms * 1e-3),
            pol_tensor_pairs_per_sec=float((its * tpairs).sum()) * world * a.steps / (ms * 1e-3),
            gpu_launches=int(launches_per_step * a.steps + (1 if world > 1 else 0) * a.steps * 0),
            clocks=sampler.summary(),
            e2e=dict(value=total_frames * a.steps / e2e_s, unit='frames/s', h2d_bytes_per_step=int(X.nbytes),
                     d2h_bytes_per_step=int(nf * (_lib.NOUT * 8 + 4))),
            roofline=dict(bound='fp64', kernel='k_pol', achieved=achieved, peak=peak, unit='TFLOP/s', frac=achieved / peak,
                          traffic=(POL_DRAM_BYTES_PER_FRAME * nf if a.config == 2 else None), traffic_source='ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of k_pol on 1184 '
                          'frames of this workload (profiles/k_pol_r01h_ncu.txt), scaled per frame to this launch',
                          ceiling=dict(achieved_by='tools/sweep_peak.cu: the sweep body alone (no solver phases) at the kernel launch shape',
                                       alg_tflops=SWEEP_BODY_TFLOPS, frac_of_ceiling=achieved / SWEEP_BODY_TFLOPS),
                          peak_source='tools/fp64_peak.cu DFMA micro-benchmark on this pool (profiles/fp64_peak_r01.json); '
                          'MEASURED_PEAKS.json carries no FP64 figure',
                          flops_per_step=pol_flops, ms_per_step=pol_ms, launches_per_step=ktimes['k_pol'][1] // a.steps, share_of_step=pol_ms / (ms / a.steps)),
            roofline_stream=dict(bound='hbm', kernel='k_frame_elect', achieved=(X.nbytes + nf * 128) / (ms_nopol / a.steps * 1e-3) / 1e9,
                                 peak=hbm, unit='GB/s', note='coordinate streaming kernel is FP64-bound by the pair work, not HBM-bound'),
            cpu_baseline=dict(value=cpu_fps, unit='frames/s', cores=os.cpu_count(), kind='port',
                              sample='%d frames of the same workload in %.1f s (NumPy/BLAS oracle with the reference cost '
                                     'structure: LU inverse + 2 DGEMM per mode)' % (cpu_n, cpu_dt)))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
