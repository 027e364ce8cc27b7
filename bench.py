#!/usr/bin/env python
"""bench.py -- trajectory samples cost-evaluated+updated per second (Panda Stoch-GPMP, H=64).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (CPU baseline arm: the oracle port)

A step = one Stoch-GPMP iteration over P=512 particles x S=64 samples x H=64 waypoints of the
7-DoF Panda (BASELINE.json configs[3], "panda_spheres"): sample from the GP prior, forward
kinematics to 50 collision spheres, collision + GP cost (+ importance-sampling term), softmax
update.  Independent particles shard across GPUs with no data-path collective (weak scaling:
512 particles per GPU).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'trajectory samples cost-evaluated+updated/sec (Panda, H=64)'
UNIT = 'samples/s'
P_PER_GPU, S, H, DOF = 512, 64, 64, 7
D, M, NS, NO = 2 * DOF, 64 * 14, 50, 16
# algorithmic work per sample (SURVEY.md 8d / BASELINE.md section 4, C4 column)
FLOP_SAMPLING = M * (M + 1)                                   # 803,712  triangular mat-vec
FLOP_FK = H * (600 + 18 * NS)                                 # 96,000
FLOP_SDF = (H - 1) * NS * (10 * NO + 3)                       # 513,450
FLOP_GP = (H - 1) * (10 * DOF + 5) + 4 * D                    # 4,781
FLOP_IS = 2 * M                                               # 1,792
FLOP_UPDATE = 2 * M + 10                                      # 1,802
BYTES_UPDATE = M * 4                                          # K3 reads every sample row once


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], bf16=p['bf16_tflops'], bf16_sustained=p.get('bf16_tflops_sustained'),
                    sm_max_mhz=p.get('sm_max_mhz', 1965.0), source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, sm_max_mhz=1965.0,
                source='fallback (B200_PROFILING.md)')


def ncu_traffic(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json), or None."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    try:
        table = json.load(open(path))
    except Exception:
        return None
    for key, rec in sorted(table.items(), key=lambda kv: -len(kv[0])):        # longest key first: cost_eval_chain2 before cost_eval
        if key in kernel_name:
            return dict(bytes=rec['dram_read_bytes'] + rec['dram_write_bytes'], unit='B', source=rec.get('source'),
                        algorithmic_bytes=rec.get('algorithmic_bytes'))
    return None


def ncu_issue_slots(kernel_key):
    """Executed warp instructions / issue slots per sample of the named kernel from the committed ncu capture
    (profiles/ncu_traffic.json), or None."""
    try:
        rec = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))[kernel_key]
        return dict(issue_slots_per_sample=rec['issue_slots_per_sample'], warp_instructions_per_sample=rec['warp_instructions_per_sample'],
                    source=rec.get('source'))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self._stop_evt = index, period, [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
                 getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonHwPowerBrakeSlowdown', 0x80): 'hw_power_brake'}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), n_samples=len(self.samples))


def physical_gpu_index(local_rank):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        try:
            return int(vis.split(',')[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_setup(P_cpu):
    """The oracle port of the reference's Stoch-GPMP iteration in its 'faithful' cost model:
    per-particle scale_tril [P,M,M] re-factorised every iteration + broadcast batched mat-vec sampler
    (mp_priors_multi.py:100-123), eager-torch cost chain, dense IS term."""
    from motion_planning_baselines_b200 import configs
    from oracle import gp_prior
    from oracle.build import TA, oracle_field, oracle_robot
    from oracle.costs import CostSpec
    cfg = configs.config('C4')
    sig = cfg['params']
    start, goal = torch.tensor(cfg['start']), torch.tensor(cfg['goal'])
    spec = CostSpec(oracle_robot(cfg['robot'], cfg['dt']), H, cfg['dt'], start, goal,
                    [oracle_field(cfg['obstacles'], cfg['robot'])], sigma_start=sig['sigma_start'], sigma_gp=sig['sigma_gp'],
                    sigma_coll=sig['sigma_coll'], sigma_goal_prior=sig['sigma_goal_prior'], tensor_args=TA)
    K_s = gp_prior.unary_K(D, sig['sigma_start_sample'], TA)
    K_g = gp_prior.unary_K(D, sig['sigma_goal_sample'], TA)
    Q = gp_prior.gp_Q_inv(DOF, cfg['dt'], sig['sigma_gp_sample'], TA)
    Sinv = gp_prior.prior_precision(H, DOF, cfg['dt'], K_s, Q, K_g, TA)
    L = gp_prior.precision_to_scale_tril(Sinv)
    mean = gp_prior.const_vel_mean(torch.cat((start, torch.zeros(DOF))), torch.cat((goal, torch.zeros(DOF))), cfg['dt'], H, DOF, TA)
    means = mean.unsqueeze(0).repeat(P_cpu, 1, 1)
    return spec, means, L, Sinv, sig


def cpu_reference_run(P_cpu, steps, warmup, faithful=True):
    from oracle import fields as ofields
    from oracle import planners as oplanners
    ofields.EXACT_SQRT = False          # time torch's own fp32 sqrt, as the reference's eager CPU path would run it
    torch.set_num_threads(os.cpu_count() or 1)
    spec, means, L, Sinv, sig = cpu_reference_setup(P_cpu)
    gen = torch.Generator().manual_seed(0)
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            eps = torch.randn(S, P_cpu, M, generator=gen)
            t0 = time.perf_counter()
            out = oplanners.stoch_gpmp_iteration(spec, means, L, Sinv, eps, sig['temperature'], sig['step_size'], faithful=faithful)
            means = out['means']
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    t = float(np.sum(times))
    return dict(value=P_cpu * S * steps / t, ms_per_step=1e3 * t / steps)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    P_cpu = args.cpu_particles
    steps, warmup = min(args.steps, 5), min(args.warmup, 1)
    r = cpu_reference_run(P_cpu, steps, warmup, faithful=True)
    cores = torch.get_num_threads()
    sample = (f'{P_cpu} particles x {S} samples x {H} waypoints Panda per step (of {P_PER_GPU}); {steps} timed steps, '
              f'{warmup} warm-up; oracle port in the reference cost model (per-particle scale_tril re-factorised every '
              f'iteration, batched mat-vec sampler, eager torch fp32)')
    cfg = workload_config(args.gpus)
    # the CPU arm times a BOUNDED sample of the workload: fewer particles per step (samples/s is flat in P on the CPU,
    # BASELINE.md section 3.4; the full P = 512 needs a 1.64 GB factor re-factorised every iteration, ~2 min per step)
    cfg.update(particles_timed=P_cpu, particles_per_gpu_workload=P_PER_GPU, noise='torch.randn on the host (outside the timed region)',
               sharding='CPU only (rank 0)', global_samples_per_step=P_cpu * S)
    line = dict(impl='reference', metric=METRIC, value=r['value'], unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=r['ms_per_step'], higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', config=cfg,
                cpu_baseline=dict(value=r['value'], unit=UNIT, cores=cores, kind='port', sample=sample),
                e2e=dict(value=r['value'], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def workload_config(n):
    return dict(workload='panda_spheres Stoch-GPMP (BASELINE.json configs[3]): 7-DoF Panda FK -> 50 collision spheres vs 16 '
                         'sphere obstacles, 512 particles x 64 samples x 64 waypoints per GPU',
                particles_per_gpu=P_PER_GPU, samples=S, waypoints=H, dof=DOF, robot_spheres=NS, obstacles=NO,
                global_samples_per_step=n * P_PER_GPU * S, sharding=f'particles x{n} (no data-path collective)',
                noise='drawn inside the sampling kernel (Philox4x32-10 keyed on the global element index), as the reference '
                      'draws inside MultiMPPrior.sample; the injected-noise variant (8 rotating 117 MB buffers resident in HBM) '
                      'is reported under "injected_noise"',
                sample_rows='kept dof-major ([dof][2H]) between the three kernels of the iteration, an internal layout with '
                            'bit-identical values (tests/test_gpu_dof_major.py); planner.state_samples converts to the '
                            "reference's [H][2 dof] on demand, outside the iteration (MPB_X_DM=0: reference layout throughout)",
                l2='per-step outputs (117 MB of samples; 235 MB with injected noise) exceed / fill the 126 MB L2 and every '
                   'step overwrites them')


def measure_fp32_peak(dev):
    """Measured FFMA rate of this device (csrc/microbench.cu): best of 5 launches, CUDA events."""
    import ctypes as C

    from motion_planning_baselines_b200 import _lib
    scratch = torch.empty(148 * 8 * 256 * 2, **dev)
    iters, nthreads = 4096, C.c_longlong(0)
    lib = _lib.lib()
    best = None
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.mpb_bench_fp32_peak(_lib.ptr(scratch), iters, C.byref(nthreads), _lib.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return 2.0 * 64 * iters * nthreads.value / (best * 1e-3) / 1e12


def parity_check(planner, cfg, sig, means0, eps, n_particles=8):
    """Replays the LAST benchmarked step of a particle subset through the CPU oracle (checker only, outside every timed
    region): costs 1e-5, flags / argmin identical -- evidence that the timed kernels computed the right thing."""
    from oracle import check
    torch.set_num_threads(1)
    P = planner.num_particles
    rng = np.random.default_rng(0)
    sub = sorted(rng.choice(P, n_particles, replace=False).tolist())
    ref = check.stoch_gpmp_subset(cfg, H, sig, means0, planner._sample_dist.scale_tril, planner.Sigma_inv, eps, sub)
    idx = torch.as_tensor(sub, device=means0.device)
    amp = float((ref['samples'] - means0[idx].cpu().unsqueeze(1)).abs().max())
    e_samples = float((planner.state_samples[idx].cpu() - ref['samples']).abs().max()) / max(amp, 1e-30)
    e_cost = check.rel_err(planner.costs[idx], ref['costs'])
    flags = bool(torch.equal(planner.free_flags.view(P, S)[idx].bool().cpu(), ref['free']))
    argmin = bool(torch.equal(planner.costs[idx].argmin(1).cpu(), ref['costs'].argmin(1)))
    torch.set_num_threads(os.cpu_count() or 1)
    return dict(particles_checked=sub, samples_max_err_over_noise_amplitude=e_samples, costs_max_rel_err=e_cost,
                collision_free_flags_identical=flags, argmin_identical=argmin,
                ok=bool(e_samples < 1e-5 and e_cost < 1e-5 and flags and argmin),
                against='oracle.check.stoch_gpmp_subset (CPU restatement of stoch_gpmp.py:235-279) on the noise the kernels drew '
                        '(mpb_philox_normal dump of the last timed step)')


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-particles', type=int, default=32, help='particles per step of the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true')
    ap.add_argument('--no-parity-check', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist

    from motion_planning_baselines_b200 import _lib, configs
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import StochGPMP
    from motion_planning_baselines_b200.robots import Robot

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback for the product path)'
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner to STDOUT on the first collective; stdout must carry exactly one JSON line,
        # so file descriptor 1 points at stderr until the communicator is up.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = dict(device=torch.device('cuda', local_rank), dtype=torch.float32)
    K, W = args.steps, max(args.warmup, 3)

    cfg = configs.config('C4')
    sig = cfg['params']
    torch.manual_seed(1000)
    robot = Robot(cfg['robot'], dt=cfg['dt'], tensor_args=dev)
    field = CollisionField(cfg['obstacles'], tensor_args=dev)
    P = P_PER_GPU
    # one global noise stream for the whole job: rank r owns particles [r*P, (r+1)*P) of world*P
    planner = StochGPMP(robot=robot, n_dof=DOF, n_support_points=H, num_particles_per_goal=P, opt_iters=1, dt=cfg['dt'],
                        start_state=torch.tensor(cfg['start']).to(**dev),
                        multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0),
                        collision_fields=[field], tensor_args=dev, num_samples=S, seed=1000,
                        noise_particle_offset=rank * P, noise_particles_global=world * P, **sig)
    assert planner._sample_dist.kron_tc_kind == 1, 'the C4 factor must take the default structured sampler'
    means0 = planner._particle_means.clone()
    planner_uses_gen = getattr(planner._sample_dist, 'scale_tril_kron_gen', None) is not None     # which K1 draws the noise

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev['device'], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def timed(step_fn, n, every=1):
        """n steps bracketed by barrier + synchronize; per-stage events inside every `every`-th step (an event record between
        two kernels keeps the second from starting under programmatic dependent launch, so the other steps run exactly as
        planner.optimize() issues them) -> (ms total max over ranks, stage ms averaged over the instrumented steps)."""
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] if i % every == 0 else None for i in range(n)]
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(n):
            step_fn(i, ev[i])
        t1.record()
        barrier()
        stage = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(4)] for e in ev if e is not None]).mean(0)
        return allmax(t0.elapsed_time(t1)), stage

    # ---- headline: K steps, noise drawn inside K1 (the way the reference's optimize() is called) -----------------------
    def headline_step(i, ev):
        # every 5th step: the three kernels as separate C calls with CUDA events between them; the others: the same three
        # launches from ONE C call (mpb_stoch_gpmp_iter_kron_gen_dm), as planner.optimize() issues them
        if ev is not None:
            planner.step_staged(None, events=ev)
        else:
            planner.step_fused()

    for i in range(W):
        planner.step_staged(None) if i % 2 else planner.step_fused()
    planner._particle_means.copy_(means0)
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    ms_total, stage_ms = timed(headline_step, K, every=5)
    # A step costs three launches issued from Python; when the host stalls (a cold or contended host core -- seen once in
    # twelve runs on fresh boxes: 0.279 instead of 0.236 ms) the GPU idles between launches and the K steps measure the host.
    # Such a region is recognisable from inside: its steps take longer than the kernels' own times add up to (the
    # uninstrumented steps OVERLAP their kernels under programmatic dependent launch, so they are normally shorter).  It is
    # then measured once more -- the same K steps -- and both figures are reported.
    first_attempt = None
    if allmax(ms_total / K / float(np.sum(stage_ms))) > 1.06:      # the same decision on every rank
        first_attempt = ms_total / K
        ms_total, stage_ms = timed(headline_step, K, every=5)
    clocks = sampler.stop()
    # launches of OUR kernels per step: K1 (tcgen05 sampler; draws the noise and, on one extra warp, computes Sigma^-1 mu), K2, K3;
    # four when the sampler in use has no mat-vec warp
    mv_in_k1 = planner_uses_gen and planner._sinv_structured
    launches_per_step = 3 if mv_in_k1 else 4
    value = world * P * S * K / (ms_total * 1e-3)
    dof_major = bool(getattr(planner, '_x_dm_fresh', False))      # the timed steps kept their sample rows dof-major (default for C4)
    free_frac = float(planner.free_flags.float().mean())
    # K3 reads only the sample rows whose weight is non-zero: count them (roofline on bytes actually needed)
    rows_nonzero = int((planner._w_buf != 0).sum())

    # ---- parity of what was just timed (rank 0; the checker runs after / outside the timed region) ----------------------
    check = None
    if rank == 0 and not args.no_parity_check:
        planner._particle_means.copy_(means0)
        desc = planner._noise.desc()
        planner.step_staged(None)
        torch.cuda.synchronize()
        eps_last = planner._sample_dist.replay_noise(desc, S)
        check = parity_check(planner, cfg, sig, means0, eps_last)
        del eps_last

    # ---- the same step on INJECTED noise resident in HBM (the parity configuration; K1 reads 117 MB more) --------------
    n_buf = 8
    eps_bufs = [torch.randn(S, P, M, **dev) for _ in range(n_buf)]
    planner._particle_means.copy_(means0)
    for i in range(3):
        planner.step_staged(eps_bufs[i % n_buf])
    planner._particle_means.copy_(means0)
    Ki = max(3, min(K, 20))
    ms_inj, stage_inj = timed(lambda i, ev: planner.step_staged(eps_bufs[i % n_buf], events=ev), Ki)
    del eps_bufs

    # ---- e2e through the public API with HOST buffers: planner.optimize(opt_iters=1) -----------------------------------
    # per step: the problem (initial particle trajectories) comes from pinned host memory, the optimised trajectories go
    # back to pinned host memory; the noise is drawn by the planner on the device, as in the reference
    # consecutive steps are independent problems: the next problem's upload and the previous result's download run on
    # copy streams and overlap this step's kernels (double-buffered device means / pinned result buffers)
    h_means = means0.cpu().pin_memory()
    h_trajs = [torch.empty(P, H, D).pin_memory() for _ in range(2)]
    h_traj = h_trajs[0]
    d_means = [torch.empty_like(means0) for _ in range(2)]
    up_stream, down_stream = torch.cuda.Stream(), torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    Ke = max(3, min(K, 20))

    def e2e_steps(n):
        # ONE copy stream and two cross-stream edges per step (each edge costs ~10 us of front-end latency on this box,
        # profiles/tools/e2e_probe.py): the main stream waits for the upload of its problem; the copy stream waits for the
        # result.  Copy-stream order: upload of problem i + 1 (its buffer was last read by step i - 1, whose download -- queued
        # earlier on this stream -- waited for that step), then the download of result i.
        torch.cuda.synchronize()
        with torch.cuda.stream(up_stream):
            d_means[0].copy_(h_means, non_blocking=True)
            ready[0].record(up_stream)
        for i in range(n):
            b = i % 2
            main_stream.wait_event(ready[b])
            planner._particle_means = d_means[b]
            traj = planner.optimize(opt_iters=1)
            done = torch.cuda.Event()
            done.record(main_stream)
            with torch.cuda.stream(up_stream):
                if i + 1 < n:
                    d_means[1 - b].copy_(h_means, non_blocking=True)
                    ready[1 - b].record(up_stream)
                up_stream.wait_event(done)
                h_trajs[b].copy_(traj, non_blocking=True)
            traj.record_stream(up_stream)
        torch.cuda.synchronize()
    e2e_steps(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps(Ke)
    e1.record()
    barrier()
    ms_e2e = allmax(e0.elapsed_time(e1))
    e2e_value = world * P * S * Ke / (ms_e2e * 1e-3)
    planner._particle_means = d_means[0]

    # e2e with INJECTED host noise (parity-style call: 117 MB of eps uploaded per step; PCIe bound) -- secondary
    h_eps = [torch.randn(S, P, M).pin_memory() for _ in range(2)]
    d_eps = [torch.empty(S, P, M, **dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream()

    def upload(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            d_eps[b].copy_(h_eps[b], non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_injected(n):
        for b in range(2):
            consumed[b].record(main_stream)
        upload(0)
        for i in range(n):
            b = i % 2
            if i + 1 < n:
                upload(i + 1)                       # overlaps the next step's upload with this step's kernels
            main_stream.wait_event(ready[b])
            traj = planner.optimize(opt_iters=1, eps=[d_eps[b]])
            consumed[b].record(main_stream)
            h_traj.copy_(traj, non_blocking=True)
        torch.cuda.synchronize()
    Kj = max(3, min(K, 10))
    e2e_injected(2)
    barrier()
    e0.record()
    e2e_injected(Kj)
    e1.record()
    barrier()
    ms_e2e_inj = allmax(e0.elapsed_time(e1))
    del h_eps, d_eps

    # N > 1: the sample-split variant (BASELINE.json configs[4]): ONE problem's samples sharded over the ranks, one
    # exchange step per iteration (all-gather of packed softmax records over NCCL/NVLink).  Strong scaling: total samples fixed.
    split_block = None
    if world > 1:
        del planner
        torch.cuda.empty_cache()
        try:
            import bench_configs
            res = bench_configs.bench_c5(dev, Ns=(100000, 1000000), split=bench_configs.TimedSplit(), iters=10,
                                         stomp_Ns=(100000,), reduce_max=allmax)
            split_block = dict(scaling='strong (total samples fixed; efficiency = t(1 GPU) / (N x t(N GPUs)), the 1-GPU times '
                                       'are the C5 rows of other_configs in the N=1 line)',
                               exchange='per iteration: all_gather_into_tensor of the packed (m, Z, cmin, argmin, sum e(x-mu)) '
                                        'records + fixed-order combine on every rank; MPPI adds a scalar all-gather (batch-summed '
                                        'obstacle cost, quirk B2) and a [T,sd] gather of the best rollout',
                               results=res)
        except Exception as exc:
            split_block = dict(error=repr(exc))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    n_samp = P * S
    fp32_measured = measure_fp32_peak(dev)
    sm_mhz_peak = clocks.get('sm_max_mhz') or pk['sm_max_mhz']
    fp32_nominal = 148 * 128 * 2 * sm_mhz_peak * 1e6 / 1e12

    def kernel_table(stage, k1_reads_eps):
        flop_k1 = DOF * (2 * H) * (2 * H + 1)       # 7 independent [128,128] triangular mat-vecs = 115,584 flop / sample
        bytes_k1 = M * 4 * (2 if k1_reads_eps else 1)
        gen = (not k1_reads_eps) and planner_uses_gen
        dm = gen and dof_major
        k1_name = ('sample_gp_kron_gen_dm3_kernel<7> (K1: tcgen05 kind::f16, work unit = 64 samples x ONE dof with the factor as the M = 128 '
                   'operand, 8-deep accumulator ring in TMEM, factors of the CTA\'s dofs bulk-loaded once, dof-major sample rows stored '
                   'straight from registers, 24 Philox producer warps in three groups whose first warp issues the group\'s MMAs, '
                   'Sigma^-1 mu shared by the eight epilogue warps)' if dm else
                   'sample_gp_kron_gen_kernel<7,32> (K1: tcgen05 kind::f16 with the factor as the M = 128 operand, 32-sample tiles, two '
                   'sets of 7 accumulators in TMEM (the epilogue of a tile overlaps the MMAs of the next), warp-specialised Philox '
                   'producers, bulk-async factor loads and row stores, Sigma^-1 mu on an extra warp)' if gen else
                   'sample_gp_kron_mma_kernel<7,64,%s> (K1: structured GP sampler, warp MMA fp16x2 split%s)'
                   % ('false' if k1_reads_eps else 'true', '' if k1_reads_eps else ', Philox noise in-kernel'))
        k1 = dict(kernel=k1_name,
                  ms=float(stage[0]), bound='hbm', achieved=bytes_k1 * n_samp / (stage[0] * 1e-3) / 1e9, peak=pk['hbm'], unit='GB/s',
                  algorithmic_bytes_per_sample=bytes_k1, algorithmic_flop_per_sample=flop_k1,
                  note='HBM floor = x written' + (' + eps read' if k1_reads_eps else '') + '; the factor decouples over the 7 dofs '
                       '(exact zeros, verified bit-exactly at setup): 115,584 flop / sample instead of the dense 803,712'
                       + ('; measured limiter (profiles/r02_k1_dm.txt, DESIGN 4f): the noise generation on the CUDA cores -- 29.4 M '
                          'normals, Philox4x32-10 + Box-Muller + fp16 split = ~35 instructions per normal on pipes that take a warp '
                          'instruction every second cycle (issue floor 0.035 ms); the MMAs of a unit take 1 400 of its ~3 400 cycles' if dm else
                          '; measured limiters (profiles/r02_k1_gen_*.txt, DESIGN 4c): the MMAs themselves -- three fp16 MMAs per '
                          'k-step for the two-term split, each fetching its 4 KiB factor tile from shared memory (55 cycles per '
                          'M128 x N32 x K16 MMA: floor 0.047 ms) -- and the noise generation on the CUDA cores (29.4 M normals: '
                          'Philox4x32-10 + Box-Muller + fp16 split = ~110 instructions per 4: issue floor 0.023 ms)' if gen else ''))
        k1['frac'] = k1['achieved'] / k1['peak']
        if gen and mv_in_k1:
            mv = dict(kernel='(no launch) Sigma^-1 mu is computed by warp 25 of K1 while the tiles run (mpb_sample_gp_kron_gen_mv); '
                             'this entry is the gap between the K1 and K2 launches', ms=float(stage[1]), bound='latency')
        else:
            mv = dict(kernel='prior_matvec_dof_kernel (Sigma^-1 mu)', ms=float(stage[1]), bound='latency')
        flop_k2 = FLOP_FK + FLOP_SDF + FLOP_GP + FLOP_IS
        a2 = flop_k2 * n_samp / (stage[2] * 1e-3) / 1e12
        slots = ncu_issue_slots('cost_eval_chain2')
        issue_peak = 148 * 4 * sm_mhz_peak * 1e6            # warp instructions / s: 4 schedulers per SM, one issue per cycle
        k2 = dict(kernel='cost_eval_chain2_kernel<7,10,2,false,true%s> (K2 packed: FK + collision + GP cost + IS dot, two waypoints per lane '
                         'on FFMA2 / FADD2 / FMUL2; sphere-only instance with the link-frame cull%s)'
                         % ((',DM', '; reads the dof-major rows in place') if dm else ('', '')), ms=float(stage[2]), bound='fp32',
                  achieved=a2, peak=fp32_measured, unit='TFLOP/s', frac=a2 / fp32_measured,
                  peak_nominal=fp32_nominal, frac_of_nominal=a2 / fp32_nominal, algorithmic_flop_per_sample=flop_k2,
                  hbm_gbs=M * 4 * n_samp / (stage[2] * 1e-3) / 1e9, hbm_frac=M * 4 * n_samp / (stage[2] * 1e-3) / 1e9 / pk['hbm'],
                  note='algorithmic flops (SURVEY 8d) count ALL 50 x 16 sphere pairs per waypoint; the broad phase proves most of '
                       'them zero and never evaluates them, which is why this figure can exceed 1: it measures the kernel against '
                       'the all-pairs formulation, not pipe utilisation.  The kernel is bound by the ISSUE rate -- see issue_view '
                       '(executed warp-instruction slots from the committed ncu capture / measured time / issue peak)')
        if slots:
            k2['issue_view'] = dict(bound='issue', achieved=slots['issue_slots_per_sample'] * n_samp / (stage[2] * 1e-3) / 1e9,
                                    peak=issue_peak / 1e9, unit='G warp-instruction slots/s',
                                    frac=slots['issue_slots_per_sample'] * n_samp / (stage[2] * 1e-3) / issue_peak,
                                    issue_slots_per_sample=slots['issue_slots_per_sample'],
                                    warp_instructions_per_sample=slots['warp_instructions_per_sample'], source=slots['source'],
                                    note='packed f32x2 instructions occupy two issue cycles (profiles/r02_ffma2_microbench.txt) and '
                                         'are counted twice; round 1 generic kernel: 9,467 warp instructions per sample')
        bytes_k3 = (rows_nonzero * M * 4 + 3 * P * M * 4 + 2 * n_samp * 4)
        a3 = bytes_k3 / (stage[3] * 1e-3) / 1e9
        k3 = dict(kernel='softmax_update_kernel (K3%s)' % (', dof-major rows' if dm else ''), ms=float(stage[3]), bound='latency', achieved=a3, peak=pk['hbm'], unit='GB/s',
                  frac=a3 / pk['hbm'], bytes_needed=bytes_k3, rows_with_nonzero_weight=rows_nonzero, rows_total=n_samp,
                  note='bytes actually needed: rows whose softmax weight is non-zero (nearly one-hot at T = 1) + means + costs / '
                       'weights; a 12 us launch + DRAM-latency chain, not a throughput kernel')
        tot = float(np.sum(stage))
        for k in (k1, mv, k2, k3):
            k['share'] = k['ms'] / tot
        return [k1, mv, k2, k3]

    kernels = kernel_table(stage_ms, k1_reads_eps=False)
    dom = max(kernels, key=lambda k: k['ms'])
    # The dominant kernel (K2) is bound by the instruction-issue rate, not by HBM or a math pipe: the headline fraction is
    # executed issue slots / issue peak (<= 1 by construction).  The SURVEY 8d algorithmic-flop figure counts all 800 sphere
    # pairs per waypoint, which the broad phase never evaluates -- it can exceed the FP32 peak and is reported next to it
    # (algorithmic_fp32_view), not as the roofline fraction.
    head = dom.get('issue_view')
    if head:
        roofline = dict(bound='issue', achieved=head['achieved'], peak=head['peak'], unit=head['unit'], frac=head['frac'],
                        issue_source=head['source'], issue_slots_per_sample=head['issue_slots_per_sample'],
                        algorithmic_fp32_view=dict(bound=dom['bound'], achieved=dom['achieved'], peak=dom['peak'], unit=dom['unit'],
                                                   frac=dom['frac'], note=dom['note']))
    else:
        roofline = dict(bound=dom['bound'], achieved=dom['achieved'], peak=dom['peak'], unit=dom['unit'], frac=dom['frac'])
    roofline.update(traffic=ncu_traffic(dom['kernel']), kernel=dom['kernel'], ms_per_launch=dom['ms'],
                    peak_source='issue peak = 148 SMs x 4 schedulers x clocks.max.sm (one warp instruction per scheduler and cycle; '
                                'FFMA2 / FADD2 / FMUL2 take two cycles: profiles/r02_ffma2_microbench.txt); FP32 FFMA rate of the '
                                'algorithmic view MEASURED in this run (mpb_bench_fp32_peak: 8 independent FMA chains per thread, '
                                'best of 6 launches); nominal 148 SMs x 128 lanes x 2 x clocks.max.sm = %.1f TFLOP/s' % fp32_nominal,
                    peak_nominal=fp32_nominal, frac_of_nominal=dom.get('frac_of_nominal'),
                    algorithmic_flop_per_sample=dom.get('algorithmic_flop_per_sample'),
                    hbm_view=dict(achieved=dom.get('hbm_gbs'), peak=pk['hbm'], unit='GB/s', frac=dom.get('hbm_frac'),
                                  peak_source=pk['source'], algorithmic_bytes_per_sample=M * 4),
                    kernels=kernels,
                    stage_events='CUDA events around the three launches inside the timed region, on every 5th timed step (an event '
                                 'record between two kernels blocks programmatic dependent launch; the other steps are the same '
                                 'three launches from ONE C call, planner.step_fused() = the iteration planner.optimize() issues)')

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total / K,
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=workload_config(world), clocks=clocks, gpu_launches=launches_per_step * K,
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=P * H * D * 4, d2h_bytes_per_step=P * H * D * 4,
                         steps=Ke, ms_per_step=ms_e2e / Ke,
                         note='planner.optimize(opt_iters=1) exactly as the reference is called (no noise argument: drawn in K1); '
                              'per step the initial particle trajectories come from pinned host memory and the optimised '
                              'trajectories go back to pinned host memory (one copy stream: the next upload / this download '
                              "overlap the kernels)",
                         injected_noise=dict(value=world * P * S * Kj / (ms_e2e_inj * 1e-3), unit=UNIT, ms_per_step=ms_e2e_inj / Kj,
                                             h2d_bytes_per_step=S * P * M * 4, d2h_bytes_per_step=P * H * D * 4,
                                             h2d_gb_per_s=S * P * M * 4 / (ms_e2e_inj / Kj * 1e-3) / 1e9,
                                             note='parity-style call planner.optimize(opt_iters=1, eps=<pinned host noise>): '
                                                  '117 MB of noise uploaded per step, PCIe bound')),
                injected_noise=dict(value=world * P * S * Ki / (ms_inj * 1e-3), unit=UNIT, ms_per_step=ms_inj / Ki, steps=Ki,
                                    note='the same step on injected noise resident in HBM (8 rotating 117 MB buffers): K1 reads eps '
                                         'instead of drawing it', kernels=kernel_table(stage_inj, k1_reads_eps=True)),
                roofline=roofline, collision_free_fraction_last_step=free_frac, parity_check=check)
    if first_attempt is not None:
        line['remeasured'] = dict(first_attempt_ms_per_step=first_attempt,
                                  reason='the first K timed steps took more than 1.06x the sum of their kernels (host-bound launch '
                                         'loop); the same K steps were timed once more and that second measurement is reported')
    if split_block is not None:
        line['sample_split'] = split_block
    if world == 1 and not args.no_other_configs:
        # the other BASELINE.json configs ("ms per planner iter"), informational: see bench_configs.py
        del planner
        torch.cuda.empty_cache()
        try:
            import bench_configs
            line['other_configs'] = bench_configs.run_other_configs(dev, quick=True)
        except Exception as exc:       # never lose the headline line to an informational extra
            line['other_configs'] = dict(error=repr(exc))
    if not args.no_cpu_baseline and world == 1:     # the CPU baseline is reported at N=1 only (rank 0)
        torch.cuda.empty_cache()
        r = cpu_reference_run(args.cpu_particles, steps=3, warmup=1, faithful=True)
        line['cpu_baseline'] = dict(
            value=r['value'], unit=UNIT, cores=torch.get_num_threads(), kind='port',
            sample=(f'{args.cpu_particles} particles x {S} samples x {H} waypoints per step, 3 timed steps + 1 warm-up; oracle '
                    'port in the reference cost model (per-iteration re-factorisation, batched mat-vec sampler, eager torch fp32)'),
            ms_per_step=r['ms_per_step'])
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
