#!/usr/bin/env python
"""Benchmark of the fused dual-NCDE training step (BASELINE.json metric: fwd+bwd samples/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sepsis] [--impl ours|reference]

A step = one pass of the hot path over one batch: dual-NCDE forward, loss, backward through the solver,
(N>1: flat-gradient all-reduce over NCCL) and the Adam update.  Default workload: PhysioNet-Sepsis (no OI)
shape, 1024 samples per GPU (weak scaling), synthetic data, random-init weights.
Prints ONE JSON line on rank 0 (see the keys below).  `--impl reference` times the reference's own CPU path on the
host cores: the unmodified reference from oracle/_ref (a copy made by oracle/make_ref.py; /root/reference itself does not
exist on the GPU box), or the restatement oracle/port.py when that copy is absent.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'dual_ncde_fwd_bwd_samples_per_sec'
WORKLOAD_TEXT = {
    'sepsis': 'PhysioNet Sepsis no-OI shape: ANCDE, 35 ch (34+time), 72 knots (times 1..72), h=hh=49, 4 layers, '
              'attention 20/20, soft timewise attention, rk4 step 1, backprop through the solver, batch 1024/GPU',
    'sepsis_oi': 'PhysioNet Sepsis OI shape: ANCDE, 69 ch, 72 knots, h=hh=49, 4 layers, soft timewise attention, '
                 'rk4 step 1, batch 1024/GPU',
    'mujoco': 'Mujoco shape: ANCDE_forecasting, 14 ch, 20 knots -> 5 steps, h=12, 2 layers, hard STE attention',
    'stock': 'Google-Stock shape: ANCDE_forecasting, 6 ch, 24 knots -> 1 step, h=12, 2 layers, hard STE attention',
    'uea': 'UEA CharacterTrajectories shape: ANCDE, 4 ch, 182 knots, h=40, 3 layers, soft timewise attention',
}


# ------------------------------------------------------------------------------------------ helpers
def algorithmic_macs(cfg):
    """MACs per sample per RK4 stage."""
    C, H, HH, A, AA, nl = cfg['C'], cfg['H'], cfg['HH'], cfg['A'], cfg['AA'], cfg['nl']
    f_hidden = C * AA + AA * A + (nl - 1) * A * A
    f_out = A * C * C
    g_hidden = H * HH + (nl - 1) * HH * HH
    g_out = HH * H * C
    return dict(f_hidden=f_hidden, f_out=f_out, f_contract=C * C, g_hidden=g_hidden, g_out=g_out,
                g_contract=H * C)


def kernel_flops(cfg, B):
    """Algorithmic FLOPs of one launch of each kernel (2 * MACs * stages * samples)."""
    m = algorithmic_macs(cfg)
    stages = 4 * (cfg['L'] - 1)
    bot = m['f_hidden'] + m['f_out'] + m['f_contract']
    top = m['g_hidden'] + m['g_out'] + m['g_contract']
    s = 2.0 * stages * B
    return {
        'ncde_fwd_bottom': s * bot, 'ncde_fwd_top': s * top,
        # recurrent backward: input-gradient pass of every layer (same MACs as the forward)
        'ncde_bwd_bottom': s * bot, 'ncde_bwd_top': s * top,
        # weight gradients: output layer, hidden layers
        'ncde_wgrad_bottom': s * m['f_out'], 'ncde_wgrad_top': s * m['g_out'],
        'ncde_hidden_wgrad_bottom': s * m['f_hidden'], 'ncde_hidden_wgrad_top': s * m['g_hidden'],
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def make_device_batch(workload, B, seed, device):
    import ancde_b200
    from ancde_b200 import synthetic
    bt = synthetic.make_batch(workload, B=B, seed=seed)
    times = bt['times'].to(device)
    coeffs = ancde_b200.natural_cubic_spline_coeffs(times, bt['X'].to(device))
    out = dict(times=times, coeffs=tuple(coeffs), final_index=bt['final_index'].to(device), y=bt['y'].to(device))
    if bt['static'] is not None:
        out['static'] = bt['static'].to(device)
    return out


def make_raw_device_batch(workload, B, seed, device):
    """The batch as the dataset holds it BEFORE the spline: raw X with NaNs (coefficients are made inside the step)."""
    from ancde_b200 import synthetic
    bt = synthetic.make_batch(workload, B=B, seed=seed)
    out = dict(times=bt['times'].to(device), X=bt['X'].to(device), final_index=bt['final_index'].to(device),
               y=bt['y'].to(device))
    if bt['static'] is not None:
        out['static'] = bt['static'].to(device)
    return out


def to_pinned_host(batch):
    host = {}
    for k, v in batch.items():
        if k == 'coeffs':
            host[k] = tuple(c.cpu().pin_memory() for c in v)
        else:
            host[k] = v.cpu().pin_memory()
    return host


def h2d(host, device):
    dev = {}
    nbytes = 0
    for k, v in host.items():
        if k == 'coeffs':
            dev[k] = tuple(c.to(device, non_blocking=True) for c in v)
            nbytes += sum(c.numel() * c.element_size() for c in v)
        else:
            dev[k] = v.to(device, non_blocking=True)
            nbytes += v.numel() * v.element_size()
    return dev, nbytes


# ------------------------------------------------------------------------------------------ CPU reference
class _ReferenceIVN(torch.nn.Module):
    """z0 = MLP(static) in front of the reference model, as experiments/sepsis_attentive.py:15-30 (that script itself
    cannot be imported: it pulls in the dataset code)."""

    def __init__(self, n_static, C, model):
        super().__init__()
        self.linear1 = torch.nn.Linear(n_static, 256)
        self.linear2 = torch.nn.Linear(256, C)
        self.model = model

    def forward(self, times, coeffs, final_index, slope, static, **kwargs):
        z0 = self.linear2(self.linear1(static).relu())
        return self.model(times, coeffs, final_index, slope, z0=z0, **kwargs)


def _h_prime_path():
    d = '/dev/shm' if os.path.isdir('/dev/shm') and os.access('/dev/shm', os.W_OK) else tempfile.gettempdir()
    return os.path.join(d, 'ancde_ref_h_prime_%d.npy' % os.getpid())


def cpu_reference_step_fn(workload, sample_B, adjoint=False):
    """Returns (one_step, kind, cores): one_step() runs forward + loss + backward of the reference on the host cores.

    kind "reference": the UNMODIFIED reference (`controldiffeq` + `experiments/models` from oracle/_ref, a byte-for-byte
    copy made by oracle/make_ref.py) on the restated torchdiffeq (oracle/torchdiffeq_shim.py), including its np.save /
    np.load of h' (to tmpfs) and its python-level stage loop.  kind "port": oracle/port.py, when no copy is present."""
    from ancde_b200 import synthetic
    from ancde_b200.train import task_loss
    cfg = synthetic.CONFIGS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(2021)
    bt = synthetic.make_batch(workload, B=sample_B, seed=2021)
    from oracle import port
    coeffs = port.natural_cubic_spline_coeffs(bt['times'], bt['X'])
    from oracle import reference_loader
    if reference_loader.available():
        path = _h_prime_path()
        model = reference_loader.make_reference_model(cfg, path)
        if cfg['static']:
            model = _ReferenceIVN(cfg['static'], cfg['C'], model)
        coeffs32 = tuple(c.float() for c in coeffs) if cfg['dtype'] == 'float32' else tuple(coeffs)

        def one():
            model.zero_grad(set_to_none=True)
            kw = dict(stream=cfg['stream'], adjoint=adjoint)
            with reference_loader.cpu_only():             # the reference hard-codes .cuda() on its vector fields
                if cfg['static']:
                    pred = model(bt['times'], coeffs32, bt['final_index'], 1.0, bt['static'], **kw)
                else:
                    pred = model(bt['times'], coeffs32, bt['final_index'], 1.0, **kw)
                loss = task_loss(cfg, pred, bt['y'])
                loss.backward()
            return float(loss.detach())
        return one, 'reference', cores

    from ancde_b200 import metamodel
    model, _, _ = metamodel.model_from_config(cfg, file='cpu')
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    sdm = {k[len('model.'):]: v for k, v in sd.items() if k.startswith('model.')} if cfg['static'] else sd

    def one():
        for v in sd.values():
            v.grad = None
        z0 = port.ivn_z0(sd, bt['static']) if cfg['static'] else None
        res = port.dual_ncde_forward(sdm, cfg, bt['times'], coeffs, bt['final_index'], 1.0, stream=cfg['stream'],
                                     z0=z0, adjoint=adjoint)
        loss = task_loss(cfg, res['pred'], bt['y'])
        loss.backward()
        return float(loss.detach())
    return one, 'port', cores


def cpu_reference_samples_per_sec(workload, sample_B, steps, warmup, adjoint=False):
    """fwd+bwd of the reference's CPU path with every host thread: (samples/s, seconds per step, cores, kind)."""
    one, kind, cores = cpu_reference_step_fn(workload, sample_B, adjoint)
    for _ in range(warmup):
        one()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return sample_B / sec, sec, cores, kind


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path, same workload, same batch, same steps.
    A step is the full batch unless one step is so slow that `steps + warmup` of them would take more than ~4 minutes;
    then every step is a bounded sample (a power-of-two fraction of the batch), and the line says so."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    B = args.batch
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    one, kind, cores = cpu_reference_step_fn(args.workload, B, args.adjoint)
    t0 = time.perf_counter()
    one()                                            # first warm-up step, timed to size the run
    first = time.perf_counter() - t0
    budget = 240.0
    sample_B = B
    while first * (sample_B / B) * (steps + warmup) > budget and sample_B > 64:
        sample_B //= 2
    if sample_B != B:
        one, kind, cores = cpu_reference_step_fn(args.workload, sample_B, args.adjoint)
        one()
    for _ in range(warmup - 1):
        one()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    value = sample_B / sec
    what = 'the unmodified reference (oracle/_ref: controldiffeq + experiments/models, incl. its np.save/np.load of h\') ' \
           'on the restated torchdiffeq' if kind == 'reference' else 'the CPU restatement oracle/port.py'
    sample = '%s; %d samples of the %s workload per step (%s), fwd+bwd, median of %d steps after %d warm-up' % (
        what, sample_B, args.workload, 'the full batch' if sample_B == B else 'bounded sample of the %d-sample batch' % B,
        steps, warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD_TEXT[args.workload], 'name': args.workload, 'batch_per_gpu': B,
                   'batch_per_step': sample_B, 'adjoint': bool(args.adjoint)},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    try:
        os.unlink(_h_prime_path())
    except OSError:
        pass
    return 0


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    import ancde_b200
    from ancde_b200 import _lib, synthetic
    from ancde_b200.train import DualNcdeTrainer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the ANCDE B200 path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    if world != args.gpus and rank == 0:
        sys.stderr.write('warning: --gpus %d but WORLD_SIZE=%d\n' % (args.gpus, world))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    cfg = synthetic.CONFIGS[args.workload]
    B = args.batch
    trainer = DualNcdeTrainer(args.workload, device, seed=2021, adjoint=args.adjoint)
    # several distinct resident batches, cycled, so no step re-reads inputs the previous one left in L2
    n_res = 4
    batches = [make_device_batch(args.workload, B, 2021 + 1000 * rank + i, device) for i in range(n_res)]
    # e2e leg: the step starts from the raw paths X (NaN = missing) and builds the spline coefficients on the GPU
    # (datasets_common.coefficients_on_the_fly): 1/4 of the host -> device bytes of precomputed coefficients
    raw_batches = [make_raw_device_batch(args.workload, B, 2021 + 1000 * rank + i, device) for i in range(2)]
    host_batches = [{k: v for k, v in to_pinned_host(b).items() if k != 'times'} for b in raw_batches]
    fp32_peak = _lib.fp32_peak_tflops(device)

    def timed(fn, steps, warmup, profile=False):
        for i in range(warmup):
            fn(i)
        barrier()
        if profile:
            _lib.profile_enable(True)
        launches0 = _lib.launches_total()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        prof = _lib.profile_read() if profile else []
        if profile:
            _lib.profile_enable(False)
        t = torch.tensor([ms], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), _lib.launches_total() - launches0, prof

    # ---- per-kernel timings (CUDA events around every library launch) from an eager pass; also counts the launches.
    # The weight-gradient kernels stay IN STREAM for this pass: on the library's side stream they share the GPU with the
    # other network's recurrent backward, and an event pair around a kernel that waits for SMs measures the wait
    # (hidden_wgrad_top 0.29 -> 1.9 ms, bwd_bottom 1.75 -> 2.3 ms), not the kernel.
    from ancde_b200 import train as _train
    _async = _train.ASYNC_WGRAD
    _train.ASYNC_WGRAD = False
    _, launches_prof, prof = timed(lambda i: trainer.step(batches[i % n_res]), 3, max(3, args.warmup), profile=True)
    _train.ASYNC_WGRAD = _async
    launches_per_step = launches_prof // 3
    prof_steps = 3

    # ---- whole steps captured in CUDA graphs (one per resident batch); eager when capture is not possible
    graphed = None
    if not args.no_graph:
        try:
            graphed = [trainer.capture(b) for b in batches]
        except Exception as e:                                   # noqa: BLE001 -- report and measure eagerly
            sys.stderr.write('CUDA graph capture failed (%s: %s); measuring eager steps\n' % (type(e).__name__, e))
            graphed = None
            torch.cuda.synchronize(device)

    def run_step(i):
        if graphed is not None:
            return graphed[i % n_res].replay()
        return trainer.step(batches[i % n_res])

    # ---- value: inputs resident in HBM
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, _, _ = timed(run_step, args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else {}
    value = world * B * args.steps / (ms * 1e-3)
    launches = launches_per_step * args.steps

    # ---- e2e: host (pinned) inputs through the public API, H2D + D2H inside the timed region
    h2d_bytes = [0]
    d2h_bytes = [0]

    # Double-buffered input pipeline: while the captured step i runs on batch buffer i % 2, the copy stream moves the
    # pinned host batch of step i + 1 into the other buffer (after the replay that last read that buffer has finished).
    # Every step's host -> device copy and its device -> host loss read happen inside the timed region.
    copy_stream = torch.cuda.Stream(device)
    ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
    ev_consumed = [torch.cuda.Event(), torch.cuda.Event()]

    # The input pipeline of the captured step: the copy stream moves the raw paths of step i + 1 to the device AND turns them into
    # spline coefficients there (one kernel, written straight into the static coefficient tensors of the captured step), all of it
    # next to the compute of step i (the recurrent kernels leave 20 SMs idle).  graphed_raw: the captured steps whose static inputs
    # the pipeline fills -- the coefficient-input graphs of the `value` leg.
    graphed_raw = None
    x_dev = [b['X'] for b in raw_batches]
    if graphed is not None:
        graphed_raw = graphed[:2]

    def prefetch(i):
        sl = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_consumed[sl])
            n = 0
            for k, v in host_batches[sl].items():
                dst = x_dev[sl] if k == 'X' else graphed_raw[sl].batch[k]
                dst.copy_(v, non_blocking=True)
                n += v.numel() * v.element_size()
            ancde_b200.natural_cubic_spline_coeffs(raw_batches[sl]['times'], x_dev[sl], out=graphed_raw[sl].batch['coeffs'])
            h2d_bytes[0] = n
            ev_copied[sl].record(copy_stream)

    def e2e_step(i):
        if graphed is not None:                   # pinned host batch -> the graph's static inputs -> replay
            sl = i % 2
            cur = torch.cuda.current_stream(device)
            cur.wait_event(ev_copied[sl])
            loss = graphed_raw[sl].replay()
            ev_consumed[sl].record(cur)
            prefetch(i + 1)                       # enqueued while the GPU runs step i (the host blocks on the loss below)
        else:
            dev_batch, nb = h2d(host_batches[i % len(host_batches)], device)
            dev_batch['times'] = raw_batches[0]['times']
            h2d_bytes[0] = nb
            loss = trainer.step(dev_batch)
        loss_host = loss.to('cpu')                # device -> host read of the step's result
        d2h_bytes[0] = loss_host.numel() * loss_host.element_size()
        return loss_host

    if graphed is not None:
        torch.cuda.synchronize(device)
        prefetch(0)
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e, _, _ = timed(e2e_step, e2e_steps, 2)
    e2e_value = world * B * e2e_steps / (ms_e2e * 1e-3)

    def finish():
        """Orderly teardown: the captured graphs (they hold NCCL kernels) are destroyed before the process group.  In
        round 1 destroy_process_group / interpreter exit could block forever with live graphs on 2 x B200, so a watchdog
        thread ends the process if the teardown has not finished after 20 s -- every result is printed by then."""
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            import threading
            dog = threading.Timer(20.0, lambda: os._exit(0))
            dog.daemon = True
            dog.start()
            graphed_all = [g for g in (graphed or [])]
            for g in graphed_all:
                g.graph.reset()
            torch.cuda.synchronize(device)
            dist.destroy_process_group()
            dog.cancel()
        return 0

    # ---- the other BASELINE.json configs, short runs in the same process on every rank (whole steps from a CUDA graph,
    # inputs resident): Sepsis OI weak (1024 / GPU) and strong (1024 in total), Sepsis with the adjoint gradient flow,
    # UEA, Mujoco, Stock.  `value` = samples of all ranks / max-over-ranks device time, like the headline.
    def measure_other(workload, B_gpu, adjoint, steps=5, warmup=3):
        out = {'workload': workload, 'batch_per_gpu': B_gpu, 'global_batch': B_gpu * world, 'adjoint': bool(adjoint)}
        try:
            tr = DualNcdeTrainer(workload, device, seed=2021, adjoint=adjoint)
            bs = [make_device_batch(workload, B_gpu, 4021 + 1000 * rank + i, device) for i in range(2)]
            gs = [tr.capture(b) for b in bs]
            ms_o, _, _ = timed(lambda i: gs[i % 2].replay(), steps, warmup)
            fl = sum(kernel_flops(synthetic.CONFIGS[workload], B_gpu).values())
            out.update(value=round(world * B_gpu * steps / (ms_o * 1e-3), 1), unit='samples/s',
                       ms_per_step=round(ms_o / steps, 4), steps=steps, warmup=warmup,
                       step_frac=round(fl * steps / (ms_o * 1e-3) / 1e12 / fp32_peak, 4))
            for g_ in gs:
                g_.graph.reset()
            del gs, bs, tr
        except Exception as e:                                   # noqa: BLE001 -- one config must not take the line down
            out['error'] = '%s: %s' % (type(e).__name__, str(e)[:200])
        torch.cuda.synchronize(device)
        torch.cuda.empty_cache()
        return out

    other_configs = []
    if not args.no_other_configs:
        full = synthetic.CONFIGS
        todo = [('sepsis_oi', full['sepsis_oi']['B'], False, 'weak')]
        if world > 1:
            todo.append(('sepsis_oi', max(1, full['sepsis_oi']['B'] // world), False, 'strong'))
            todo.append(('sepsis', max(1, full['sepsis']['B'] // world), False, 'strong'))
        todo += [('sepsis', full['sepsis']['B'], True, 'weak'), ('uea', full['uea']['B'], False, 'weak'),
                 ('mujoco', full['mujoco']['B'], False, 'weak'), ('stock', full['stock']['B'], False, 'weak')]
        for wl, bg, adj, scaling in todo:
            if wl == args.workload and bg == B and adj == bool(args.adjoint):
                continue
            r = measure_other(wl, bg, adj)
            r['scaling'] = scaling
            other_configs.append(r)

    barrier()                       # the last collective of every rank
    if rank != 0:
        return finish()

    # ---- per-kernel roofline from the CUDA-event timings taken inside the timed region
    flops = kernel_flops(cfg, B)
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of one `ncu --set full` capture of this
    # workload at this batch (profiles/r02_ncu_set_full.txt); a measured constant, only valid for the captured shape
    NCU_TRAFFIC = {('sepsis', 1024): {
        'ncde_fwd_bottom': 0.0415e9 + 1.6618e9, 'ncde_fwd_top': 0.0947e9 + 2.3874e9,
        'ncde_bwd_top': 2.4501e9 + 0.2799e9, 'ncde_wgrad_top': 2.2732e9 + 0.0050e9, 'ncde_hidden_wgrad_top': 0.4597e9 + 0.0041e9,
        'ncde_bwd_bottom': 1.6847e9 + 0.1542e9, 'ncde_wgrad_bottom': 1.5843e9 + 0.0048e9,
        'ncde_hidden_wgrad_bottom': 0.2542e9 + 0.0045e9}}
    traffic = NCU_TRAFFIC.get((args.workload, B), {})
    agg = {}
    for name, t in prof:
        a = agg.setdefault(name, [0.0, 0])
        a[0] += t
        a[1] += 1
    total_kernel_ms = sum(a[0] for a in agg.values())
    kernels = []
    for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        avg = t / n
        k = {'name': name, 'launches': n, 'avg_ms': round(avg, 4), 'share_of_kernel_time': round(t / total_kernel_ms, 4)}
        if name in flops:
            k['algorithmic_gflop_per_launch'] = round(flops[name] / 1e9, 3)
            k['achieved_tflops'] = round(flops[name] / (avg * 1e-3) / 1e12, 3)
            k['frac_of_fp32_peak'] = round(flops[name] / (avg * 1e-3) / 1e12 / fp32_peak, 4)
        if name in traffic:
            k['dram_bytes_per_launch'] = traffic[name]
        kernels.append(k)
    dom = next((k for k in kernels if 'achieved_tflops' in k), None)
    total_flops = sum(flops.values())
    roofline = None
    if dom is not None:
        roofline = {'bound': 'fp32', 'kernel': dom['name'], 'achieved': dom['achieved_tflops'],
                    'peak': round(fp32_peak, 2), 'unit': 'TFLOP/s', 'frac': dom['frac_of_fp32_peak'],
                    'peak_source': 'FFMA microbenchmark (ancde_fp32_peak_kernel) measured in this run; '
                                   'MEASURED_PEAKS.json holds no FP32 figure',
                    'traffic': traffic.get(dom['name']),
                    'traffic_source': 'profiles/r02_ncu_set_full.txt (bytes per launch)' if dom['name'] in traffic else None,
                    'step_frac': round(total_flops * args.steps / (ms * 1e-3) / 1e12 / fp32_peak, 4),
                    'step_algorithmic_gflop': round(total_flops / 1e9, 2)}

    # ---- HBM roofline of the output-layer weight-gradient kernel (tcgen05; reads the saved tanh outputs exactly once)
    roofline_wgrad = None
    wg = next((k for k in kernels if k['name'] == 'ncde_wgrad_top'), None)
    if wg is not None:
        stages = 4 * (cfg['L'] - 1)
        c4 = (cfg['C'] + 3) // 4 * 4
        g_bytes = stages * B * cfg['H'] * c4 * 4                                  # tanh outputs, read once
        small = stages * B * (cfg['HH'] + cfg['H'] + c4) * 4                      # last hidden activation, cotangent, dU
        nbytes = g_bytes + small
        peak, src = hbm_peak()
        ach = nbytes / (wg['avg_ms'] * 1e-3) / 1e9
        roofline_wgrad = {'bound': 'hbm', 'kernel': 'ncde_wgrad_top', 'achieved': round(ach, 1), 'peak': peak, 'unit': 'GB/s',
                          'frac': round(ach / peak, 4), 'traffic': traffic.get('ncde_wgrad_top'), 'peak_source': src,
                          'avg_ms': wg['avg_ms'],
                          'algorithmic_bytes_per_launch': nbytes,
                          'tensor_pipe': 'tcgen05.mma kind::f16, FP16 hi/lo x3 = FP32-grade; accumulator in tensor memory'}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, sec, cores, kind = cpu_reference_samples_per_sec(args.workload, args.cpu_sample, 3, 1, args.adjoint)
        cpu = {'value': round(v, 2), 'unit': 'samples/s', 'cores': cores, 'kind': kind,
               'sample': '%d samples of the same workload, fwd+bwd, median of 3 after 1 warm-up (%.2f s each)'
                         % (args.cpu_sample, sec)}

    line = {
        'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD_TEXT[args.workload], 'name': args.workload, 'batch_per_gpu': B,
                   'global_batch': B * world, 'adjoint': bool(args.adjoint), 'optimizer': 'Adam (fused) inside the step',
                   'cuda_graph': graphed is not None,
                   'parallelism': 'dp%d' % world,
                   'l2': '%d distinct resident batches cycled; each step streams ~%.1f GB of saved activations '
                         'through HBM (>> 126 MB L2)' % (n_res, saved_gb(cfg, B))},
        'clocks': clk,
        'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d_bytes[0],
                'd2h_bytes_per_step': d2h_bytes[0], 'ms_per_step': ms_e2e / e2e_steps, 'steps': e2e_steps,
                'input_pipeline': ('double-buffered: the copy of the raw paths of step i+1 and their spline coefficients (one kernel on the '
                                   'copy stream, written into the captured step\'s static inputs) overlap the compute of step i')
                                  if graphed is not None else 'serial'},
        'gpu_launches': launches,
        'kernel_profile_steps': prof_steps,
        'roofline': roofline,
        'roofline_spline': spline_roofline(device, cfg) if world == 1 else None,
        'roofline_wgrad': roofline_wgrad,
        'kernels': kernels,
        'cpu_baseline': cpu,
        'other_configs': other_configs,
    }
    print(json.dumps(line))
    return finish()


def hbm_peak():
    peak, src = 6650.0, 'fallback 6.65 TB/s (B200_PROFILING.md)'
    try:
        mp = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        peak, src = float(mp['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs'
    except (OSError, KeyError, ValueError):
        pass
    return peak, src


def spline_roofline(device, cfg, n_paths_target=2 ** 20):
    """HBM roofline of the spline-coefficient kernel on a dataset-split sized input of the workload's shape."""
    import ancde_b200
    from ancde_b200 import _lib
    L, C = cfg['L'], cfg['C']
    N = max(1024, n_paths_target // C)
    g = torch.Generator(device='cpu').manual_seed(7)
    X = (0.1 * torch.randn(N, L, C, generator=g)).cumsum(1)
    X[torch.rand(N, L, C, generator=g) < 0.9] = float('nan')
    X = X.to(device)
    times = (torch.arange(L, dtype=torch.float32) + cfg['t0']).to(device)
    for _ in range(3):
        ancde_b200.natural_cubic_spline_coeffs(times, X)
    torch.cuda.synchronize(device)
    _lib.profile_enable(True)
    reps = 10
    for _ in range(reps):
        ancde_b200.natural_cubic_spline_coeffs(times, X)
    torch.cuda.synchronize(device)
    ts = [t for name, t in _lib.profile_read() if name.startswith('spline_coeffs')]
    _lib.profile_enable(False)
    ms = sum(ts) / len(ts)
    nbytes = N * C * (L + 4 * (L - 1)) * 4           # algorithmic: read X once, write a, b, two_c, three_d once
    peak, src = hbm_peak()
    achieved = nbytes / (ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of exactly this input
    # (profiles/r02_spline_ncu.txt); only valid for the captured shape
    traffic = 1.5751e9 if (N, L, C) == (29959, 72, 35) else None
    return {'bound': 'hbm', 'kernel': 'spline_coeffs_f32', 'achieved': round(achieved, 1), 'peak': peak, 'unit': 'GB/s',
            'frac': round(achieved / peak, 4), 'traffic': traffic,
            'traffic_source': 'profiles/r02_spline_ncu.txt (bytes per launch)' if traffic else None,
            'peak_source': src, 'avg_ms': round(ms, 4),
            'paths': N * C, 'knots': L, 'algorithmic_bytes_per_launch': nbytes,
            'input': '%d x %d x %d float32, 90%% of entries missing' % (N, L, C)}


def saved_gb(cfg, B):
    stages = 4 * (cfg['L'] - 1)
    c4 = lambda c: (c + 3) // 4 * 4
    return stages * B * 4 * (cfg['H'] * c4(cfg['C']) + cfg['C'] * c4(cfg['C'])) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default='sepsis', choices=sorted(WORKLOAD_TEXT))
    ap.add_argument('--batch', type=int, default=None, help='samples per GPU (default: the workload\'s batch)')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--adjoint', action='store_true', help="reference's adjoint=True gradient flow")
    ap.add_argument('--cpu-sample', type=int, default=512, help='samples per CPU-baseline step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true', help='skip the short runs of the other BASELINE configs')
    ap.add_argument('--no-graph', action='store_true', help='launch every step eagerly instead of replaying CUDA graphs')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    from ancde_b200 import synthetic
    if args.batch is None:
        args.batch = synthetic.CONFIGS[args.workload]['B']
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
