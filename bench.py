#!/usr/bin/env python
"""bench.py — gated-GCRNN sequences/second, forward+backward (BASELINE.json metric).

Workload (N=1): cfg3 of SURVEY.md §8d — synthetic dense graph N=1024 (density 0.3, S = W/lambda_max), F=64 state
features, G=1 input feature, K=5 taps, T=64, global batch 4096 sequences, time-gated GGCRNNCell, random-init
weights (reference init, seed 0), X ~ N(0,1), h0 = 0, dH = ones.  A "step" = forward + backward of the whole
global batch (in micro-batches that fit HBM) + one all-reduce of the parameter-gradient bucket when N > 1
(batch sharded over ranks: strong scaling, configs[3]).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference        # the reference algorithm on the host CPUs (oracle port), bounded sample
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SEQ_FWD_BWD = 82.82e9      # SURVEY.md §8d, cfg3 (G=1): algorithmic flops per sequence, fwd+bwd
CFG3 = dict(N=1024, F=64, G=1, K=5, T=64, B=4096, density=0.3)


_JSON_OUT = None


def json_only_stdout():
    """Multi-rank runs: libraries (NCCL's version banner) write to file descriptor 1 behind Python's back.  Keep a private copy of
    the real stdout for the ONE JSON line and point fd 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(out):
    f = _JSON_OUT or sys.stdout
    f.write(json.dumps(out) + '\n')
    f.flush()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], bf16=d['bf16_tflops'], bf16_sustained=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src='fallback')


def ncu_traffic(files, kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of the first matching kernel row in a committed
    `ncu --set full ... --page raw --csv` export under profiles/ (captured with `bench.py --once`, same launch shape)."""
    import csv
    unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    for f in files:
        path = os.path.join(ROOT, 'profiles', f)
        if not os.path.isfile(path):
            continue
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], dict(zip(rows[0], rows[1]))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            if kernel_substr in d.get('Kernel Name', '') and 'dram__bytes_read.sum' in d:
                tot = sum(float(d[k].replace(',', '')) * unit.get(units[k], 1.0) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
                return tot, f
    return None, None


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port, fp64 as the reference scripts run it) on a bounded sample
# ------------------------------------------------------------------------------------------------------
def cpu_sample(cfg, Bs, Ts, repeats, warm=1, S=None, tg=True, sg=None):
    """Time forward+backward of the time-gated cell on the host CPUs, fp64 (as the reference scripts run it).

    kind 'reference': the UNMODIFIED reference `Utils.graphML.GGCRNNCell` from the git-ignored copy under baseline/_ref
    (staged by `__graft_entry__.build()`); kind 'port': the oracle's restatement when no copy travelled with the repo."""
    import torch
    from oracle import gcrnn_oracle as orc, ref_shim
    import gated_gcrnns_b200 as gg
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if S is None:
        S = gg.graphs.dense_random(cfg['N'], cfg['density'], seed=0).double()
    X = torch.randn(Bs, Ts, cfg['G'], cfg['N'], dtype=torch.float64)
    h0 = torch.zeros(Bs, cfg['F'], cfg['N'], dtype=torch.float64)
    dH = torch.ones(Bs, Ts, cfg['F'], cfg['N'], dtype=torch.float64)
    torch.manual_seed(0)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        if ref_shim.available():
            kind = 'reference'
            gml = ref_shim.load()
            cell = gml.GGCRNNCell(cfg['G'], cfg['F'], cfg['K'], cfg['K'], torch.tanh, tg, sg, 1, True)
            cell.addGSO(S)

            def run():
                cell.zero_grad()
                torch.autograd.backward(cell(X, h0), dH)
        else:
            kind = 'port'
            p = orc.init_cell_params(cfg['G'], cfg['F'], cfg['K'], cfg['K'], cfg['N'], tg, sg, 1, True)

            def run():
                orc.cell_forward_backward(p, S, X, h0, dH, tg, sg)
        times = []
        for i in range(warm + repeats):
            t0 = time.perf_counter()
            run()
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    finally:
        torch.set_default_dtype(prev)
    return times, cores, kind


def cpu_baseline_dict(cfg, times, cores, Bs, Ts, kind):
    best = min(times)
    seqs = Bs / (best * cfg['T'] / Ts)          # linear extrapolation in T (recurrence cost is linear in B*T)
    what = ('unmodified reference Utils.graphML.GGCRNNCell (copy of the reference tree staged under baseline/_ref)' if kind == 'reference'
            else 'oracle port of the reference GGCRNNCell')
    return dict(value=seqs, unit='sequences/s', cores=cores, kind=kind,
                sample=f'{what}, fp64, torch CPU with {cores} threads, fwd+bwd on cfg3 shapes with '
                       f'B={Bs}, T={Ts} of {cfg["T"]}; min of {len(times)} runs = {best:.3f} s, extrapolated linearly in T')


def run_reference(args):
    cfg = dict(CFG3)
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    Bs, Ts = args.cpu_batch, args.cpu_T
    times, cores, kind = cpu_sample(cfg, Bs, Ts, args.steps, warm=min(max(args.warmup, 1), 2))
    ms = 1e3 * statistics.mean(times)
    seqs = Bs / (statistics.mean(times) * cfg['T'] / Ts)
    cb = cpu_baseline_dict(cfg, times, cores, Bs, Ts, kind)
    cb['value'] = seqs
    out = dict(impl='reference', metric='GCRNN sequences/sec fwd+bwd', value=seqs, unit='sequences/s', n_gpus=args.gpus,
               steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling='strong',
               vs_baseline=None, dtype='f64', data='synthetic',
               config=dict(workload='cfg3: dense N=1024 F=64 G=1 K=5 T=64 time-gated GGCRNNCell, global batch 4096',
                           sample=f'B={Bs}, T={Ts} per step on the host CPUs'),
               cpu_baseline=cb, e2e=dict(value=seqs, unit='sequences/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    emit(out)


# ------------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------------
class Clocks(threading.Thread):
    REASONS = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
               0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown',
               0x100: 'display_clock_setting'}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.mask, self.stop_flag, self.max_mhz = index, [], 0, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=['unavailable'])
        reasons = [n for b, n in self.REASONS.items() if self.mask & b and n != 'gpu_idle']
        return dict(sm_mhz=statistics.median(self.samples), sm_max_mhz=self.max_mhz, reasons=reasons)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
PREC_DTYPE = {'fp32': 'f32', 'bf16': 'bf16', 'bf16x2': 'bf16x2 (split bf16 hi+lo operands, fp32 accumulate)'}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gated_gcrnns_b200 as gg
    from gated_gcrnns_b200 import _lib, graph as ggraph

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        # the product path: ONE all-reduce per step of every parameter gradient on one flat bucket
        # (gated_gcrnns_b200.dist.allreduce_gradients), through torch.distributed or the library's own NCCL transport
        gg.dist.enable(native=bool(args.native_allreduce))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'

    cfg = dict(CFG3)
    cfg['B'] = args.batch
    N, F, G, K, T = cfg['N'], cfg['F'], cfg['G'], cfg['K'], cfg['T']
    lo, hi = gg.dist.shard_range(cfg['B'], rank, world)
    Bl = hi - lo
    mb = min(args.microbatch, Bl)
    assert Bl % mb == 0, 'local batch must be a multiple of the micro-batch'
    gg.set_precision(args.precision)
    S = gg.graphs.dense_random(N, cfg['density'], seed=0)
    torch.manual_seed(0)
    sg = 'node' if getattr(args, 'cfg3_spatial', 'none') == 'node' else None      # extra: the same dense config with node gates on top
    cell = gg.GGCRNNCell(G, F, K, K, torch.tanh, True, sg, 1, True)
    cell.addGSO(S)
    cell = cell.to(dev)
    used = [dict(cell.named_parameters())[n] for _, _, n in gg.cell_param_slots(True, sg, True)]
    gen = torch.Generator(device='cpu').manual_seed(1234 + rank)
    X_host = torch.randn(Bl, T, G, N, generator=gen).pin_memory()
    h0_host = torch.zeros(mb, F, N).pin_memory()
    X_dev = X_host.to(dev)
    h0_dev = torch.zeros(mb, F, N, device=dev)
    dH = torch.ones(mb, T, F, N, device=dev)
    L = _lib.lib()
    if args.gemm_pair is not None:
        gg.options.set('gemm_pair', args.gemm_pair)
    if args.bwd_fused is not None:
        gg.options.set('bwd_fused', args.bwd_fused)
    pair = gg.options.get('gemm_pair')

    # e2e leg: pinned host -> device copies of every micro-batch's X and h0 run on a side stream, double-buffered, so the
    # PCIe transfer of micro-batch i+1 overlaps the kernels of micro-batch i (all of it inside the timed region); the first
    # micro-batch of step s+1 is prefetched while step s computes (what a training loop's data loader does)
    copy_stream = torch.cuda.Stream(device=dev)
    xbuf = [torch.empty(mb, T, G, N, device=dev) for _ in range(2)]
    hbuf = [torch.empty(mb, F, N, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    ctr = {'mb': 0, 'prefetched': False}

    def prefetch(i, slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            xbuf[slot].copy_(X_host[i:i + mb], non_blocking=True)
            hbuf[slot].copy_(h0_host, non_blocking=True)
            ready[slot].record(copy_stream)

    def step(host_inputs, prefetch_next=False, nb=None):
        nb = Bl if nb is None else nb
        for p in used:
            p.grad = None
        main = torch.cuda.current_stream()
        if host_inputs and not ctr['prefetched']:
            for sl in range(2):
                freed[sl].record(main)
            prefetch(0, ctr['mb'] & 1)
        ctr['prefetched'] = False
        for i in range(0, nb, mb):
            if host_inputs:
                slot = ctr['mb'] & 1
                if i + mb < nb:
                    prefetch(i + mb, slot ^ 1)
                elif prefetch_next:
                    prefetch(0, slot ^ 1)
                    ctr['prefetched'] = True
                main.wait_event(ready[slot])
                x, h = xbuf[slot], hbuf[slot]
                ctr['mb'] += 1
            else:
                x, h = X_dev[i:i + mb], h0_dev
            H = cell(x, h)
            torch.autograd.backward(H, dH)
            del H
            if host_inputs:
                freed[slot].record(main)
        if world > 1:
            bucket = gg.dist.allreduce_gradients(used, op='sum')      # product path: one collective per step
        else:
            bucket = torch.cat([p.grad.reshape(-1) for p in used])
        if host_inputs:
            return bucket.to('cpu', non_blocking=False)          # D2H read of the step's result
        return bucket

    def timed(host_inputs, steps, warmup):
        for w in range(warmup):
            step(host_inputs, prefetch_next=host_inputs)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clk = Clocks(local)
        clk.start()
        l0 = L.gcrnn_debug_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s_ in range(steps):
            step(host_inputs, prefetch_next=host_inputs and s_ + 1 < steps)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = L.gcrnn_debug_launch_count() - l0
        c = clk.result()
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), launches, c

    if args.once:                    # profiling aid (ncu --set full): one forward+backward of one micro-batch, no timing
        step(False, nb=mb)
        torch.cuda.synchronize()
        return
    ms_dev, launches, clocks = timed(False, args.steps, args.warmup)
    ms_e2e, _, _ = timed(True, args.steps, max(1, args.warmup - 2))
    seqs = cfg['B'] * args.steps / (ms_dev * 1e-3)
    seqs_e2e = cfg['B'] * args.steps / (ms_e2e * 1e-3)

    # ---- the other precisions of the same workload (device-resident inputs, short run): seq/s for EVERY mode ---------------
    modes = {args.precision: dict(value=seqs, ms_per_step=ms_dev / args.steps, steps=args.steps)}
    for other in [m for m in args.also if m != args.precision]:
        gg.set_precision(other)
        o_steps = 2 if other != 'fp32' else 1
        if other == 'fp32':            # the exact path is far slower at cfg3: time ONE micro-batch of 16 sequences
            xs, hs, ds = X_dev[:16], h0_dev[:16], dH[:16]
            for _ in range(2):
                for p in used:
                    p.grad = None
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                torch.autograd.backward(cell(xs, hs), ds)
                e1.record(); torch.cuda.synchronize()
            modes[other] = dict(value=16 * world / (e0.elapsed_time(e1) * 1e-3), ms_per_step=e0.elapsed_time(e1), steps=1,
                                note='fp32 exact path (CUDA cores, sparse shift over the 30 %-dense operator): one 16-sequence micro-batch per GPU, second of two runs')
        else:
            ms_o, _, _ = timed(False, o_steps, 1)
            modes[other] = dict(value=cfg['B'] * o_steps / (ms_o * 1e-3), ms_per_step=ms_o / o_steps, steps=o_steps)
    gg.set_precision(args.precision)

    # ---- parity of the timed mode, measured in-run (outside the timed region) on a B = 2 slice at cfg3's own T = 64 ---------
    # reference = this library's fp32 exact path (itself held to the fp64 oracle by tests/); errors are max-norm relative to max|ref|
    parity = None
    if rank == 0 and args.precision != 'fp32' and not args.no_parity:
        def fb(prec, Tg):
            gg.set_precision(prec)
            for p in used:
                p.grad = None
            xs, hs = X_dev[:2], h0_dev[:2]
            d2 = torch.zeros(2, T, F, N, device=dev)
            g2 = torch.Generator(device='cpu').manual_seed(99)
            d2[:, :Tg] = torch.randn(2, Tg, F, N, generator=g2).to(dev)
            Hh = cell(xs, hs)
            torch.autograd.backward(Hh, d2)
            return Hh.detach(), torch.cat([p.grad.reshape(-1) for p in used]).clone()
        Hr, gr16 = fb('fp32', 16)
        Hm, gm16 = fb(args.precision, 16)
        hmax = Hr.abs().max()
        curve = ((Hm - Hr).abs().amax(dim=(0, 2, 3)) / hmax).tolist()
        parity = dict(mode=args.precision, reference='fp32 exact path of this library (CUDA cores, sparse shift; pinned to the fp64 oracle by tests/)',
                      T=T, B=2, init='reference init, seed 0', max_rel_H_first16=max(curve[:16]), max_rel_H=max(curve),
                      max_rel_grad_T16=((gm16 - gr16).abs().max() / gr16.abs().max()).item(),
                      H_curve_every8=[curve[t] for t in range(0, T, 8)] + [curve[-1]],
                      note='the recurrence is chaotic under the reference init: fp32 itself drifts from fp64 (tests/test_gpu_tc.py::'
                           'test_tc_cfg3_full_horizon_vs_oracle, profiles/r02_horizon_*.json); gradients are those of the T = 16 prefix problem')
        gg.set_precision(args.precision)

    # ---- data-parallel arithmetic check (N > 1): sharded + all-reduced gradients vs a single-rank recompute ------------------
    grad_check = None
    if world > 1:
        Bc = 64
        gc = torch.Generator(device='cpu').manual_seed(4321)
        Xc = torch.randn(Bc, T, G, N, generator=gc).to(dev)
        hc = torch.zeros(Bc, F, N, device=dev)
        dc = torch.ones(Bc, T, F, N, device=dev)
        clo, chi = gg.dist.shard_range(Bc, rank, world)
        grad_check = dict(B=Bc, what='gradient bucket of 64 sequences sharded over the ranks and summed by gated_gcrnns_b200.dist.'
                                     'allreduce_gradients vs the same 64 sequences on rank 0 alone (max-norm relative), at T = 8 and at the '
                                     'full T = 64 (where float-atomic summation order is amplified by the chaotic recurrence: two single-GPU '
                                     'runs differ by the same amount)')
        for Tc in (8, T):
            for p in used:
                p.grad = None
            torch.autograd.backward(cell(Xc[clo:chi, :Tc], hc[clo:chi]), dc[clo:chi, :Tc])
            red = gg.dist.allreduce_gradients(used, op='sum').clone()
            if rank == 0:
                for p in used:
                    p.grad = None
                torch.autograd.backward(cell(Xc[:, :Tc], hc), dc[:, :Tc])
                full = torch.cat([p.grad.reshape(-1) for p in used])
                grad_check[f'rel_err_T{Tc}'] = ((red - full).abs().max() / full.abs().max()).item()

    # ---- roofline of the dominant kernel (tcgen05 shift GEMM), timed live with CUDA events on its stream -------------
    pk = peaks()
    roof = None
    if args.precision != 'fp32':
        P = 2 if args.precision == 'bf16x2' else 1
        g = ggraph.get(cell.S, dev, keep_dense=True)
        g.set_option('gemm_pair', pair)
        R = mb * F
        A = torch.randn(R, P * N, device=dev).to(torch.bfloat16)
        out = torch.empty(R, P * N, dtype=torch.bfloat16, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        reps = 40

        def one():
            _lib.check(L.gcrnn_debug_shift_gemm(g.ptr, 0, C.c_void_p(A.data_ptr()), R, P, C.c_void_p(out.data_ptr()), P, C.c_void_p(0), st), 'gemm')
        for _ in range(5):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            one()
        e1.record()
        torch.cuda.synchronize()
        t_k = e0.elapsed_time(e1) * 1e-3 / reps
        ach = 2.0 * R * N * N / t_k / 1e12               # ALGORITHMIC flops of the shift (one product), whatever the operand split
        pipe = P * ach                                   # flops the tensor pipe executed (P K-concatenated products)
        step_ach = FLOP_PER_SEQ_FWD_BWD * seqs / world / 1e12
        traffic, tsrc, tshape = ncu_traffic_for_shape(R, N, P, pair)
        eq = pk['bf16'] / P
        roof = dict(bound='tensor', achieved=ach, peak=pk['bf16'], unit='TFLOP/s', frac=ach / pk['bf16'], traffic=traffic,
                    traffic_note=tshape,
                    tensor_pipe_achieved=pipe, tensor_pipe_frac=pipe / pk['bf16'],
                    mode_equivalent_peak=eq, frac_of_mode_equivalent_peak=ach / eq,
                    mode_note=(f'{args.precision}: every product runs as {P} K-concatenated bf16 MMAs into one fp32 accumulator; `achieved`/`frac` count the '
                               f'algorithmic 2*M*N^2 flops once against the measured bf16 peak, `tensor_pipe_*` count what the tensor pipe executed, '
                               f'`mode_equivalent_peak` = bf16 peak / {P} (SURVEY.md 8d: a tf32-class mode is held to half the bf16 peak)'),
                    kernel=f'{"shift_gemm2_kernel (cta_group::2, 256x256 pair tiles)" if pair else "shift_gemm_kernel<256> (cta_group::1)"} '
                           f'[{R}x{P}*{N}]x[{N}x{N}] bf16, {t_k * 1e6:.1f} us/launch, peak = {pk["src"]} burst bf16',
                    step_achieved=step_ach, step_peak=pk['bf16_sustained'], step_frac=step_ach / pk['bf16_sustained'],
                    step_frac_of_mode_equivalent_peak=step_ach / (pk['bf16_sustained'] / P),
                    step_note='algorithmic 82.82 GFLOP/sequence fwd+bwd x sequences/s per GPU vs sustained bf16 peak (' + pk['src'] + ')')
    else:
        step_ach = FLOP_PER_SEQ_FWD_BWD * seqs / world / 1e12
        roof = dict(bound='tensor', achieved=step_ach, peak=pk['bf16_sustained'], unit='TFLOP/s', frac=step_ach / pk['bf16_sustained'],
                    traffic=None, kernel='fp32 sparse exact path (no tensor cores); whole-step algorithmic flops')

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times, cores, kind = cpu_sample(dict(CFG3), args.cpu_batch, args.cpu_T, 2)
        cb = cpu_baseline_dict(dict(CFG3), times, cores, args.cpu_batch, args.cpu_T, kind)

    # ---- secondary workloads in the SAME line (N = 1 only): cfg5 (sparse kNN graph, HBM roofline) and cfg1 (the reference's own
    # CPU-runnable configuration), each measured by the code path `--workload cfg5 / cfg1` runs, with fewer steps ------------------
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        import copy
        peak_gb = round(torch.cuda.max_memory_allocated() / 1e9, 1)
        del X_dev, dH, xbuf, hbuf, h0_dev
        cell.zero_grad(set_to_none=True)
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        secondary = {}
        # cfg5 twice: the graph in the generator's RANDOM node order (the library renumbers it itself: graph option 'reorder'), and
        # numbered along a Hilbert curve by the library's graph builder
        for name, wl, st_, wu_, order in (('cfg5', 'cfg5', 10, 2, 'random'), ('cfg5-hilbert', 'cfg5', 5, 2, 'hilbert'), ('cfg1', 'cfg1', 20, 5, None),
                                          ('cfg2-node', 'cfg2-node', 20, 5, None), ('cfg2-edge', 'cfg2-edge', 20, 5, None)):
            a2 = copy.copy(args)
            a2.workload, a2.steps, a2.warmup, a2.no_cpu_baseline, a2.batch = wl, st_, wu_, True, CFG3['B']
            if order:
                a2.cfg5_order, a2.cfg5_reorder = order, 'auto'
            try:
                o = run_cfg5(a2, emit_line=False) if wl == 'cfg5' else run_small(a2, emit_line=False)
                secondary[name] = {k: o[k] for k in ('value', 'unit', 'steps', 'warmup', 'ms_per_step', 'dtype', 'config', 'roofline', 'e2e', 'gpu_launches',
                                                     'whole_step') if k in o}
            except Exception as e:                      # a secondary leg must never take the headline line down
                secondary[name] = dict(error=f'{type(e).__name__}: {e}')
            gc.collect()
            torch.cuda.empty_cache()
        try:
            secondary['cfg3-node'] = cfg3_node_extra(dev, cfg, mb)
        except Exception as e:
            secondary['cfg3-node'] = dict(error=f'{type(e).__name__}: {e}')
        gc.collect()
        torch.cuda.empty_cache()
        gg.set_precision(args.precision)
    else:
        peak_gb = round(torch.cuda.max_memory_allocated() / 1e9, 1)

    if rank == 0:
        out = dict(metric='GCRNN sequences/sec fwd+bwd', value=seqs, unit='sequences/s', n_gpus=world, steps=args.steps,
                   warmup=args.warmup, ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling='strong',
                   vs_baseline=None, dtype=PREC_DTYPE[args.precision], data='synthetic',
                   config=dict(workload='cfg3: dense N=1024 F=64 G=1 K=5 T=64 time-gated GGCRNNCell fwd+bwd'
                                        + (' + NODE gates (extra; step_frac still counts the time-gated cell\'s 82.82 GFLOP/sequence)' if sg else ''),
                               global_batch=cfg['B'], per_gpu_batch=Bl, microbatch=mb, precision=args.precision,
                               parallelism=f'dp{world} (batch sharded, one gradient all-reduce per step'
                                           + (f' through gated_gcrnns_b200.dist.allreduce_gradients, transport {"gcrnn_allreduce_sum (C ABI)" if args.native_allreduce else "torch.distributed nccl"})' if world > 1 else ')'),
                               l2=f'inputs larger than L2 (X 1 GiB, H {mb * T * F * N * 4 / 1e9:.1f} GB per micro-batch); no explicit flush'),
                   roofline=roof, cpu_baseline=cb, clocks=clocks,
                   e2e=dict(value=seqs_e2e, unit='sequences/s', h2d_bytes_per_step=int(X_host.numel() * 4 + (Bl // mb) * h0_host.numel() * 4),
                            d2h_bytes_per_step=int(sum(p.numel() for p in used) * 4), ms_per_step=ms_e2e / args.steps),
                   modes=modes, parity=parity, grad_check=grad_check, secondary=secondary,
                   gpu_launches=int(launches), peak_hbm_gb=peak_gb)
        emit(out)
    if world > 1:
        gg.dist.disable()
        dist.destroy_process_group()


def cfg3_node_extra(dev, cfg, mb, precisions=('bf16x2', 'bf16')):
    """cfg3 with NODE gates on top of the time gates (SURVEY.md 8d extra): device-resident forward + backward of one micro-batch on the
    tensor-core path (csrc/tc_node.cuh), per operand mode; 1 warm-up + 2 timed micro-batches."""
    import torch
    import gated_gcrnns_b200 as gg
    N, F, G, K, T = cfg['N'], cfg['F'], cfg['G'], cfg['K'], cfg['T']
    S = gg.graphs.dense_random(N, cfg['density'], seed=0)
    out = {}
    X = torch.randn(mb, T, G, N, device=dev)
    h0 = torch.zeros(mb, F, N, device=dev)
    dH = torch.ones(mb, T, F, N, device=dev)
    try:
        for prec in precisions:
            gg.set_precision(prec)
            torch.manual_seed(0)
            cell = gg.GGCRNNCell(G, F, K, K, torch.tanh, True, 'node', 1, True)
            cell.addGSO(S)
            cell = cell.to(dev)

            def once():
                cell.zero_grad(set_to_none=True)
                H = cell(X, h0)
                torch.autograd.backward(H, dH)
                del H
            once()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                once()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            out[prec] = dict(value=mb / (ms * 1e-3), unit='sequences/s', ms_per_microbatch=ms)
            del cell
    finally:
        del X, h0, dH
    out['config'] = dict(workload=f'cfg3 + node gates: dense N={N} F={F} G={G} K={K} T={T}, time + node gated GGCRNNCell fwd+bwd, micro-batch {mb}, '
                                  'inputs resident in HBM, tensor-core path')
    return out


def ncu_traffic_for_shape(R, N, P, pair):
    """DRAM bytes per launch of the shift GEMM from a committed `ncu --set full` export captured at EXACTLY this launch shape
    (rows R, planes P); None with the reason when no export of this shape is committed (never another shape's number)."""
    name = 'shift_gemm2_kernel' if pair else 'shift_gemm_kernel'
    fn = f'r02_ncu_gemm2_R{R}_P{P}.raw.csv' if P > 1 or not os.path.isfile(os.path.join(ROOT, 'profiles', f'r01_ncu_gemm2_mb{R // 64}.raw.csv')) \
        else f'r01_ncu_gemm2_mb{R // 64}.raw.csv'
    traffic, src = ncu_traffic([fn], name)
    alg = (2 * R * N * 2 * P + N * N * 2) / 1e6
    if src is None:
        return None, None, f'no committed ncu --set full export for this launch shape ([{R}x{P}*{N}] bf16 in and out); algorithmic operand bytes per launch = {alg:.0f} MB'
    return traffic, src, (f'DRAM read+write bytes per launch from profiles/{src} (ncu --set full, same launch shape [{R}x{P}*{N}]); '
                          f'algorithmic operand bytes per launch = {alg:.0f} MB')


# ------------------------------------------------------------------------------------------------------
# secondary workload: cfg5 (sparse kNN graph, CSR SpMM path, edge-gated) — HBM-bound; `--workload cfg5`
# ------------------------------------------------------------------------------------------------------
CFG5 = dict(N=100_000, knn=16, F=32, G=1, K=3, T=32, B=256, mb=32)
BYTES_PER_SEQ_CFG5 = 42 * 32 * (32 * 100_000 * 4)      # 42 P per (sample, step), P = F*N*4 B, T = 32 -> 17.2 GB


def cfg5_cpu_baseline(cfg, S_csr, Bs=1, Ts=2):
    """The reference algorithm on the host CPUs for cfg5: the dense reference cannot hold N = 1e5 (S alone is 80 GB), so this
    times the sparse fp64 oracle port (oracle/gcrnn_oracle.py, pinned against the reference's golden vectors) on a bounded
    sample: B = 1 sequence, T = 2 of 32 steps, forward + backward, extrapolated linearly in T."""
    import time
    import torch
    from oracle import gcrnn_oracle as orc
    N, F, G, K = cfg['N'], cfg['F'], cfg['G'], cfg['K']
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    prev = torch.get_default_dtype(); torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(0)
        p = orc.init_cell_params(G, F, K, K, N, False, 'edge', 1, True)
    finally:
        torch.set_default_dtype(prev)
    S = S_csr.double().to_sparse_coo().coalesce()
    X, h0, dH = torch.randn(Bs, Ts, G, N).double(), torch.zeros(Bs, F, N).double(), torch.ones(Bs, Ts, F, N).double()
    times = []
    for _ in range(2):
        t0 = time.perf_counter()
        orc.cell_forward_backward(p, [S], X, h0, dH, False, 'edge')
        times.append(time.perf_counter() - t0)
    best = min(times)
    return dict(value=Bs / (best * cfg['T'] / Ts), unit='sequences/s', cores=cores, kind='port',
                sample=f'sparse fp64 oracle port of the reference GGCRNNCell (edge-gated; the dense reference cannot hold N = 1e5), torch CPU '
                       f'with {cores} threads, fwd+bwd on the cfg5 graph with B={Bs}, T={Ts} of {cfg["T"]}; min of 2 runs = {best:.2f} s, '
                       f'extrapolated linearly in T')


def cfg5_dram_bytes_per_sequence(cfg):
    """Measured DRAM bytes per sequence from the committed `ncu --set full` export of one launch of every per-step kernel
    (profiles/r01_ncu_sparse_v2.raw.csv, captured at micro-batch 32): sum over a forward + backward step, x T, / 32 sequences."""
    per_step = [('spmm32_v2_k', 2), ('gather_contract_k<3, 0', 1), ('rowstats_v2_k', 1), ('aggregate_v2_k', 1),
                ('bwd_rows_v2_k', 1), ('bwd_node_v2_k', 1), ('gather_contract_k<3, 1', 1)]
    tot = 0.0
    for name, n in per_step:
        b, _ = ncu_traffic(['r01_ncu_sparse_v2.raw.csv'], name)
        if b is None:
            return None
        tot += n * b
    return tot * cfg['T'] / 32


def run_cfg5(args, emit_line=True):
    import torch
    import torch.distributed as dist
    import gated_gcrnns_b200 as gg
    from gated_gcrnns_b200 import _lib
    world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cfg = dict(CFG5)
    if args.batch != CFG3['B']:
        cfg['B'] = args.batch
    N, F, G, K, T = cfg['N'], cfg['F'], cfg['G'], cfg['K'], cfg['T']
    lo, hi = gg.dist.shard_range(cfg['B'], rank, world)
    Bl = hi - lo
    mb = min(cfg['mb'], Bl)
    assert Bl % mb == 0
    gg.set_precision('fp32')
    # the graph is built by the library's on-device builder (csrc/builders.cu): 'hilbert' = the library renumbers the nodes along a
    # Hilbert curve (locality owned by the library), 'random' = the generator's random node order is kept, 'host' = round 1's scipy path
    order = getattr(args, 'cfg5_order', 'hilbert')
    t_build = time.perf_counter()
    if order == 'host':
        rp, ci, va = gg.graphs.knn_csr(N, cfg['knn'], seed=0)
    else:
        rp, ci, va, _ = gg.graphs.knn_csr_gpu(N, cfg['knn'], seed=0, device=dev, reorder=(order == 'hilbert'))
    t_build = time.perf_counter() - t_build
    S = gg.graphs.csr_to_torch_sparse(rp, ci, va, N)
    torch.manual_seed(0)
    cell = gg.GGCRNNCell(G, F, K, K, torch.tanh, False, 'edge', 1, True)
    cell.addGSO(S)
    cell = cell.to(dev)
    # node renumbering owned by the library (graph option 'reorder', include/gcrnn_b200.h): the same cached handle the cell uses
    gh = gg.graph.get(cell.S, dev, keep_dense=False)
    gh.set_option('reorder', {'off': 0, 'auto': 1, 'force': 2}[args.cfg5_reorder])
    reorder = dict(mode=args.cfg5_reorder, reordered=bool(gh.get_option('reordered')),
                   tile_rows_before=gh.get_option('tile_rows_before_x100') / 100, tile_rows_after=gh.get_option('tile_rows_after_x100') / 100,
                   note='distinct neighbour rows per 128-node tile / 128 in the caller\'s node order and in the library\'s; the library '
                        'renumbers (breadth-first balls of 128 nodes) when that lowers it by >= 1.5x')
    used = [dict(cell.named_parameters())[n] for _, _, n in gg.cell_param_slots(False, 'edge', True)]
    gen = torch.Generator(device='cpu').manual_seed(1234 + rank)
    X_host = torch.randn(Bl, T, G, N, generator=gen).pin_memory()
    X_dev = X_host.to(dev)
    h0 = torch.zeros(mb, F, N, device=dev)
    dH = torch.ones(mb, T, F, N, device=dev)
    L = _lib.lib()

    # e2e leg: the pinned-host -> device copy of micro-batch i+1 runs on a side stream while micro-batch i computes (double buffer),
    # all inside the timed region; the first micro-batch of a step is copied at the start of that step
    copy_stream = torch.cuda.Stream(device=dev)
    xbuf = [torch.empty(mb, T, G, N, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i, slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            xbuf[slot].copy_(X_host[i:i + mb], non_blocking=True)
            ready[slot].record(copy_stream)

    def step(host_inputs):
        for p in used:
            p.grad = None
        starts = list(range(0, Bl, mb))
        if host_inputs:
            prefetch(starts[0], 0)
        for j, i in enumerate(starts):
            if host_inputs:
                slot = j & 1
                if j + 1 < len(starts):
                    prefetch(starts[j + 1], slot ^ 1)
                torch.cuda.current_stream().wait_event(ready[slot])
                x = xbuf[slot]
            else:
                x = X_dev[i:i + mb]
            H = cell(x, h0)
            torch.autograd.backward(H, dH)
            del H
            if host_inputs:
                freed[slot].record(torch.cuda.current_stream())
        bucket = torch.cat([p.grad.reshape(-1) for p in used])
        if world > 1:
            dist.all_reduce(bucket)
        return bucket.to('cpu') if host_inputs else bucket

    def timed(host_inputs, steps, warmup):
        for _ in range(warmup):
            step(host_inputs)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clk = Clocks(local); clk.start()
        l0 = L.gcrnn_debug_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(host_inputs)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), L.gcrnn_debug_launch_count() - l0, clk.result()

    if args.once:
        Bl = mb
        step(False)
        torch.cuda.synchronize()
        return
    ms_dev, launches, clocks = timed(False, args.steps, args.warmup)
    ms_e2e, _, _ = timed(True, args.steps, 1)
    seqs = cfg['B'] * args.steps / (ms_dev * 1e-3)
    seqs_e2e = cfg['B'] * args.steps / (ms_e2e * 1e-3)
    pk = peaks()
    ach = BYTES_PER_SEQ_CFG5 * seqs / world / 1e9
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cfg5_cpu_baseline(cfg, S)
    traffic = cfg5_dram_bytes_per_sequence(cfg) if cfg['K'] == 3 else None
    if rank == 0:
        out = dict(metric='GCRNN sequences/sec fwd+bwd', value=seqs, unit='sequences/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f32', data='synthetic',
                   config=dict(workload='cfg5: sparse directed 16-NN graph N=100000 (CSR SpMM path), F=32 G=1 K=3 T=32 edge-gated GGCRNNCell fwd+bwd',
                               node_order=order, library_reorder=reorder, graph_build_s=round(t_build, 2),
                               global_batch=cfg['B'], per_gpu_batch=Bl, microbatch=mb, precision='fp32',
                               parallelism=f'dp{world} (batch sharded, one gradient all-reduce per step)',
                               l2='per-micro-batch working set (H 13 GB) larger than L2; no explicit flush'),
                   roofline=dict(bound='hbm', achieved=ach, peak=pk['hbm'], unit='GB/s', frac=ach / pk['hbm'], traffic=traffic,
                                 traffic_note='measured DRAM read+write bytes per SEQUENCE (the unit of `achieved`): per-launch bytes of the seven '
                                              'per-step kernels from profiles/r01_ncu_sparse_v2.raw.csv (ncu --set full, micro-batch 32) x T / 32; '
                                              'algorithmic bytes per sequence = 17.2e9',
                                 kernel='whole step: algorithmic 17.2 GB per sequence (42 passes over an [F,N] fp32 signal per (sample, step), '
                                        'SURVEY.md 8d) x sequences/s per GPU, peak = ' + pk['src'] + ' HBM copy bandwidth'),
                   cpu_baseline=cb, clocks=clocks,
                   e2e=dict(value=seqs_e2e, unit='sequences/s', h2d_bytes_per_step=int(X_host.numel() * 4),
                            d2h_bytes_per_step=int(sum(p.numel() for p in used) * 4), ms_per_step=ms_e2e / args.steps),
                   gpu_launches=int(launches))
        if emit_line:
            emit(out)
    if world > 1:
        dist.destroy_process_group()
    return out if rank == 0 else None


# ------------------------------------------------------------------------------------------------------
# small reference configurations (parity + latency cases of SURVEY.md 8d; launch-bound, no roofline claim):
#   cfg1      time-gated cell on the N=80 SBM graph of kStepPredGRNNs.py (F=20, K=5, T=5, B=100)
#   cfg2-node / cfg2-edge   node- / edge-gated cell on the Adj.p seismograph graph (N=59, F=20, K=4, T=20, B=100)
# The graphs come from the committed golden fixtures (generated from the reference), signals are synthetic.
# ------------------------------------------------------------------------------------------------------
SMALL = {'cfg1': ('cell_cfg1_time', True, None, 5, 5), 'cfg2-node': ('cell_cfg2_node', False, 'node', 4, 20),
         'cfg2-edge': ('cell_cfg2_edge', False, 'edge', 4, 20)}


def run_small(args, emit_line=True):
    import numpy as np
    import torch
    import gated_gcrnns_b200 as gg
    from gated_gcrnns_b200 import _lib
    fixture, tg, sg, K, T = SMALL[args.workload]
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    dev = torch.device('cuda', 0)
    S = torch.tensor(np.load(os.path.join(ROOT, 'tests', 'golden', fixture + '.npz'), allow_pickle=True)['S'])
    N, F, G, B = S.shape[-1], 20, 1, 100
    gg.set_precision('fp32')
    torch.manual_seed(0)
    cell = gg.GGCRNNCell(G, F, K, K, torch.tanh, tg, sg, 1, True)
    cell.addGSO(S)
    cell = cell.to(dev)
    X_host = torch.randn(B, T, G, N).pin_memory()
    h0_host = torch.zeros(B, F, N).pin_memory()
    X_dev, h0_dev = X_host.to(dev), h0_host.to(dev)
    dH = torch.ones(B, T, F, N, device=dev)
    L = _lib.lib()

    def step(host):
        cell.zero_grad(set_to_none=True)
        x, h = (X_host.to(dev, non_blocking=True), h0_host.to(dev, non_blocking=True)) if host else (X_dev, h0_dev)
        H = cell(x, h)
        torch.autograd.backward(H, dH)
        g = cell.weight_B.grad
        return g.cpu() if host else g

    def timed(host, steps, warmup):
        for _ in range(warmup):
            step(host)
        torch.cuda.synchronize()
        clk = Clocks(0); clk.start()
        l0 = L.gcrnn_debug_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(host)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), L.gcrnn_debug_launch_count() - l0, clk.result()

    steps = max(args.steps, 20)
    ms, launches, clocks = timed(False, steps, max(args.warmup, 5))
    ms_e, _, _ = timed(True, steps, 3)

    # whole TRAINING step (reordering gather + recurrence + per-node readout + L1 loss + backward + Adam) as ONE CUDA graph
    # (gated_gcrnns_b200.train.GraphedStep) next to the same step launched eagerly
    whole = None
    if not args.no_whole_step:
        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.cell = gg.GGCRNNCell(G, F, K, K, torch.tanh, tg, sg, 1, True)
                self.cell.addGSO(S)
                self.readout = torch.nn.Linear(F, 1)

            def forward(self, x, h):
                return self.readout(self.cell(x, h).transpose(2, 3)).squeeze(-1).unsqueeze(2)
        torch.manual_seed(0)
        net = Net().to(dev)
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)
        Y = torch.randn(B, T, 1, N, device=dev)
        order = list(np.random.RandomState(0).permutation(N))
        oidx = torch.as_tensor(order, device=dev)
        l1 = torch.nn.L1Loss()

        def eager_step():
            opt.zero_grad(set_to_none=True)
            loss = l1(net(X_dev.index_select(-1, oidx), h0_dev), Y)
            loss.backward()
            opt.step()
            return loss

        def time_fn(fn, n):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        ms_eager = time_fn(eager_step, steps)
        gstep = gg.train.GraphedStep(net, l1, opt, X_dev, h0_dev, target=Y, order=order)
        ms_graph = time_fn(lambda: gstep(X_dev, h0_dev, target=Y), steps)
        whole = dict(what='one training step: node-reordering gather + GGCRNNCell + per-node Linear(F,1) readout + L1 loss + backward + Adam',
                     eager_seq_per_s=B / (ms_eager * 1e-3), graphed_seq_per_s=B / (ms_graph * 1e-3), ms_eager=ms_eager, ms_graphed=ms_graph)
    cb = None
    if not args.no_cpu_baseline:
        cfg = dict(N=N, F=F, G=G, K=K, T=T, density=None)
        times, cores, kind = cpu_sample(cfg, B, T, 3, warm=1, S=S.double(), tg=tg, sg=sg)
        cb = cpu_baseline_dict(cfg, times, cores, B, T, kind)
        cb['sample'] = cb['sample'].replace('cfg3 shapes', args.workload + ' shapes (full size, no extrapolation)')
    out = dict(metric='GCRNN sequences/sec fwd+bwd', value=B * steps / (ms * 1e-3), unit='sequences/s', n_gpus=1, steps=steps,
               warmup=max(args.warmup, 5), ms_per_step=ms / steps, higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f32',
               data='synthetic',
               config=dict(workload=f'{args.workload}: N={N} F={F} G={G} K={K} T={T} B={B} time_gating={tg} spatial_gating={sg}, fp32 sparse exact path',
                           l2='working set fits L2 (launch/latency-bound case; no roofline claim, SURVEY.md 8d)'),
               roofline=None, cpu_baseline=cb, clocks=clocks, whole_step=whole,
               e2e=dict(value=B * steps / (ms_e * 1e-3), unit='sequences/s', h2d_bytes_per_step=int(X_host.numel() * 4 + h0_host.numel() * 4),
                        d2h_bytes_per_step=int(cell.weight_B.numel() * 4), ms_per_step=ms_e / steps),
               gpu_launches=int(launches))
    if emit_line:
        emit(out)
    return out


def main():
    # stdout carries exactly one JSON line: NCCL's own version banner (written to fd 1 by the library) goes to stderr
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        json_only_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=CFG3['B'], help='global batch (sequences per step); default = cfg3')
    ap.add_argument('--microbatch', type=int, default=2048)
    ap.add_argument('--precision', default='bf16x2', choices=['bf16x2', 'bf16', 'fp32'],
                    help='bf16x2 (default): split-bf16 tensor-core mode with the tight stated bound; bf16: plain bf16 operands '
                         '(fast, short horizons only); fp32: exact CUDA-core path')
    ap.add_argument('--also', default='bf16,fp32', type=lambda v: [m for m in v.split(',') if m],
                    help='other precisions of the same workload to time briefly for the `modes` object (comma separated; "" = none)')
    ap.add_argument('--no-parity', action='store_true', help='skip the in-run parity measurement')
    ap.add_argument('--cfg3-spatial', default='none', choices=['none', 'node'],
                    help='cfg3 extra: add node gates to the time-gated dense cell (tensor-core path, csrc/tc_node.cuh)')
    ap.add_argument('--cfg5-reorder', default='auto', choices=['auto', 'off', 'force'],
                    help='library-owned node renumbering of the fused sparse path (graph option "reorder")')
    ap.add_argument('--cfg5-order', default='hilbert', choices=['hilbert', 'random', 'host'],
                    help='cfg5 graph: built on the GPU with library-owned Hilbert renumbering (default), on the GPU in the random input order, or by the host generator')
    ap.add_argument('--no-secondary', action='store_true', help='N = 1: skip the cfg5 / cfg1 / cfg2 secondary measurements in the same line')
    ap.add_argument('--no-whole-step', action='store_true', help='small workloads: skip the whole-training-step CUDA-graph leg')
    ap.add_argument('--native-allreduce', type=int, default=0, help='N > 1: 1 = the library\'s own NCCL transport (gcrnn_allreduce_sum)')
    ap.add_argument('--cpu-batch', type=int, default=16)
    ap.add_argument('--cpu-T', type=int, default=8)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--gemm-pair', type=int, default=None, help='A/B switch: 1 = CTA-pair shift GEMM, 0 = single-CTA')
    ap.add_argument('--bwd-fused', type=int, default=None, help='A/B switch: 1 = fused reverse-time step kernel, 0 = separate kernels')
    ap.add_argument('--workload', default='cfg3', choices=['cfg3', 'cfg5', 'cfg1', 'cfg2-node', 'cfg2-edge'],
                    help='cfg3 = the headline dense config; cfg5 = sparse kNN graph; cfg1 / cfg2-* = the small reference configurations')
    ap.add_argument('--opt', action='append', default=[], metavar='NAME=VALUE',
                    help='library tuning switch (gated_gcrnns_b200.options), e.g. sparse_v2=0 for the first-generation sparse kernels')
    ap.add_argument('--once', action='store_true', help='run one micro-batch forward+backward and exit (for ncu captures)')
    args = ap.parse_args()
    if args.opt and args.impl != 'reference':
        import gated_gcrnns_b200 as gg
        for kv in args.opt:
            name, value = kv.split('=')
            gg.options.set(name, int(value))
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'cfg5':
        run_cfg5(args)
    elif args.workload in SMALL:
        run_small(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
