#!/usr/bin/env python
"""bench.py -- AAConv2d fwd+bwd throughput on B200 (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--shape T1|T2|T3|T1_512]
    python bench.py --impl reference ...      # the reference algorithm's CPU port on the host cores

A step = one forward + backward of the AAConv2d module at the Transition-1 shape (B=16 per GPU, 256 ch,
80x80 -> 40x40, 8 heads, dk=160, dv=8) on synthetic relu(randn) input, through the
public nn.Module / autograd.Function (which calls the C ABI).  Under torchrun (N>1) every rank runs its own
batch (weak scaling) and the parameter gradients are all-reduced over NCCL inside the step.

Prints ONE JSON line on rank 0.  `value` = algorithmic TFLOP/s (fwd+bwd = 3 x fwd), inputs
resident in HBM; `e2e` = same with pinned-host x/dy copied in and y + parameter grads copied out each step.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPES = {  # name -> (Cin, Hin, Cout, dk, dv)   
    'T1': (256, 80, 128, 160, 8),
    'T2': (512, 40, 256, 160, 24),
    'T3': (1024, 20, 512, 160, 48),
    'T1_512': (256, 128, 128, 160, 8),
}


def flops_fwd(B, cin, hin, cout, dk, dv, nh=8, ks=3):
    H = W = hin // 2
    L = H * W
    dkh, dvh = dk // nh, dv // nh
    dense = 2 * B * L * (ks * ks * cin * (cout - dv) + cin * (2 * dk + dv) + dv * dv)
    attn = 2 * B * nh * L * (L * dkh + dkh * ((2 * W - 1) + (2 * H - 1)) + L * dvh)
    return dense + attn


# dram bytes per launch of the dominant kernels at T1/B=16, copied from the ncu --set full captures under profiles/
NCU_DRAM_BYTES_PER_LAUNCH = {   # profiles/r01_e_final.md (MB rd + MB wr columns)
    'attn_fwd_cc': 111.2e6, 'attn_bwd_dkv_cc': 116.8e6, 'attn_bwd_dq_cc': 154.7e6, 'aug_build_fwd': 80.0e6, 'rel_bwd': 103.1e6,
    'conv_qkv_fprop_tc': 60.9e6, 'conv_qkv_dgrad_tc': 83.4e6, 'conv_qkv_wgrad_tc': 82.4e6, 'pack_nhwc_bf16': 130.0e6,
}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        j = json.load(open(p))
        return j['bf16_tflops'], j['hbm_gbs'], 'measured'
    return 1590.0, 6650.0, 'fallback'


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap',
                     nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: 'hw_power_brake'}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.05)
        except Exception as e:  # NVML missing: report nothing rather than fail the bench
            self.reasons.add(f'nvml_unavailable:{type(e).__name__}')

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def cpu_port(shape, B, steps, warmup, threads):
    """The reference algorithm (oracle port of attn_aug_conv.py:65-97 + autograd) on the host cores."""
    from oracle import aaconv_oracle as O
    cin, hin, cout, dk, dv = SHAPES[shape]
    s = O.AAConvShape(cin, cout, 3, 2, dk, dv, 8, True, (hin // 2, hin // 2))
    torch.set_num_threads(threads)
    m = O.SequentialAAConv2d(s, O.init_params(s, seed=0))
    g = torch.Generator().manual_seed(0)
    x = torch.relu(torch.randn(B, cin, hin, hin, generator=g)).requires_grad_(True)
    dy = torch.randn(B, cout, hin // 2, hin // 2, generator=torch.Generator().manual_seed(1))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        m.zero_grad(set_to_none=True)
        x.grad = None
        m(x).backward(dy)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cin, hin, cout, dk, dv = SHAPES[args.shape]
    # bounded sample: per-step batch chosen so the run stays within a few minutes
    B = args.ref_batch
    sec = cpu_port(args.shape, B, args.steps, min(args.warmup, 1), threads)
    tf = 3 * flops_fwd(B, cin, hin, cout, dk, dv) / sec / 1e12
    line = {'impl': 'reference', 'metric': 'aaconv_fwd_bwd_tflops', 'value': tf, 'unit': 'TFLOP/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': min(args.warmup, 1), 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'AAConv2d {args.shape} fwd+bwd (configs[1])', 'shape': args.shape, 'batch_per_gpu': args.batch,
                       'cin': cin, 'hin': hin, 'cout': cout, 'dk': dk, 'dv': dv, 'nh': 8, 'precision': 'fp32 (reference arithmetic)',
                       'sample_batch_per_step': B},
            'cpu_baseline': {'value': tf, 'unit': 'TFLOP/s', 'cores': threads, 'kind': 'port',
                             'sample': f'oracle.SequentialAAConv2d (reference op order, torch CPU fp32) {args.shape} B={B} '
                                       f'fwd+bwd x{args.steps}'},
            'e2e': {'value': tf, 'unit': 'TFLOP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def run_model(args):
    """configs[2]: whole-model data-parallel training step of aadensenet121 (chexpert.py:152-165) on synthetic radiographs."""
    import torch.distributed as dist
    from chexpert_b200 import _lib
    from chexpert_b200.train import TrainStep, synthetic_batch
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if args.cuda_graph:
            os.environ.setdefault('TORCH_NCCL_ASYNC_ERROR_HANDLING', '0')   # no watchdog aborts while collectives are being captured
        dist.init_process_group('nccl', device_id=dev)
    W, K, B = max(args.warmup, 3), args.steps, args.batch
    ts = TrainStep(dev, size=args.size, precision=args.precision, cuda_graph=args.cuda_graph, channels_last=args.channels_last)
    xh, th = synthetic_batch(B, size=args.size, seed=1000 + rank)
    xh, th = xh.pin_memory(), th.pin_memory()
    x, t = xh.to(dev), th.to(dev)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    def e2e_step():
        xd = xh.to(dev, non_blocking=True)
        td = th.to(dev, non_blocking=True)
        loss_host.copy_(ts(xd, td), non_blocking=True)

    for _ in range(W):
        ts(x, t)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    ms = timed(lambda: ts(x, t), K)
    launches = _lib.launch_count() - l0
    if args.cuda_graph:          # replays do not pass through the host-side launch counter: count what one replay holds
        launches = ts.graph_launches * K
    barrier()
    for _ in range(2):
        e2e_step()
    barrier()
    ms_e2e = timed(e2e_step, max(3, K // 2))
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = tt.tolist()
    if rank == 0:
        line = {'metric': 'aadensenet121_train_images_per_s', 'value': B * world / (ms * 1e-3), 'unit': 'images/s', 'n_gpus': world,
                'steps': K, 'warmup': W, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
                'config': {'workload': 'aadensenet121 training step (configs[2])', 'batch_per_gpu': B, 'image': args.size,
                           'precision': args.precision, 'optimizer': 'SGD nesterov momentum 0.9', 'parallelism': f'dp{world}',
                           'cuda_graph': bool(args.cuda_graph),
                           'l2': 'activations of one step (>> 126 MB) stream through L2; no explicit flush',
                           'dense_blocks': 'torchvision _DenseBlock under torch.autocast(bf16), as the reference wires them'},
                'e2e': {'value': B * world / (ms_e2e * 1e-3), 'unit': 'images/s', 'ms_per_step': ms_e2e,
                        'h2d_bytes_per_step': xh.numel() * 4 + th.numel() * 4, 'd2h_bytes_per_step': 4},
                'gpu_launches': launches, 'clocks': sampler.summary(), 'roofline': None}
        print(json.dumps(line), flush=True)
    if world > 1:
        ts.release()
        barrier()
        if args.cuda_graph:          # process-group teardown after captured NCCL work has been seen to hang: leave without it
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('AACONV_BENCH_PRECISION', 'bf16'), choices=['bf16', 'fp32'])
    ap.add_argument('--shape', default='T1', choices=list(SHAPES))
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--ref-batch', type=int, default=4, help='per-step batch of the CPU reference arm (bounded sample)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-flush', action='store_true')
    ap.add_argument('--workload', default='layer', choices=['layer', 'model'],
                    help="layer: AAConv2d fwd+bwd microbench (configs[1], the headline); model: aadensenet121 training step "
                         "(configs[2]: batch 16/GPU, 320x320, SGD-nesterov, gradient all-reduce), images/s")
    ap.add_argument('--size', type=int, default=320)
    ap.add_argument('--channels-last', action='store_true', help='model workload: channels_last memory format for the dense blocks')
    ap.add_argument('--cuda-graph', action='store_true', help='model workload: capture the whole training step in a CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.workload == 'model':
        return run_model(args)

    import torch.distributed as dist
    import chexpert_b200 as cb
    from chexpert_b200 import _lib

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    cin, hin, cout, dk, dv = SHAPES[args.shape]
    B = args.batch
    H = hin // 2

    torch.manual_seed(0)
    m = cb.AAConv2d(cin, cout, 3, 2, dk, dv, 8, True, (H, H), precision=args.precision)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Conv2d):
            torch.nn.init.kaiming_normal_(mod.weight)
    m = m.to(dev)
    params = [p for p in m.parameters()]
    g = torch.Generator().manual_seed(1000 + rank)
    x_host = torch.relu(torch.randn(B, cin, hin, hin, generator=g)).pin_memory()
    dy_host = torch.randn(B, cout, H, H, generator=g).pin_memory()
    x = x_host.to(dev).requires_grad_(True)
    dy = dy_host.to(dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(xin, dyin):
        for p in params:
            p.grad = None
        xin.grad = None
        y = m(xin)
        y.backward(dyin)
        if world > 1:   # data-parallel exchange step: one bucket with every parameter gradient of the module
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            flat.div_(world)
        return y

    def timed(fn, n):
        """n calls, each bracketed by its own CUDA events on the current stream; L2 flushed in between."""
        evs = []
        for _ in range(n):
            if not args.no_flush:
                flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / n   # ms

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step(x, dy)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    t_wall = time.perf_counter()
    ms = timed(lambda: step(x, dy), K)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = _lib.launch_count() - l0

    # end-to-end: pinned host buffers in, y + parameter grads out, every step
    y_host = torch.empty(B, cout, H, H).pin_memory()
    g_host = [torch.empty(p.shape).pin_memory() for p in params]

    # The way a training loop feeds a GPU: pinned host buffers, the uploads of step k+1 on a copy stream while step k
    # computes, results read back on a third stream from double-buffered device staging.  Every step's H2D and D2H copies
    # are inside the timed region; they overlap compute (a 118 MB upload takes 1.9 ms, the step 0.8 ms).
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    xbuf = [torch.empty_like(x) for _ in range(2)]
    dybuf = [torch.empty_like(dy) for _ in range(2)]
    ystage = [torch.empty(B, cout, H, H, device=dev) for _ in range(2)]
    gstage = [[torch.empty_like(p) for p in params] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]       # upload into input buffer i finished
    freed = [torch.cuda.Event() for _ in range(2)]       # compute on input buffer i finished
    staged = [torch.cuda.Event() for _ in range(2)]      # results of a step are in staging buffer i
    drained = [torch.cuda.Event() for _ in range(2)]     # staging buffer i has been read back
    state = {'k': 0, 'n': 0}

    def upload(i):
        with torch.cuda.stream(up):
            up.wait_event(freed[i])
            xbuf[i].copy_(x_host, non_blocking=True)
            dybuf[i].copy_(dy_host, non_blocking=True)
            ready[i].record(up)

    def e2e_step():
        cur = torch.cuda.current_stream(dev)
        k = state['k']
        i = k & 1
        if k == 0:
            upload(i)
        if k + 1 < state['n']:
            upload(i ^ 1)                                # next step's inputs travel while this step computes
        cur.wait_event(ready[i])
        y = step(xbuf[i].detach().requires_grad_(True), dybuf[i])
        freed[i].record(cur)
        cur.wait_event(drained[i])
        ystage[i].copy_(y.detach())
        for gs, p in zip(gstage[i], params):
            gs.copy_(p.grad)
        staged[i].record(cur)
        with torch.cuda.stream(down):
            down.wait_event(staged[i])
            y_host.copy_(ystage[i], non_blocking=True)
            for gh, gs in zip(g_host, gstage[i]):
                gh.copy_(gs, non_blocking=True)
            drained[i].record(down)
        state['k'] = k + 1

    def e2e_timed(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        state['k'], state['n'] = 0, n
        torch.cuda.synchronize()
        cur = torch.cuda.current_stream(dev)
        for ev in freed + drained:
            ev.record(cur)
        a.record()
        for _ in range(n):
            e2e_step()
        cur.wait_stream(down)                            # the last read-back is part of the region
        cur.wait_stream(up)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    e2e_timed(3)
    barrier()
    ms_e2e = e2e_timed(max(6, K // 2))
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()

    # per-kernel device times of one extra forward+backward (dominant-kernel roofline)
    barrier()
    if not args.no_flush:
        flush_buf.zero_()
    torch.cuda.synchronize()
    _lib.profile_begin(torch.cuda.current_stream().cuda_stream)
    prof = []
    for _ in range(3):
        step(x, dy)
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    agg = {}
    for name, t in prof:
        a = agg.setdefault(name, [0.0, 0])
        a[0] += t
        a[1] += 1
    kern = {n: {'ms_total_per_step': v[0] / 3, 'launches_per_step': v[1] / 3} for n, v in agg.items()}

    if rank == 0:
        f_fwd = flops_fwd(B, cin, hin, cout, dk, dv)
        f_tot = 3 * f_fwd
        peak_tf, peak_gbs, how = peaks()
        tf = f_tot * world / (ms * 1e-3) / 1e12
        top = max(kern.items(), key=lambda kv: kv[1]['ms_total_per_step']) if kern else (None, None)
        roof = None
        if top[0]:
            # algorithmic flops of the dominant kernel per launch (DESIGN.md "work model")
            L = H * H
            nh, dkh, dvh = 8, dk // 8, dv // 8
            attn_f = 2 * B * nh * L * (L * dkh + dkh * (4 * H - 2) + L * dvh)
            dense = 2 * B * L * (9 * cin * (cout - dv) + cin * (2 * dk + dv))
            # algorithmic flops per launch; recompute is not counted
            work = {'attn_fwd': attn_f, 'attn_bwd_dq': attn_f, 'attn_bwd_dkv': attn_f,
                    'conv_qkv_fprop': dense, 'conv_qkv_dgrad': dense, 'conv_qkv_wgrad': dense,
                    'conv_fwd': 2 * B * L * 9 * cin * (cout - dv), 'conv_bwd_data': 2 * B * L * 9 * cin * (cout - dv),
                    'conv_bwd_weight': 2 * B * L * 9 * cin * (cout - dv),
                    'qkv_fwd': 2 * B * L * cin * (2 * dk + dv), 'qkv_bwd_data': 2 * B * L * cin * (2 * dk + dv),
                    'qkv_bwd_weight': 2 * B * L * cin * (2 * dk + dv)}
            key = next((k for k in sorted(work, key=len, reverse=True) if top[0].startswith(k)), None)
            dur = top[1]['ms_total_per_step'] * 1e-3
            ach = work[key] / dur / 1e12 if key else None
            nl = max(top[1]['launches_per_step'], 1.0)
            # every score is exponentiated once per attention kernel: the MUFU.EX2 floor (15.9 ex2/clk/SM measured,
            # tools/mufu_bench.cu) is the binding roofline of the attention kernels, reported beside the tensor one
            exps = B * nh * L * L if top[0].startswith('attn_') else 0
            clk = (sampler.summary()['sm_mhz'] or 1965) * 1e6
            roof = {'kernel': top[0], 'bound': 'tensor', 'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s',
                    'frac': (ach / peak_tf) if ach else None,
                    'traffic': NCU_DRAM_BYTES_PER_LAUNCH.get(top[0]) if args.shape == 'T1' and B == 16 else None,
                    'traffic_source': 'ncu --set full dram__bytes_read.sum + dram__bytes_write.sum, profiles/ (T1, B=16)',
                    'peak_source': how, 'ms_per_step': top[1]['ms_total_per_step'],
                    'launches_per_step': top[1]['launches_per_step'],
                    'us_per_launch': top[1]['ms_total_per_step'] * 1e3 / nl,
                    'mufu_exp_per_launch': exps,
                    'mufu_floor_us': exps / (148 * 15.9 * clk) * 1e6 if exps else None,
                    'mufu_frac': (exps / (148 * 15.9 * clk)) / (top[1]['ms_total_per_step'] * 1e-3 / nl) if exps else None,
                    'module_frac_of_bf16_peak': tf / world / peak_tf}
        h2d = x_host.numel() * 4 + dy_host.numel() * 4
        d2h = y_host.numel() * 4 + sum(t.numel() * 4 for t in g_host)
        line = {'metric': 'aaconv_fwd_bwd_tflops', 'value': tf, 'unit': 'TFLOP/s', 'n_gpus': world, 'steps': K, 'warmup': W,
                'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
                'config': {'workload': f'AAConv2d {args.shape} fwd+bwd (configs[1])', 'shape': args.shape,
                           'batch_per_gpu': B, 'cin': cin, 'hin': hin, 'cout': cout, 'dk': dk, 'dv': dv, 'nh': 8,
                           'precision': args.precision, 'l2': 'flushed between timed steps (256 MiB memset)' if not args.no_flush else 'not flushed',
                           'gflop_per_step_per_gpu': f_tot / 1e9, 'wall_s_timed_region': t_wall},
                'e2e': {'value': f_tot * world / (ms_e2e * 1e-3) / 1e12, 'unit': 'TFLOP/s', 'ms_per_step': ms_e2e,
                        'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'pipeline': 'pinned host buffers; upload of step k+1 and read-back of step k on side streams overlap compute'},
                'gpu_launches': launches, 'clocks': sampler.summary(), 'roofline': roof, 'kernels': kern}
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            cb_B = args.ref_batch
            sec = cpu_port(args.shape, cb_B, 2, 1, threads)
            line['cpu_baseline'] = {'value': 3 * flops_fwd(cb_B, cin, hin, cout, dk, dv) / sec / 1e12, 'unit': 'TFLOP/s',
                                    'cores': threads, 'kind': 'port',
                                    'sample': f'oracle.SequentialAAConv2d {args.shape} B={cb_B} fwd+bwd, 1 warm-up + 2 timed'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
