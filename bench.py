"""bench.py -- frame-graph update-op iterations/sec on B200 (BASELINE.json metric), with roofline + CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp16|fp32]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" = ONE pass of the hot path over one frame graph = FactorGraph.update (factor_graph.py:227-307):
4-level correlation lookup for all edges + DynamicUpdateModule (incl. GraphAgg) + weight / mask post-processing +
2 Gauss-Newton iterations of the dense BA.  Workload at N=1: BASELINE.json configs[1] -- 8 keyframes at 480x640
(60x80 grid), neighbourhood radius 3 => 36 edges, r=3 correlation window, half-precision volume and update operator
(what the reference's inference autocasts to), state fed back from step to step.  The volume build (per new edge,
not per iteration) is reported separately as ``edge_add_ms``.  At N>1 each rank owns the out-edges of 8 keyframes of
one 8N-keyframe graph (weak scaling) and the reduced camera system is all-reduced over NCCL once per BA iteration.

JSON keys follow the driver contract: value = whole-job steps/s with the state resident in HBM; e2e = the same step
driven through the public API with the mutable graph state arriving from / returning to pinned HOST buffers every step
(copies of neighbouring steps overlap the compute on separate streams); roofline = the dominant kernel (tensor-core
448->256 GRU gate convolution) vs the measured bf16 peak in MEASURED_PEAKS.json, with its DRAM traffic from the committed ncu
capture (profiles/ncu_traffic.json); cpu_baseline = the CPU oracle (a port of the reference's Python path) on the host
cores; reference_gpu_path = the reference's own GPU code for the stages that can be had on this box, timed on the same GPU
(its correlation kernel and its BA kernels compiled into oracle/_ref/, its update operator's module structure on cuDNN).
--workload vkitti / fpn run BASELINE configs[2] shape / configs[3] (fusion / loop: the video branch of the segmenter, the whole
per-frame loop); --radius 8 at 8 GPUs = BASELINE configs[4] (64 keyframes, banded-8 graph, 952 edges); --impl reference = the
reference's CPU path.  Further keys of the line (round 2): use_inactive_step = the frontend's mode (stored edges in the BA) timed the
same way; roofline_other.corr_lookup / corr_volume = the HBM-bound kernels (the volume build warm, per new edge); parity_gate = the
benchmarked half path's deviation from float32 next to the reference's own autocast run on the same GPU (a checker leg: it runs the
staged reference bytecode of oracle/_ref/py through tests/ref_harness.py, outside every timed region; --no-gate skips it);
multi_gpu_consistency = N > 1 only: pose replicas bitwise equal across ranks and equal to a single-GPU solve of the whole graph,
asserted during warm-up; numa = CPUs this rank bound itself to before allocating its pinned staging buffers.
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

METRIC = 'VO update-op iters/sec (480x640, 8 KF, r=3)'
UNIT = 'iters/s'
F_PER_GPU, RADIUS, HT, WD = 8, 3, 60, 80


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src='fallback')


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith('active'):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


def pin_to_gpu_numa(index):
    """Bind this process to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function) so that the pinned staging
    buffers of the end-to-end leg are first-touched on that NUMA node: 8 NUMA-blind ranks staging ~60 MB each way per step
    made e2e regress from 4 to 8 GPUs in round 1.  Best effort; returns what it did for the bench line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
            NUMA['cpus'] = len(cpus)
    except Exception as ex:
        NUMA['error'] = repr(ex)[:80]


NUMA = {}


# ------------------------------------------------------------------------------------------------ workloads
def make_case(nf, seed=1):
    from tests.golden import cases
    return cases.graph_case(F=nf, radius=RADIUS, h=HT, w=WD, seed=seed)


def update_flops(E, K, hw):
    """MACs*2 of DynamicUpdateModule as executed: per edge and per source frame."""
    per_px_edge = (196 * 128 + 128 * 128 * 9 + 8 * 49 * 128 + 128 * 64 * 9 + 128 * 128 + 3 * 448 * 128 * 9 +
                   5 * 128 * 128 * 9 + 4 * 128 * 2 * 9)
    per_px_frame = 128 * 128 * 9 + 128 * 9
    return 2.0 * hw * (E * per_px_edge + K * per_px_frame)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pvo_b200 import _lib
    from pvo_b200.depth_video import DepthVideo
    from pvo_b200.droid_net import DynamicUpdateModule
    from pvo_b200.factor_graph import FactorGraph

    pin_to_gpu_numa(local_rank)                 # before any pinned allocation: staging buffers become NUMA-local to the GPU
    torch.cuda.set_device(local_rank)
    dev = 'cuda:%d' % local_rank
    nf = F_PER_GPU * world
    c = make_case(nf)
    ii, jj = c['ii'], c['jj']
    if world > 1:                               # shard: this rank owns the out-edges of its 8 keyframes
        own = (ii // F_PER_GPU) == rank
        ii, jj = ii[own], jj[own]
    E = ii.shape[0]
    segm_on = args.workload == 'vkitti'
    v = DepthVideo(image_size=(8 * HT, 8 * WD), buffer=max(nf + 8, 16), device=dev, store_images=False, segm_filter=segm_on, thresh=0.8)
    if segm_on:                                 # synthetic panoptic ids: 4 stuff bands + 8 rectangular things per frame
        g = torch.Generator().manual_seed(5)
        seg = (torch.arange(HT)[:, None] * 4 // HT + 1).repeat(1, WD)[None].repeat(nf, 1, 1) * 10000
        for f in range(nf):
            for k in range(8):
                y, x = int(torch.randint(0, HT - 8, (1,), generator=g)), int(torch.randint(0, WD - 16, (1,), generator=g))
                seg[f, y:y + 8, x:x + 16] = 120000 + k + 1
        v.segms[:nf, 0] = seg.to(v.segms.dtype).to(dev)
    v.poses[:nf] = c['poses'].to(dev)
    v.disps[:nf] = c['disps'].to(dev)
    v.intrinsics[:] = c['intrinsics'][0].to(dev)
    v.fmaps[:nf] = c['fmaps'].to(dev)
    v.nets[:nf] = c['nets'].to(dev)
    v.inps[:nf] = c['inps'].to(dev)
    v.counter.value = nf
    torch.manual_seed(0)                        # random-init weights of the reference architecture (PyTorch default init)
    m = DynamicUpdateModule(precision=args.precision).to(dev)
    fg = FactorGraph(v, m, device=dev, precision=args.precision, use_cuda_graph=not args.no_graph)
    if world > 1:
        from pvo_b200 import dist as pdist
        pdist.attach(fg, rank, world, F_PER_GPU)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- edge add (volume build), timed separately
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    ev0.record()
    fg.add_factors(ii, jj)
    ev1.record()
    sync()
    edge_add_ms = ev0.elapsed_time(ev1)
    pose0, disp0 = v.poses.clone(), v.disps.clone()
    state0 = None

    def reset_state():
        """keep the iteration numerically tame over many timed steps: restart from the initial graph state"""
        v.poses.copy_(pose0)
        v.disps.copy_(disp0)

    def step():
        fg.update(t0=1, t1=nf, itrs=2)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    # the per-step working set (volume 2.2 GB + activations ~0.9 GB) is far larger than L2, so consecutive steps
    # cannot hit in L2; the explicit flush below is belt and braces and sits OUTSIDE the timed events.
    consistency = None
    if world > 1:
        # multi-GPU correctness, asserted in the bench itself: after the first warm-up step every rank must hold the same poses
        # (bitwise: each rank sums the pushed systems in rank order), and they must equal a single-GPU solve of the WHOLE graph
        step()
        sync()
        mine = v.poses[:nf].clone()
        allp = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        same = all(torch.equal(allp[0], q) for q in allp)
        assert same, 'sharded BA: pose replicas differ across ranks'
        single = None
        if rank == 0:
            v1 = DepthVideo(image_size=(8 * HT, 8 * WD), buffer=max(nf + 8, 16), device=dev, store_images=False, segm_filter=False, thresh=0.8)
            for name in ('poses', 'disps', 'fmaps', 'nets', 'inps'):
                getattr(v1, name)[:nf] = c[name].to(dev)
            v1.intrinsics[:] = c['intrinsics'][0].to(dev)
            v1.counter.value = nf
            fg1 = FactorGraph(v1, m, device=dev, precision=args.precision)
            fg1.add_factors(c['ii'], c['jj'])
            fg1.update(t0=1, t1=nf, itrs=2)
            torch.cuda.synchronize()
            single = float((v1.poses[:nf] - mine).abs().max().item())
            del fg1, v1
            torch.cuda.empty_cache()
            assert single < 5e-3, 'sharded BA differs from the single-GPU solve: %g' % single
        consistency = {'replicas_bitwise_equal': same, 'max_abs_pose_diff_vs_single_gpu': single}
        reset_state()
    for _ in range(max(args.warmup, 3)):
        step()
    sync()
    times = []
    clocks = ClockSampler(local_rank)
    clocks.start()
    for k in range(args.steps):
        flush.zero_()
        if k % 10 == 0:
            reset_state()
        sync()
        ev0.record()
        step()
        ev1.record()
        sync()
        times.append(ev0.elapsed_time(ev1))
    clocks.stop_flag = True
    clocks.join(timeout=2)
    total_ms = sum(times)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    # work unit = one config-2 graph's worth of edges (36): the sharded graph has c['E'] edges in total (interior ranks own
    # 48, the two end ranks 42), so the whole job processes c['E'] / 36 units per step
    E_total = int(c['E']) if world > 1 else E
    units_per_step = E_total / 36.0 if world > 1 else 1.0
    value = units_per_step * 1000.0 / ms_per_step

    # ---- the frontend's mode: use_inactive=True (droid_frontend.py:49,62,81,87) with stored edges in the BA, same timing recipe
    inactive = None
    if world == 1 and not args.profile_only:
        far_i = torch.arange(0, nf - 4)
        xi, xj = torch.cat([far_i, far_i + 4]), torch.cat([far_i + 4, far_i])           # the 2 (F - 4) edges at distance 4
        fg.add_factors(xi.to(dev), xj.to(dev))
        fg.rm_factors((fg.ii - fg.jj).abs() == 4, store=True)                          # ... retired into the inactive set
        ts = []
        for k in range(3 + min(args.steps, 10)):
            flush.zero_()
            if k % 10 == 0:
                reset_state()
            sync()
            ev0.record()
            fg.update(t0=1, t1=nf, itrs=2, use_inactive=True)
            ev1.record()
            sync()
            if k >= 3:
                ts.append(ev0.elapsed_time(ev1))
        inactive = {'ms_per_step': round(sum(ts) / len(ts), 4), 'active_edges': int(fg.ii.shape[0]), 'inactive_edges_in_ba': int(fg._cache['inac']['n']),
                    'path': 'CUDA graph replay' if not args.no_graph else 'launch plan', 'ratio_to_plain_step': round(sum(ts) / len(ts) / ms_per_step, 3)}
        reset_state()

    # ---- warm volume build (per NEW edge, not per iteration): all-pairs GEMM on tcgen05 + 4-level pyramid in the epilogue
    volume = None
    if world == 1 and not args.profile_only and args.precision == 'fp16':
        from pvo_b200.modules.corr import build_pyramid_tiled
        g1, g2 = v.fmaps[ii.to(dev)].contiguous(), v.fmaps[jj.to(dev)].contiguous()
        ts = []
        for k in range(5):
            flush.zero_()
            sync()
            ev0.record()
            lv = build_pyramid_tiled(g1, g2)
            ev1.record()
            sync()
            if k >= 2:
                ts.append(ev0.elapsed_time(ev1))
            del lv
        vb = E * HT * WD * sum((HT >> l_) * (WD >> l_) for l_ in range(4)) * 2 + 2 * E * 128 * HT * WD * 2
        vms = sum(ts) / len(ts)
        volume = {'bound': 'hbm', 'kernel': 'corr_volume_tc_kernel<80,tiled> (+ 2 layout kernels)', 'ms': round(vms, 3), 'ms_per_edge': round(vms / E, 4),
                  'achieved': round(vb / vms / 1e6, 1), 'peak': peaks()['hbm'], 'unit': 'GB/s', 'frac': round(vb / vms / 1e6 / peaks()['hbm'], 4),
                  'bytes': vb, 'what': 'warm build of all %d edges: pyramid written once (61 MB/edge) + features read' % E}

    # ---- instrumented pass: per-launch CUDA events on the launching stream
    eng = fg._engine()
    kern = {}
    eng.timers = kern
    fg.timers = kern
    for _ in range(min(args.steps, 10)):
        step()
    torch.cuda.synchronize()
    eng.timers = None
    fg.timers = None
    kernel_ms = {k: sum(a.elapsed_time(b) for a, b in evs) / len(evs) for k, evs in kern.items()}
    launches = getattr(fg, 'launch_count', None)
    if args.profile_only:
        if rank == 0:
            print(json.dumps({'profile_only': True, 'ms_per_step': round(ms_per_step, 4), 'kernel_ms': kernel_ms}))
        return

    # ---- end to end through the public API with HOST state buffers
    K = fg._cache['K']
    def blob(specs, device, pinned=False):
        """one contiguous buffer + named views: a step's state crosses PCIe as ONE copy per direction instead of seven"""
        off, lay = 0, {}
        for n, t_ in specs:
            nb = t_.numel() * t_.element_size()
            lay[n] = (off, nb, t_.dtype, tuple(t_.shape))
            off += (nb + 255) // 256 * 256
        raw = torch.empty(off, dtype=torch.uint8, device=device)
        if pinned:
            raw = raw.pin_memory()
        return raw, {n: raw[o:o + nb].view(dt).view(shp) for n, (o, nb, dt, shp) in lay.items()}

    in_specs = [('poses', v.poses[:nf]), ('disps', v.disps[:nf]), ('target', fg.target_cam), ('weight', fg.weight),
                ('raw_mask', fg.raw_mask), ('delta_dy', fg.delta_dy), ('net', eng.X[:, :128].contiguous())]
    host_raw, host = blob(in_specs, 'cpu', pinned=True)
    host['poses'].copy_(pose0[:nf])
    host['disps'].copy_(disp0[:nf])
    for n, t_ in (('target', fg.target_cam), ('weight', fg.weight), ('raw_mask', fg.raw_mask), ('delta_dy', fg.delta_dy)):
        host[n].copy_(t_)
    host['net'].copy_(eng.X[:, :128])
    out_specs = [('poses', host['poses']), ('disps', host['disps']), ('net', host['net']), ('full_flow', host['target'])]
    out_raw, out_host = blob(out_specs, 'cpu', pinned=True)
    h2d = sum(t_.numel() * t_.element_size() for t_ in host.values())
    d2h = sum(t_.numel() * t_.element_size() for t_ in out_host.values())

    # Three streams: H2D of step k+1 and D2H of step k-1 overlap the compute of step k (each step's inputs come from the
    # host, so consecutive steps are independent).  Host <-> device traffic goes through double-buffered device staging so
    # a copy never touches the live graph state while a step is reading it.
    live_in = {'poses': v.poses[:nf], 'disps': v.disps[:nf], 'target': fg.target_cam, 'weight': fg.weight, 'raw_mask': fg.raw_mask,
               'delta_dy': fg.delta_dy, 'net': eng.X[:, :128]}
    live_out = lambda: {'poses': v.poses[:nf], 'disps': v.disps[:nf], 'full_flow': fg.full_flow, 'net': eng.X[:, :128]}
    stage_in_raw, stage_in, stage_out_raw, stage_out = [], [], [], []
    for _ in range(2):
        r_, v_ = blob(in_specs, dev)
        stage_in_raw.append(r_); stage_in.append(v_)
        r_, v_ = blob(out_specs, dev)
        stage_out_raw.append(r_); stage_out.append(v_)
    s_in, s_out, cur = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
    E_ = lambda: torch.cuda.Event()
    in_ready, in_free, out_ready, out_free = [E_(), E_()], [E_(), E_()], [E_(), E_()], [E_(), E_()]
    for b in range(2):
        in_free[b].record(cur)
        out_free[b].record(cur)

    def e2e_step(k):
        b = k & 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(in_free[b])
            stage_in_raw[b].copy_(host_raw, non_blocking=True)
            in_ready[b].record(s_in)
        cur.wait_event(in_ready[b])
        for n, t_ in live_in.items():
            t_.copy_(stage_in[b][n], non_blocking=True)
        in_free[b].record(cur)
        fg.update(t0=1, t1=nf, itrs=2)
        cur.wait_event(out_free[b])
        for n, t_ in live_out().items():
            stage_out[b][n].copy_(t_, non_blocking=True)
        out_ready[b].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(out_ready[b])
            out_raw.copy_(stage_out_raw[b], non_blocking=True)
            out_free[b].record(s_out)

    for k in range(4):
        e2e_step(k)
    sync()
    n_e2e = max(4, min(args.steps, 20))
    ev0.record()
    for k in range(n_e2e):
        e2e_step(k)
    cur.wait_stream(s_out)                      # the last results are on the host before the clock stops
    ev1.record()
    sync()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = units_per_step * 1000.0 * n_e2e / float(t.item())

    if rank != 0:
        return
    pk = peaks()
    hw = HT * WD
    # dominant kernel: the 448->256 (z|r) 3x3 gate convolution on tensor cores
    zr_flops = 2.0 * E * hw * 448 * 9 * 256
    zr_ms = kernel_ms.get('conv_zr')
    roof = None
    if zr_ms:
        ach = zr_flops / (zr_ms * 1e-3) / 1e12
        roof = {'bound': 'tensor', 'kernel': 'conv_tc_halo_kernel<256,EpZR> (GRU z|r gates, 448->256 3x3)', 'achieved': round(ach, 1),
                'peak': pk['tf_burst'], 'unit': 'TFLOP/s', 'frac': round(ach / pk['tf_burst'], 4), 'traffic': None,
                'peak_source': pk['src'] + ' (burst: kernel timed per launch)', 'ms_per_launch': round(zr_ms, 4),
                'flops_per_launch': zr_flops}
    try:                                        # DRAM traffic of the same kernels from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
    except Exception:
        traffic = {}
    if roof is not None and traffic:
        roof['traffic'] = traffic.get('conv_tc_halo_kernel<256,EpZR>')
        roof['traffic_source'] = traffic.get('source')
    look_ms = kernel_ms.get('corr_lookup')
    secondary = {}
    if look_ms:
        lb = E * hw * (4 * 64 * 2 + 8 + 208 * 2)
        secondary['corr_lookup'] = {'bound': 'hbm', 'achieved': round(lb / (look_ms * 1e-3) / 1e9, 1), 'peak': pk['hbm'], 'unit': 'GB/s',
                                    'frac': round(lb / (look_ms * 1e-3) / 1e9 / pk['hbm'], 4), 'bytes_per_launch': lb,
                                    'traffic': traffic.get('corr_lookup_tiled_kernel'), 'kernel': 'corr_lookup_tiled_kernel<1> (4 levels, tiled volume, cp.async zero-fill staging)'}
    if volume is not None:
        secondary['corr_volume'] = volume
    ba_ms = sum(v_ for k_, v_ in kernel_ms.items() if k_.startswith('ba'))
    upd_ms = sum(v_ for k_, v_ in kernel_ms.items() if k_.startswith('conv') or k_ in ('colmean', 'glo_matvec', 'heads2', 'segment_mean', 'eta'))
    cpu = cpu_baseline_sample()
    gpu_ref = gpu_reference_sample(fg, c, dev) if world == 1 else None
    line = {
        'metric': METRIC, 'value': round(value, 3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': round(ms_per_step, 4), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f16' if args.precision == 'fp16' else 'f32', 'data': 'synthetic (seeded random features / random-init weights)',
        'config': {'workload': ('BASELINE configs[2] shape (VKITTI2 240x808, segment filter on): ' if args.workload == 'vkitti' else 'BASELINE configs[1]: ') + '%d-keyframe factor graph per GPU, r=3 corr, %d edges/GPU, 2 BA iters per update; '
                               'update-op iterations feed back state' % (F_PER_GPU, E),
                   'edges': E_total, 'edges_this_rank': E, 'keyframes': nf, 'grid': [HT, WD], 'precision': args.precision,
                   'l2': 'per-step working set (2.2 GB volume + 0.9 GB activations) >> 126 MB L2; 256 MB flush between timed steps',
                   'parallelism': ('edges sharded by source keyframe; reduced camera system all-reduced per BA iteration by ' + getattr(fg, 'allreduce_kind', 'ncclAllReduce')) if world > 1 else 'single GPU'},
        'e2e': {'value': round(e2e_value, 3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'what': 'every step: mutable graph state (poses, disps, target, weight, masks, hidden state) H2D from pinned host memory, '
                        'FactorGraph.update, results (poses, disps, flow, hidden state) D2H to pinned host memory; copies of '
                        'neighbouring steps overlap the compute on separate streams (double-buffered staging, one packed copy per direction)'},
        'gpu_launches': launches,
        'clocks': clocks.result(),
        'roofline': roof, 'roofline_other': secondary, 'use_inactive_step': inactive, 'multi_gpu_consistency': consistency, 'numa': NUMA,
        'parity_gate': parity_gate(args) if world == 1 else None,
        'kernel_ms': {k: round(x, 4) for k, x in sorted(kernel_ms.items())},
        'stage_ms': {'update_operator': round(upd_ms, 4), 'ba_2_iterations': round(ba_ms, 4), 'corr_lookup': look_ms,
                     'edge_add_ms': round(edge_add_ms, 3)},
        'update_op_tflops': round(update_flops(E, K, hw) / (upd_ms * 1e-3) / 1e12, 1) if upd_ms else None,
        'cpu_baseline': cpu,
        'reference_gpu_path': gpu_ref,
    }
    print(json.dumps(line))


def parity_gate(args):
    """CHECKER leg (not timed, not on the product path; `--no-gate` skips it): how far the benchmarked half / tensor-core path is
    from float32 after two update iterations at config 2, next to how far the REFERENCE's own inference mode is (its unmodified
    FactorGraph + DynamicUpdateModule under torch.autocast on cuDNN, staged bytecode in oracle/_ref/py, same GPU, same inputs).
    tests/test_config2_gpu.py asserts ours <= 2 x reference; the bench line prints both."""
    if args.no_gate or args.precision != 'fp16' or args.workload != 'vo':
        return None
    try:
        from tests import ref_harness
        import torch
        r = ref_harness.fp16_gate(steps=2)
        torch.cuda.empty_cache()
        r['ok'] = all(r['ours_fp16'][k] <= 2.0 * r['reference_autocast'][k] + 1e-3 for k in r['ours_fp16'])
        return {k: ({a: float('%.3g' % b) for a, b in v.items()} if isinstance(v, dict) else v) for k, v in r.items()}
    except Exception as ex:
        return {'unavailable': repr(ex)[:160]}


# ------------------------------------------------------------------------------------------------ CPU arms
def _oracle_step_time(nf, steps=1, warm=0, threads=None):
    import torch
    torch.set_grad_enabled(False)
    if threads:
        torch.set_num_threads(threads)
    from oracle import step as ostep, update_op as oupd
    c = make_case(nf)
    sd = oupd.init_params(0)
    st = ostep.make_state(c)
    ts = []
    for k in range(warm + steps):
        t0 = time.perf_counter()
        ostep.graph_update(sd, st, t0=1, t1=nf, itrs=2)
        if k >= warm:
            ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), c['E']


def best_threads(cores):
    """oneDNN/OpenMP do not always scale to every hardware thread: time ONE step of the FULL config-2 graph (36 edges; a smaller
    graph favours fewer threads than the real workload) at a few counts, keep the best."""
    cands = sorted(set([c for c in (16, 32, 64, cores) if c <= cores] or [cores]))
    best, best_t = cores, None
    for n in cands:
        t, _ = _oracle_step_time(F_PER_GPU, steps=1, warm=0, threads=n)
        if best_t is None or t < best_t:
            best, best_t = n, t
    return best


def cpu_baseline_sample():
    """Bounded sample of the same workload on the host cores (oracle = port of the reference's Python/CPU path)."""
    import torch
    cores = os.cpu_count() or 1
    try:
        nthr = best_threads(cores)
        t, E = _oracle_step_time(F_PER_GPU, steps=1, warm=0, threads=nthr)
        return {'value': round(1.0 / t, 4), 'unit': UNIT, 'cores': nthr, 'kind': 'port', 'host_cores': cores,
                'sample': '1 full update step of the config-2 graph (36 edges, 60x80) on the CPU oracle, %d threads '
                          '(fastest of a small sweep): %.2f s' % (nthr, t)}
    except Exception as ex:       # the baseline must never take the GPU number down with it
        return {'value': None, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': 'failed: %r' % (ex,)}


def gpu_reference_sample(fg, c, dev):
    """The reference's own GPU path for the stages that can be had on this box, timed on the same B200 next to ours
    (SURVEY.md section 8(d) "reference timed beside it"):
      * lookup: the reference's corr_index_forward kernel itself (correlation_kernels.cu compiled by oracle/build_ref.py
        into oracle/_ref/, called the way CorrBlock.__call__ calls it: 4 launches + zero fills + cat), on OUR volume.
      * update operator: the reference's module structure as plain torch convolutions (cuDNN) under half autocast --
        what DynamicUpdateModule.forward executes -- through the restatement in oracle/update_op.py on CUDA tensors.
      * BA: droid_kernels.cu cannot be built (vendored Eigen is incomplete) -> no number.
    Baseline only: nothing here is on the product path."""
    import torch
    out = {'lookup_ms': None, 'update_operator_ms': None, 'ba_ms': None,
           'note': 'reference BA = droid_kernels.cu with its Eigen sparse LLT replaced by a dense float64 Cholesky (the vendored Eigen is '
                   'incomplete; patch documented in oracle/build_ref.py), host round trips as in the reference'}

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    try:
        from oracle import build_ref
        ref = build_ref.load()
        E = c['E']
        coords = (fg.coords0[None, None] + 0.37).repeat(1, E, 1, 1, 1)
        if ref is not None:
            pyr = fg.corr.corr_pyramid

            def ref_lookup():                                  # modules/corr.py:40-50
                cs = coords[0].permute(0, 3, 1, 2).contiguous()
                outs = []
                for i, vol in enumerate(pyr):
                    o, = ref.corr_index_forward(vol, (cs / 2 ** i).contiguous(), 3)
                    outs.append(o.view(1, E, -1, HT, WD))
                return torch.cat(outs, dim=2).contiguous()
            out['lookup_ms'] = round(timed(ref_lookup), 4)
        else:
            out['note'] += '; oracle/_ref/droid_corr_ref.so absent (built only where the reference tree is mounted)'
    except Exception as ex:
        out['note'] += '; reference lookup failed: %r' % (ex,)
    try:
        from oracle import build_ref
        refba = build_ref.load_ba()
        if refba is not None:
            v_ = fg.video
            nf_ = v_.counter.value
            ii_, jj_ = fg.ii, fg.jj
            tgt = fg.target_cam[0].permute(0, 3, 1, 2).contiguous()        # (E,2,h,w) as factor_graph.py:297-300 passes them
            wgt = fg.weight[0].permute(0, 3, 1, 2).contiguous()
            kx_ = torch.unique(torch.cat([torch.arange(1, nf_, device=dev), ii_]))
            eta_ = (0.2 * fg.damping[kx_] + 1e-7).contiguous()
            intr_ = v_.intrinsics[0].contiguous()
            out['ba_ms'] = round(timed(lambda: refba.ba(v_.poses.clone(), v_.disps.clone(), intr_, tgt, wgt, eta_, ii_, jj_, 1, nf_, 2,
                                                         1e-4, 0.1, False)), 4)
        else:
            out['note'] += '; oracle/_ref/droid_ba_ref.so absent'
    except Exception as ex:
        out['note'] += '; reference BA failed: %r' % (ex,)
    try:
        from oracle import update_op as oupd
        sd = {k: v.to(dev).half() for k, v in oupd.init_params(0).items()}
        E = c['E']
        g = torch.Generator(device=dev).manual_seed(0)
        net = torch.tanh(torch.randn(1, E, 128, HT, WD, device=dev, generator=g)).half()
        inp = torch.relu(torch.randn(1, E, 128, HT, WD, device=dev, generator=g)).half()
        corr = torch.randn(1, E, 196, HT, WD, device=dev, generator=g).half()
        flow = torch.randn(1, E, 8, HT, WD, device=dev, generator=g).half()
        ii = c['ii'].to(dev)
        with torch.autocast('cuda', dtype=torch.float16):
            out['update_operator_ms'] = round(timed(lambda: oupd.update_op(sd, net, inp, corr, flow, ii)), 4)
    except Exception as ex:
        out['note'] += '; cuDNN update operator failed: %r' % (ex,)
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU-capable path (its Python stack; droid_backends has no CPU kernels,
    BASELINE.md section 3) as restated by the oracle, all host threads, same config/metric/unit."""
    if rank != 0:
        return
    import torch
    cores = best_threads(os.cpu_count() or 1)
    torch.set_num_threads(cores)
    total = args.steps + args.warmup
    t_full, E = _oracle_step_time(F_PER_GPU, steps=1, warm=0, threads=cores)
    budget = 200.0
    if t_full * total <= budget:
        t, _ = _oracle_step_time(F_PER_GPU, steps=args.steps, warm=args.warmup, threads=cores)
        sample = 'full config-2 step (36 edges) per step'
        scale = 1.0
    else:                       # bounded sample: a 4-keyframe graph (12 edges), cost extrapolated linearly in edges
        t, Es = _oracle_step_time(4, steps=args.steps, warm=args.warmup, threads=cores)
        scale = E / float(Es)
        sample = '4-keyframe sub-graph (%d of %d edges) per step, time scaled by %.2f' % (Es, E, scale)
    value = 1.0 / (t * scale)
    line = {'impl': 'reference', 'metric': METRIC, 'value': round(value, 4), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': round(t * scale * 1e3, 2), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic (seeded random features / random-init weights)',
            'config': {'workload': 'BASELINE configs[1]: 8-keyframe 480x640 factor graph, r=3 corr, 36 edges, 2 BA iters per update',
                       'grid': [HT, WD], 'precision': 'fp32', 'parallelism': 'host cores'},
            'cpu_baseline': {'value': round(value, 4), 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port', 'sample': sample},
            'e2e': {'value': round(value, 4), 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='fp16', choices=['fp16', 'fp32'])
    ap.add_argument('--no-graph', action='store_true', help='launch kernels eagerly instead of replaying a CUDA graph')
    ap.add_argument('--profile-only', action='store_true', help='skip the e2e and CPU legs (for ncu runs)')
    ap.add_argument('--no-gate', action='store_true', help='skip the fp16 parity-gate leg (reference autocast run, oracle/_ref/py)')
    ap.add_argument('--radius', type=int, default=None, help='neighbourhood radius of the frame graph (default 3; 8 = BASELINE configs[4] banded-8 graph)')
    ap.add_argument('--workload', default='vo', choices=['vo', 'vkitti', 'fpn', 'fusion', 'loop'],
                    help='vo = BASELINE configs[1] (the contract line); vkitti = configs[2] shape: 240x808 frames (30x101 grid), 12 keyframes, '
                         'radius 2, panoptic segment filter ON; fpn = configs[3] (delegates to tools/bench_fpn.py); fusion = Panoptic-FPN video branch '
                         '(tools/bench_fusion.py); loop = the whole per-frame loop at the configs[2] shape (tools/bench_loop.py)')
    args, rest = ap.parse_known_args()
    if args.workload == 'fpn':
        sys.argv = [os.path.join(ROOT, 'tools', 'bench_fpn.py'), '--steps', str(min(args.steps, 10)), '--warmup', str(args.warmup)] + rest
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import bench_fpn
        bench_fpn.main()
        return
    if args.workload in ('fusion', 'loop'):
        name = 'bench_' + args.workload
        sys.argv = [os.path.join(ROOT, 'tools', name + '.py')] + rest
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        __import__(name).main()
        return
    global F_PER_GPU, RADIUS, HT, WD, METRIC
    if args.radius:
        RADIUS = args.radius
    if args.workload == 'vkitti':
        F_PER_GPU, RADIUS, HT, WD = 12, 2, 30, 101
        METRIC = 'VO update-op iters/sec (240x808 VKITTI2 shape, 12 KF, radius 2, segment filter on)'
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    import torch
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda:%d' % local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
