#!/usr/bin/env python
"""Benchmark of the LAS hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path, one JSON line on rank 0
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on host cores

Headline workload (`config.workload`): C4 of BASELINE.json -- LAS data-parallel TRAINING step, batch 256 per GPU,
T=512 frames of 80-dim fbanks, 40 target characters, conf/default.yaml model (S=256, mlp 128, tf_rate 0.9),
Adadelta(lr=1, eps=1e-8) + clip 5 as in trainer.py:131-148,401-403.  One "step" = zero_grad, forward, loss,
backward, (gradient all-reduce), clip, optimiser step on one synthetic batch.  `value` times it with the batch
resident in HBM; `e2e` times the same step through the drop-in module API with the batch in pinned HOST memory
(H2D copy of x,y and D2H read of the loss and attention maps inside the timed region).
Greedy decoding (C3) and fbank extraction (C2) are timed after it and reported under `extra`.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = dict(output_dim=50, encoder_state_size=256, decoder_state_size=256, mlp_out_size=128, feature_dim=80)
C4 = dict(B=256, T=512, F=80, U=40)
CPU_SAMPLE = dict(B=16, T=512, F=80, U=40)


# ----------------------------------------------------------------------------------------------------------
def synth_batch(B, T, F, U, seed=1234, n_tokens=50):
    """SURVEY.md §8d common synthetic recipe (same generator the tests use)."""
    g = torch.Generator('cpu').manual_seed(seed)
    lens = torch.randint(3 * T // 4, T + 1, (B,), generator=g)
    lens, _ = torch.sort(lens, descending=True)
    lens[0] = T
    x = torch.randn(B, T, F, generator=g)
    x = x * (torch.arange(T)[None, :, None] < lens[:, None, None]).to(x.dtype)
    ylen = torch.randint(max(1, U // 2), U + 1, (B,), generator=g)
    y = torch.zeros(B, U + 2, dtype=torch.long)
    tok = torch.randint(3, n_tokens, (B, U + 2), generator=g)
    for i in range(B):
        n = int(ylen[i])
        y[i, 1:1 + n] = tok[i, 1:1 + n]
        y[i, 1 + n] = 1
    if int(ylen.max()) < U:
        y[0, 1:1 + U] = tok[0, 1:1 + U]
        y[0, 1 + U] = 1
    return x, [int(v) for v in lens], y


def train_flops(B, T, F, U, S=256, Sd=256, M=128, C=50):
    """Algorithmic FLOPs of one training step per kernel family (multiply-add = 2)."""
    rows = [B * T, B * (T // 2), B * (T // 4), B * (T // 8)]
    Ks = [F, 4 * S, 4 * S, 4 * S]
    gemm = 0.0
    for i, (r, k) in enumerate(zip(rows, Ks)):
        gemm += 2.0 * r * 8 * S * k          # input projection
        gemm += 2.0 * r * 8 * S * k          # dW_ih
        if i > 0:
            gemm += 2.0 * r * 8 * S * k      # dX
        gemm += 2 * 2.0 * r * 4 * S * S      # dW_hh, both directions
    rec = sum(r * 16.0 * S * S for r in rows)  # 2 dirs x 2*S*4S per row
    Tp, E = T // 8, 2 * S
    X1, X2 = 2 * Sd + E, 2 * Sd
    sp = 2.0 * B * Tp * E * M + U * B * (2.0 * 4 * Sd * (X1 + X2)) + 2.0 * B * U * Sd * C
    att = U * B * (2.0 * Sd * M + 2.0 * Tp * M + 2.0 * Tp * E)
    gemm += 3 * sp
    return {'gemm': gemm, 'rec_fwd': rec, 'rec_bwd': rec, 'attn_fwd': att, 'attn_bwd': 2 * att,
            'total': gemm + 2 * rec + 3 * att}


def train_bytes(B, T, F, U, S=256):
    """Algorithmic HBM bytes of one training step for the streaming part of the recurrent kernels (per row of a layer:
    forward = pre-activations in (layer 1: the bf16 features, its projection is fused in), activations / h / c fp32 + bf16 h
    out; backward = activations, dhout, c in, bf16 dG out)."""
    rows_l = [B * T, B * (T // 2), B * (T // 4), B * (T // 8)]
    out_b = 8 * S * 4 + 2 * S * 4 + 2 * S * 4 + 2 * S * 2
    fwd = rows_l[0] * (((F + 7) // 8 * 8) * 2 + out_b) + sum(r * (8 * S * 4 + out_b) for r in rows_l[1:])
    bwd = sum(rows_l) * (8 * S * 4 + 2 * S * 4 + 2 * S * 4 + 8 * S * 2)
    return {'rec_fwd_tc': float(fwd), 'rec_bwd_tc': float(bwd)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.f = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def lines(self):
        try:
            with open(self.path) as f:
                return sum(1 for _ in f)
        except OSError:
            return 0

    def wait_ready(self, timeout=2.0):
        """Blocks until nvidia-smi has written its first sample: its start-up (NVML initialisation over every GPU of the box,
        hundreds of ms in the first process on a fresh box) stalls kernel submission and must not fall into the timed region."""
        if self.proc is None:
            return
        t0 = time.time()
        while self.lines() == 0 and time.time() - t0 < timeout and self.proc.poll() is None:
            time.sleep(0.02)

    def stop(self, first_line=0):
        """first_line: number of samples already in the file when the timed region started (they are not reported)."""
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        all_lines = open(self.path).read().splitlines()
        if len(all_lines) >= first_line > 0:      # keep the last sample of the warm-up (under load) for very short timed regions
            all_lines = all_lines[first_line - 1:]
        for line in all_lines:
            p = [v.strip() for v in line.split(',')]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU PyTorch path (oracle port), host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_train_utt_per_s(steps, warmup, sample=CPU_SAMPLE):
    from oracle import las_oracle as O
    from oracle import las_port as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.make_state_dict(seed=1, **DIMS)
    port = P.Port(sd, tf_rate=0.9)
    optim = torch.optim.Adadelta(port.parameters(), lr=1.0, eps=1e-8)
    x, lens, y = synth_batch(sample['B'], sample['T'], sample['F'], sample['U'])
    rng = random.Random(1)
    for _ in range(warmup):
        port.train_step(x, lens, y, optim, rng=rng)
    t0 = time.perf_counter()
    for _ in range(steps):
        port.train_step(x, lens, y, optim, rng=rng)
    dt = time.perf_counter() - t0
    return sample['B'] * steps / dt, dt / steps, cores


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    v, spstep, cores = cpu_train_utt_per_s(args.steps, args.warmup)
    sample = ('%d-utterance batch of the C4 recipe (T=512,F=80,U=40, tf_rate 0.9) per step on the host CPU, torch %s, '
              '%d threads; the reference is pure Python/torch and is timed through oracle/las_port.py (same torch '
              'calls as src/asr.py + trainer.py:415-438)' % (CPU_SAMPLE['B'], torch.__version__, cores))
    line = {'impl': 'reference', 'metric': 'asr_train_utt_per_s', 'value': v, 'unit': 'utt/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': spstep * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.gpus),
            'cpu_baseline': {'value': v, 'unit': 'utt/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': 'utt/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def workload_config(n):
    return {'workload': 'C4 LAS data-parallel training step (BASELINE.json configs[3]): batch 256/GPU, T=512, F=80, '
                        'U=40, S_enc=S_dec=256, mlp=128, tf_rate=0.9, Adadelta+clip5',
            'global_batch': C4['B'] * n, 'per_gpu_batch': C4['B'], 'frames': C4['T'], 'feature_dim': C4['F'],
            'decode_steps': C4['U'] + 1, 'parallelism': 'dp%d' % n,
            'l2': 'working set >> L2 (layer-1 gate buffer alone is 1.07 GB per step); no explicit flush needed'}


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from ss_asr_b200 import _lib
    from ss_asr_b200.asr import ASR
    from ss_asr_b200.functional import asr_loss
    from ss_asr_b200 import preprocess as PP
    from ss_asr_b200.parallel import GradSync

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the '
                           'CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION prints it to stdout) out of it
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    cfg = dict(C4)
    if args.small:
        cfg.update(B=32, T=128)
    B, T, F, U = cfg['B'], cfg['T'], cfg['F'], cfg['U']

    torch.manual_seed(1)
    random.seed(1)
    model = ASR(tf_rate=0.9, **DIMS).to(dev)
    model.train_precision = args.precision
    model.train()
    from ss_asr_b200.optim import FusedAdadelta
    optim = FusedAdadelta(model.parameters(), lr=1.0, eps=1e-8)     # torch.optim.Adadelta + Solver.step fused on the device
    sync = GradSync(model, world)
    from ss_asr_b200 import functional as Fk
    Fk.set_overlap_wgrad(True)     # encoder weight-gradient GEMMs on a second stream under the next layer's recurrent kernel
    x, lens, y = synth_batch(B, T, F, U, seed=1234 + rank)
    ans_len = int(max((y != 0).sum(-1) + 1)) - 1
    x_host, y_host = x.pin_memory(), y.pin_memory()
    x_dev, y_dev = x.to(dev), y.to(dev)

    def step(xd, yd, att_on_device):
        model.att_on_device = att_on_device
        optim.zero_grad(set_to_none=True)
        _, logits, att = model(xd, ans_len, teacher=yd, state_len=lens)
        loss = asr_loss(logits, yd)
        sync.backward(loss)
        optim.step_clipped(5.0)            # trainer.py:144-148: clip_grad_norm_(5) + NaN-skip + Adadelta step, on the device
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()            # started (and waited for) BEFORE the warm-up: it samples every 200 ms from here on
        sampler.wait_ready()
    for _ in range(max(args.warmup, 3)):
        step(x_dev, y_dev, True)
    first_sample = sampler.lines() if rank == 0 else 0
    lib.ssasr_launch_count_reset()
    ms = timed(lambda: step(x_dev, y_dev, True), args.steps)
    launches = int(lib.ssasr_launch_count())
    clocks = sampler.stop(first_sample) if rank == 0 else {}
    value = world * B * args.steps / (ms / 1e3)

    # e2e: host buffers in, loss + attention maps out, through the drop-in module API.  Every step copies its batch
    # from pinned host memory (on the copy stream of HostBatchPipeline, overlapping the previous step) and reads the loss
    # and the attention maps of that step back to the host (both land in pinned memory without blocking; the host waits for
    # them after it has enqueued the rest of the step, so it never stalls the GPU; all device work of every step is inside
    # the timed region, which ends with a device synchronisation).
    from ss_asr_b200.parallel import HostBatchPipeline
    pipe = HostBatchPipeline(dev)
    model.att_async = True
    e2e_state = {'left': 0, 'att_probe': 0.0, 'loss_host': torch.zeros((), pin_memory=True), 'loss_ready': torch.cuda.Event()}

    def e2e_step():
        xd, yd = pipe.take()
        e2e_state['left'] -= 1
        if e2e_state['left'] > 0:
            pipe.submit(x_host, y_host)
        model.att_on_device = False
        optim.zero_grad(set_to_none=True)
        _, logits, att = model(xd, ans_len, teacher=yd, state_len=lens)
        loss = asr_loss(logits, yd)
        # the step's result goes to the host as soon as it exists: non-blocking copy of the loss into pinned memory behind the
        # (already enqueued) copy of the attention maps, one event; backward and the optimiser are enqueued before the host waits
        e2e_state['loss_host'].copy_(loss.detach(), non_blocking=True)
        e2e_state['loss_ready'].record()
        sync.backward(loss)
        optim.step_clipped(5.0)
        e2e_state['loss_ready'].synchronize()    # host sync: the loss and the attention maps of THIS step are on the host
        v = float(e2e_state['loss_host'])
        e2e_state['att_probe'] = float(att[0, 0, 0])
        return v

    def e2e_run(n):
        e2e_state['left'] = n
        pipe.submit(x_host, y_host)              # the first batch's copy is inside the timed region as well
        for _ in range(n):
            e2e_step()
    e2e_run(2)
    ms_e2e = timed(lambda: e2e_run(args.steps), 1)
    model.att_async = False
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = x_host.numel() * 4 + y_host.numel() * 8
    d2h = 4 + B * ans_len * (T // 8) * 4

    # per-family CUDA-event timing (separate pass, not part of the numbers above)
    Fk.set_overlap_wgrad(False)          # kernels timed one at a time: no second stream sharing the SMs during this pass
    lib.ssasr_profile_enable(1)
    _lib.profile_read()
    nprof = 2
    for _ in range(nprof):
        step(x_dev, y_dev, True)
    prof = _lib.profile_read()
    lib.ssasr_profile_enable(0)
    Fk.set_overlap_wgrad(True)
    fl = train_flops(B, T, F, U)
    fam_ms = {k: v[0] / nprof for k, v in prof.items() if v[1] > 0}
    fam_n = {k: v[1] / nprof for k, v in prof.items() if v[1] > 0}
    fam_flops = {'gemm_f32': fl['gemm'], 'gemm_tc': fl['gemm'], 'rec_fwd_f32': fl['rec_fwd'], 'rec_bwd_f32': fl['rec_bwd'],
                 'rec_fwd_tc': fl['rec_fwd'], 'rec_bwd_tc': fl['rec_bwd'], 'attn_fwd': fl['attn_fwd'],
                 'attn_bwd': fl['attn_bwd']}
    if fam_ms.get('gemm_f32') and fam_ms.get('gemm_tc'):   # both GEMM kinds active: split by time is not meaningful
        fam_flops['gemm_f32'] = None
    dom = max(fam_ms, key=fam_ms.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback (B200_PROFILING.md)'
    ach = (fam_flops.get(dom) or 0.0) / (fam_ms[dom] / 1e3) / 1e12 if fam_ms[dom] > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'))).get(dom, {}).get('bytes_per_launch')
    except Exception:
        pass
    # every family against both ceilings (the graded `roofline` object below is the dominant family's entry)
    peak_bw = peaks.get('hbm_gbs', 6546.6)
    by = train_bytes(B, T, F, U)
    dep_steps = T + T // 2 + T // 4 + B
    rooflines = {}
    for k, v in fam_ms.items():
        ent = {'ms_per_step': round(v, 3), 'launches_per_step': fam_n[k]}
        if fam_flops.get(k):
            ent['tflops'] = round(fam_flops[k] / (v / 1e3) / 1e12, 2)
            ent['tensor_frac'] = round(ent['tflops'] / peak_tf, 4)
        if k in by:
            ent['hbm_gbps'] = round(by[k] / (v / 1e3) / 1e9, 1)
            ent['hbm_frac'] = round(ent['hbm_gbps'] / peak_bw, 4)
            ent['us_per_dependent_step'] = round(v * 1e3 / dep_steps, 3)
        rooflines[k] = ent
    # the dominant family is reported against the ceiling it is closer to (both fractions are in rooflines[dom])
    dom_ent = rooflines[dom]
    if dom_ent.get('hbm_frac', 0.0) > dom_ent.get('tensor_frac', 0.0):
        roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': dom_ent['hbm_gbps'], 'peak': peak_bw, 'unit': 'GB/s',
                    'frac': dom_ent['hbm_gbps'] / peak_bw, 'traffic': traffic,
                    'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs)' if peaks else 'fallback (B200_PROFILING.md)'}
    else:
        roofline = {'kernel': dom, 'bound': 'tensor', 'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s',
                    'frac': ach / peak_tf, 'traffic': traffic, 'peak_source': peak_src}
    roofline.update({
                'note': 'recurrent kernels are bound by the latency of %d dependent steps per pass (exchange + MMA issue + cell '
                        'math per step), not by a throughput ceiling: see rooflines[*].us_per_dependent_step and DESIGN.md §4'
                        % dep_steps,
                'rooflines': rooflines,
                'launches_per_step': fam_n[dom], 'ms_per_step_in_kernel': fam_ms[dom],
                'share_of_step': fam_ms[dom] / (ms / args.steps),
                'per_family_ms_per_step': {k: round(v, 3) for k, v in sorted(fam_ms.items(), key=lambda kv: -kv[1])},
                'whole_step_tflops': fl['total'] / (ms / args.steps / 1e3) / 1e12})

    extra = {}
    if not args.no_extras and world == 1:
        extra = run_extras(model, dev, args, peaks)
    elif not args.no_extras:
        extra = run_extras(model, dev, args, peaks, shard=(rank, world), dist=dist)

    cpu = None
    if rank == 0 and not args.no_cpu:
        v, spstep, cores = cpu_train_utt_per_s(1, 1 if not args.small else 0)
        cpu = {'value': v, 'unit': 'utt/s', 'cores': cores, 'kind': 'port',
               'sample': '1 timed + 1 warm-up train step of a %d-utterance batch of the same recipe (T=512,F=80,U=40) '
                         'through oracle/las_port.py (the reference torch call sequence), %.1f s/step'
                         % (CPU_SAMPLE['B'], spstep)}
    if rank == 0:
        line = {'metric': 'asr_train_utt_per_s', 'value': value, 'unit': 'utt/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
                'config': workload_config(world) if not args.small else dict(workload_config(world), small=cfg),
                'clocks': clocks,
                'e2e': {'value': e2e_value, 'unit': 'utt/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'ms_per_step': ms_e2e / args.steps},
                'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu, 'extra': extra}
        print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extras(model, dev, args, peaks, shard=(0, 1), dist=None):
    """C3 greedy decode (1000 utterances, bs=1 semantics, sharded by utterance) and C2 fbank (4096 x 10 s)."""
    from ss_asr_b200 import _lib
    from ss_asr_b200 import preprocess as PP
    rank, world = shard
    out = {}

    def tmax(ms):
        t = torch.tensor([ms], device=dev)
        if dist is not None and world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    # ---- decode
    n_total = 1000 if not args.small else 64
    g = torch.Generator().manual_seed(4321)
    Ts = sorted([int(v) for v in torch.randint(256, 513, (n_total,), generator=g)], reverse=True)
    mine = Ts[rank::world]                      # length-balanced round-robin shard
    xb = torch.zeros(len(mine), mine[0], 80)
    for i, t in enumerate(mine):
        xb[i, :t] = torch.randn(t, 80, generator=g)
    xb = xb.to(dev)
    model.decode_batch(xb, mine)                # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ids = model.decode_batch(xb, mine)
    e1.record()
    torch.cuda.synchronize()
    ms = tmax(e0.elapsed_time(e1))
    out['decode'] = {'workload': 'C3 greedy decode, %d utterances T~U[256,512], bs=1 semantics, 200-char cap, lm_weight 0'
                                 % n_total, 'utt_per_s': n_total / (ms / 1e3), 'ms': ms,
                     'chars_per_s': sum(len(i) for i in ids) * world / (ms / 1e3),
                     'precision': 'fp32 SIMT exact path'}
    model.decode_batch(xb, mine, precision='tf32x3')
    torch.cuda.synchronize()
    e0.record()
    ids2 = model.decode_batch(xb, mine, precision='tf32x3')
    e1.record()
    torch.cuda.synchronize()
    ms2 = tmax(e0.elapsed_time(e1))
    out['decode']['utt_per_s_tf32x3'] = n_total / (ms2 / 1e3)
    # decoding steps actually executed (the loop stops once every utterance has emitted EOS; 201 = nobody did before the cap)
    out['decode']['steps_run'] = int(getattr(model, 'last_decode_steps', 0))
    out['decode']['tf32x3_identical_transcripts'] = sum(a == b for a, b in zip(ids, ids2)) / max(1, len(ids))
    # headline decode figure: the fastest path whose transcripts are identical to the fp32 SIMT path in this very run
    out['decode']['utt_per_s_fp32_simt'] = out['decode']['utt_per_s']
    if out['decode']['tf32x3_identical_transcripts'] == 1.0 and ms2 < ms:
        out['decode'].update(utt_per_s=n_total / (ms2 / 1e3), ms=ms2, chars_per_s=sum(len(i) for i in ids2) * world / (ms2 / 1e3),
                             precision='tf32x3 / bf16x3 tensor-core exact path (transcripts identical to the fp32 SIMT path)')
    if rank == 0 and world == 1:
        # secondary: same utterances with the (randomly initialised, as in ASRTester) CharLM at lm_weight 0.5
        import torch.nn as nn

        class _LM(nn.Module):
            def __init__(self):
                super().__init__()
                self.emb = nn.Embedding(50, 128)
                self.layer_1 = nn.GRUCell(128, 128)
                self.layer_2 = nn.GRUCell(128, 128)
                self.out = nn.Linear(128, 50)
        torch.manual_seed(7)
        lm = _LM().to(dev)
        model.decode_batch(xb, mine, rnn_lm=lm, lm_weight=0.5)
        torch.cuda.synchronize()
        e0.record()
        model.decode_batch(xb, mine, rnn_lm=lm, lm_weight=0.5)
        e1.record()
        torch.cuda.synchronize()
        out['decode']['utt_per_s_lm05'] = n_total / (e0.elapsed_time(e1) / 1e3)
    del xb
    if rank == 0 and world == 1 and not args.small:
        out['long_c5'] = run_c5(dev)
    # ---- fbank
    n_utt = (4096 if not args.small else 256) // world
    n = 160000
    audio = 0.1 * torch.randn(n_utt * n, device=dev, generator=torch.Generator(device=dev).manual_seed(1234 + rank))
    off = [i * n for i in range(n_utt + 1)]
    plan = PP.FbankPlan(off, 16000, 80, device=dev)
    fb = plan.run(audio)
    torch.cuda.synchronize()
    reps = 5
    e0.record()
    for _ in range(reps):
        plan.run(audio, out=fb)
    e1.record()
    torch.cuda.synchronize()
    ms = tmax(e0.elapsed_time(e1)) / reps
    byts = n_utt * (4 * n + 4 * 80 * 1001)
    hbm = peaks.get('hbm_gbs', 6650.0)
    out['fbank'] = {'workload': 'C2 log-mel fbank, %d x 10 s @16 kHz, 80 mels (input+output %.2f GB > L2)'
                                % (n_utt * world, byts / 1e9), 'utt_per_s': n_utt * world / (ms / 1e3), 'ms': ms,
                    'roofline': {'bound': 'hbm', 'achieved': byts / (ms / 1e3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                                 'frac': byts / (ms / 1e3) / 1e9 / hbm, 'traffic': None}}
    return out


def run_c5(dev):
    """C5 (BASELINE.json configs[4]): long utterances (1600 frames), 512-dim BLSTM, B=32, U=100 -- train step timing."""
    from ss_asr_b200.asr import ASR
    from ss_asr_b200.functional import asr_loss
    torch.manual_seed(1)
    m = ASR(50, 512, 256, 128, 80, 0.9).to(dev)
    m.train_precision = 'bf16'
    m.train()
    m.att_on_device = True
    opt = torch.optim.Adadelta(m.parameters(), lr=1.0, eps=1e-8)
    x, lens, y = synth_batch(32, 1600, 80, 100)
    xd, yd = x.to(dev), y.to(dev)
    ans = int(max((y != 0).sum(-1) + 1)) - 1

    def step():
        opt.zero_grad(set_to_none=True)
        _, logits, _ = m(xd, ans, teacher=yd, state_len=lens)
        asr_loss(logits, yd).backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
        opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    return {'workload': 'C5 long-utterance LAS train step: B=32, T=1600, F=80, S_enc=512, S_dec=256, U=100',
            'ms_per_step': ms, 'utt_per_s': 32 / ms * 1e3}


_JSON_OUT = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--small', action='store_true', help='debug-sized shapes (not a bench number)')
    ap.add_argument('--no-extras', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'],
                    help='bf16: tcgen05 gate GEMMs (fp32 accumulate, fp32 recurrence); fp32: exact SIMT path')
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line.  Libraries write there too (NCCL's version banner when NCCL_DEBUG=VERSION comes
    # from the environment or from nccl.conf): file descriptor 1 is pointed at stderr for the whole run and the JSON line goes
    # to a private duplicate of the original stdout.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
    _JSON_OUT.flush()


if __name__ == '__main__':
    main()
