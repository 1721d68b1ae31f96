#!/usr/bin/env python
"""Benchmark of the closed-loop decode hot path (BASELINE.json metric: agent-steps/sec, 64 agents x 91 steps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4] [--scenes B]

A *step* is one closed-loop rollout - `InfGenAgentDecoder.inference` (reference agent_decoder.py:1605-2389): setup,
prefill, S decode iterations each with its insertion stage, read-back - of one batch of synthetic Waymo-shaped scenes.
Every configuration runs the decoder the way the reference's shipped configs do: insertion stage ON (the reference
cannot switch it off, infgen/model/infgen.py:75-76), top-5 motion tokens, top-10 insertion cells, random-init weights.

  --config 1   BASELINE configs[1]: one scene, 64 agents, 91 steps (S = 16)                  (default at --gpus 1)
  --config 2   BASELINE configs[2]: 32 scenes x 64 agents, 91 steps, one GPU
  --config 3   BASELINE configs[3]: 32 scenes per GPU (256 at 8 GPUs), scenes dealt round-robin from ONE global list
               (infgen_b200/sharding.py = Lightning's DistributedSampler, run.py:130-139), per-scene metrics gathered
               over NCCL after the rollouts                                                  (default at --gpus > 1)
  --config 4   BASELINE configs[4]: 150 s rollouts (num_recurrent_steps_val = 1500, S = 300), 8 scenes per GPU

  value        whole-job agent-steps/s (agents of the scenes x raw steps per rollout / time) with the scene tensors already
               resident in HBM, CUDA-event timed on the engine stream, L2 flushed between steps, max over ranks
  e2e          the same metric through the public call `B200AgentDecoder.inference_batch(data, map_enc)` with host
               tensors: host setup, pinned H2D, rollout, D2H, output dicts - wall clock
  roofline     dominant kernel class of a profiled replay of the same steps (CUDA events around every launch)
  cpu_baseline the reference's own CPU implementation on the host cores (oracle/_ref through oracle/shims when staged,
               else the CPU port oracle/agent_decoder_oracle.py), bounded sample

`--impl reference` times that CPU implementation alone on the same configuration.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault('TQDM_DISABLE', '1')

N_AGENTS, N_MAP, N_STEPS, NH = 64, 2048, 91, 11
METRIC, UNIT = 'agent-steps/sec', 'agent-steps/s'
SCENE_SEED0 = 13

CONFIGS = {
    1: dict(name='configs[1]', scenes_per_gpu=1, n_rec=-1,
            text='1 scene x 64 agents x 91 steps (16 decode iterations)'),
    2: dict(name='configs[2]', scenes_per_gpu=32, n_rec=-1,
            text='32 scenes x 64 agents x 91 steps (16 decode iterations) on one GPU'),
    3: dict(name='configs[3]', scenes_per_gpu=32, n_rec=-1,
            text='32 scenes per GPU x 64 agents x 91 steps (16 decode iterations), one global scene list sharded round-robin'),
    4: dict(name='configs[4]', scenes_per_gpu=8, n_rec=1500,
            text='8 scenes per GPU x 64 agents, 150 s rollouts (num_recurrent_steps_val = 1500: 300 decode iterations)'),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=0, choices=[0, 1, 2, 3, 4], help='BASELINE.json configs[i]; 0 = 1 at --gpus 1, 3 otherwise')
    ap.add_argument('--scenes', type=int, default=0, help='override the scenes per GPU of the configuration')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the side measurements (motion-only, configs[2], configs[4] shapes)')
    return ap.parse_args()


def workload(args):
    c = args.config or (1 if args.gpus == 1 else 3)
    w = dict(CONFIGS[c])
    w['id'] = c
    if args.scenes:
        w['scenes_per_gpu'] = args.scenes
    return w


def decoder_config(w, motion_only=False):
    from infgen_b200.config import DecoderConfig
    return DecoderConfig(motion_beam_size=5, insert_beam_size=10, disable_insertion=motion_only,
                         num_recurrent_steps_val=w['n_rec'])


def config_block(w, world):
    """Identical in both arms (the driver compares it)."""
    return {'workload': f"{w['name']}: {w['text']}; insertion stage on, top-5 motion / top-10 insertion sampling, 2048 map "
                        f"tokens per scene, random-init weights (seed 0), scene seeds {SCENE_SEED0}+i",
            'scenes_total': w['scenes_per_gpu'] * world, 'scenes_per_gpu': w['scenes_per_gpu'],
            'raw_steps_per_agent': NH + (w['n_rec'] if w['n_rec'] > 0 else N_STEPS - NH)}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return {'hbm_gbs': float(p['hbm_gbs']), 'bf16_tflops': float(p.get('bf16_tflops', 0) or 0),
                    'bf16_tflops_sustained': float(p.get('bf16_tflops_sustained', 0) or p.get('bf16_tflops', 0)),
                    'source': 'measured (MEASURED_PEAKS.json)'}
        except Exception:
            pass
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0,
            'source': 'fallback (B200_PROFILING.md)'}


def make_scenes(ids, cfg):
    from infgen_b200.synth import make_scene
    return [make_scene(SCENE_SEED0 + i, num_agents=N_AGENTS, num_map_tokens=N_MAP, num_steps=N_STEPS, ragged=0.0,
                       ego_index=5, cfg=cfg) for i in ids]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(', ') for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                continue
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation of the path
# ---------------------------------------------------------------------------------------------------------------------
class CpuReference:
    """`InfGenAgentDecoder.inference` on the host cores: the UNMODIFIED reference modules staged under oracle/_ref
    (oracle/make_ref.py) run through oracle/shims when present (kind 'reference'), else the CPU port (kind 'port')."""

    def __init__(self, sd, cfg):
        import torch
        self.sd, self.cfg = sd, cfg
        try:
            torch.set_num_threads(len(os.sched_getaffinity(0)))      # torchrun pins OMP_NUM_THREADS=1
        except Exception:
            pass
        self.cores = torch.get_num_threads()
        self.kind, self.dec = 'port', None
        try:
            from oracle import ref_runner
            if ref_runner.reference_available():
                self.dec = ref_runner.build_reference_decoder(sd, cfg)
                self.run_reference = ref_runner.run_reference
                self.kind = 'reference'
        except Exception as ex:                                       # fall back to the port, say why
            self.why_port = repr(ex)[:160]
        self.detail = ('the unmodified reference modules (oracle/_ref: infgen/modules/agent_decoder.py et al., staged by '
                       'oracle/make_ref.py) through the torch_geometric / torch_cluster stand-ins of oracle/shims'
                       if self.kind == 'reference' else
                       'CPU port of InfGenAgentDecoder.inference (oracle/agent_decoder_oracle.py, torch CPU ops)')

    def rollout(self, scene, max_iters=None):
        """One rollout (or its first max_iters iterations); returns seconds."""
        import torch
        t0 = time.perf_counter()
        if self.kind == 'reference':
            cfg = self.cfg
            if max_iters is not None:
                # the reference has no iteration limit: shorten the horizon instead (whole multiples of `shift` raw steps;
                # it cannot run a horizon shorter than the scene, agent_decoder.py:1638, so the scene is cut to fit)
                import copy, dataclasses
                n_rec = max_iters * cfg.shift
                cfg = dataclasses.replace(cfg, num_recurrent_steps_val=n_rec)
                scene = _cut_scene(scene, NH + n_rec) if NH + n_rec < N_STEPS else scene
                self.dec.num_recurrent_steps_val = n_rec
            torch.manual_seed(2024)
            self.run_reference(scene, self.sd, cfg, capture=False, decoder=self.dec)
        else:
            from oracle.agent_decoder_oracle import rollout
            rollout(scene, self.sd, self.cfg, seed=2024, scene_id=0, max_iters=max_iters)
        return time.perf_counter() - t0


def _cut_scene(scene, n_steps):
    """The first n_steps raw steps of a synthetic scene (token columns cut accordingly)."""
    import torch
    T = n_steps // 5
    out = {k: (dict(v) if isinstance(v, dict) else v) for k, v in scene.items()}
    ag = out['agent']
    for k, v in list(ag.items()):
        if isinstance(v, torch.Tensor) and v.dim() >= 2 and v.shape[0] == ag['token_idx'].shape[0]:
            if v.shape[1] == scene['agent']['token_idx'].shape[1]:
                ag[k] = v[:, :T].clone()
            elif v.shape[1] == N_STEPS:
                ag[k] = v[:, :n_steps].clone()
    return out


def cpu_measure(ref, scenes, S, steps, warmup, budget_s):
    """Bounded sample: the first `iters` decode iterations of one scene per step, sized so that the whole run stays within
    budget_s; throughput pro-rated to whole rollouts (early iterations have the fewest rows, which flatters the CPU)."""
    dt = ref.rollout(scenes[0], 1)
    iters = int(max(1, min(S, budget_s / max(1e-6, dt * (steps + warmup)))))
    if iters < S:
        dt2 = ref.rollout(scenes[0], min(S, 2 * iters))          # iterations get slower as columns fill: re-calibrate
        iters = int(max(1, min(S, min(S, 2 * iters) * budget_s / max(1e-6, dt2 * (steps + warmup)))))
    for i in range(warmup):
        ref.rollout(scenes[i % len(scenes)], iters)
    t0 = time.perf_counter()
    for i in range(steps):
        ref.rollout(scenes[i % len(scenes)], iters)
    total = time.perf_counter() - t0
    return total, iters


def run_reference(args, rank, world):
    """Reference arm: rank 0 alone runs the CPU implementation; the other ranks exit without work."""
    if rank != 0:
        return
    from infgen_b200.weights import make_state_dict
    w = workload(args)
    cfg = decoder_config(w)
    sd = make_state_dict(0)
    n_sample = min(w['scenes_per_gpu'] * world, 4)
    scenes = make_scenes(range(n_sample), cfg)
    ref = CpuReference(sd, cfg)
    steps_per_agent = NH + (w['n_rec'] if w['n_rec'] > 0 else N_STEPS - NH)
    S = (steps_per_agent - NH) // 5
    total, iters = cpu_measure(ref, scenes, S, args.steps, args.warmup, 150.0)
    agent_steps = N_AGENTS * steps_per_agent * iters / S * args.steps
    value = agent_steps / total
    sample = (f'{iters} of {S} decode iterations of one 64-agent scene per step (scenes {SCENE_SEED0}..{SCENE_SEED0 + n_sample - 1} of the '
              f'workload in turn), pro-rated to whole rollouts')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': total / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config_block(w, world),
        'impl_detail': ref.detail,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': ref.cores, 'kind': ref.kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def kernel_profile(dec, hb, hosts, n_prof=2):
    """Per-kernel-class CUDA-event timing of profiled rollouts + the algorithmic work of each class per launch."""
    import numpy as np
    pk = peaks()
    dec.set_profile(True)
    S = hb.S
    e_t = e_m = e_a = 0
    rows_sum = 0
    for rep in range(n_prof):
        dec.load(hb, hosts)
        dec.prefill()
        for it in range(S):
            dec.step(1)
            if rep == 0 and (S <= 20 or it % max(1, S // 20) == 0):
                e_t += int(dec.debug_read('t_cnt', (hb.R,), np.int32).sum())
                e_m += int(dec.debug_read('m_cnt', (hb.R,), np.int32).sum())
                e_a += int(dec.debug_read('a_cnt', (hb.R,), np.int32).sum())
                rows_sum += int(dec.debug_read('n_rows', (hb.n_scenes,), np.int32).sum())
                e_t, e_m, e_a = e_t, e_m, e_a
        dec.synchronize()
    prof = dec.profile()
    dec.set_profile(False)
    sampled = S if S <= 20 else len(range(0, S, max(1, S // 20)))
    rows = rows_sum / max(1, sampled)                  # average active rows per iteration (inserted rows included)
    et, em, ea = e_t / sampled, e_m / sampled, e_a / sampled
    # algorithmic work per launch (fp32; DESIGN.md "roofline accounting").  Attention gather: per edge the K
    # row, the V row (512 B each) and the relative embedding row (512 B, recomputable but here materialised once per
    # iteration and read by the six layers of its stack) + 4 B source index.  Dense half: folded formulation, 2 * MAC.
    per_edge_kv, per_edge_all = 1024, 512 + 512 + 512 + 4
    post_mac, pre_mac, pre_kv_mac = 196608, 49152, 32768     # Wvr+gate+out+ffn | q,s,Wkr fold | k,v
    w_layer = 1.125e6 + 0.17e6                                # post + pre chunks of one layer (fp32 bytes)
    edge_flops = 2.0 * 2 * (128 + 16) * 8                     # score + weighted sum, 8 heads
    tm_bytes = rows * (1024 + 5120 + 1024 + 1024) + (et + em) * per_edge_all + 2 * w_layer
    a_bytes = rows * (1024 + 5120 + 5120 + 1024) + ea * per_edge_all + w_layer
    tm_flops = rows * 2.0 * (2 * post_mac + 2 * pre_mac + pre_kv_mac) + (et + em) * edge_flops
    a_flops = rows * 2.0 * (post_mac + pre_mac + pre_kv_mac) + ea * edge_flops
    st_bytes = rows * (1024 + 12 * 1024) + 6 * (et + em + ea) * per_edge_all + 18 * w_layer
    st_flops = 6 * (tm_flops + a_flops)
    e_avg = (et + em + ea) / 3.0
    attn_kv_bytes = e_avg * per_edge_kv                       # SURVEY 8d's canonical figure: K, V rows only
    attn_all_bytes = e_avg * per_edge_all + rows * (512 + 4096 + 512 + 4096 + 32)
    attn_flops = e_avg * edge_flops
    node_flops = rows * 2.0 * (post_mac + pre_mac + pre_kv_mac * 2.0 / 3.0)
    node_bytes = 1.05e6 + rows * (1024 + 512 + 4096 + 32 + 512 + 512 + 512 + 4096 + 1024 * 2.0 / 3.0)
    model = {
        'k_layer:stack18': ('hbm', st_bytes, st_flops),
        'k_layer:temporal+map': ('hbm', tm_bytes, tm_flops),
        'k_layer:agent': ('hbm', a_bytes, a_flops),
        'k_attn': ('hbm', attn_kv_bytes, attn_flops),
        'k_node': ('tensor', node_flops, node_bytes),
        'k_fourier:edges': ('tensor', (et * (4 * (132 * 128 + 128 * 128) + 128 * 128) +
                                       (em + ea) * (3 * (132 * 128 + 128 * 128) + 128 * 128)) * 2.0),
        'k_embed_column': ('tensor', rows * 2.0 * (2 * (132 * 128 + 128 * 128) + 128 * 128 + 512 * 128 + 2 * 128 * 128)),
        'k_heads': ('tensor', rows * 2.0 * (2 * 128 * 128 + 128 * 2048 + 128 * 128)),
    }
    pipes = {'k_fourier:edges': 'tcgen05.mma kind::tf32, 3xTF32 split (3 tensor-core passes per algorithmic flop), fp32 '
                                'accumulate in TMEM',
             'k_node': getattr(dec, 'node_pipe', 'fp32 FFMA (reference numerics are fp32)')}
    klist = []
    tot = sum(v['ms'] for v in prof.values())
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
        avg_us = v['ms'] * 1e3 / v['launches']
        ent = {'kernel': name, 'share': v['ms'] / tot, 'avg_us': avg_us, 'launches_per_rollout': v['launches'] / n_prof}
        if name in model:
            bound, work = model[name][0], model[name][1]
            if bound == 'hbm':
                ach = work / (avg_us * 1e-6) / 1e9
                ent.update({'bound': 'hbm', 'achieved': ach, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                            'frac': ach / pk['hbm_gbs'], 'algorithmic_bytes_per_launch': work})
                if name == 'k_attn':
                    ent['bytes_definition'] = 'K and V rows of every edge (1,024 B per edge, SURVEY.md 8d), nothing else'
                    ent['achieved_incl_rhat_and_handover'] = attn_all_bytes / (avg_us * 1e-6) / 1e9
                    ent['frac_incl_rhat_and_handover'] = ent['achieved_incl_rhat_and_handover'] / pk['hbm_gbs']
                if len(model[name]) > 2:
                    fl = model[name][2] / (avg_us * 1e-6) / 1e12
                    ent.update({'algorithmic_flops_per_launch': model[name][2], 'fp32_tflops': fl,
                                'frac_of_fp32_ffma_peak': fl / 72.0})
            else:
                ach = work / (avg_us * 1e-6) / 1e12
                ent.update({'bound': 'tensor', 'achieved': ach, 'peak': pk['bf16_tflops_sustained'],
                            'unit': 'TFLOP/s', 'frac': ach / pk['bf16_tflops_sustained'],
                            'algorithmic_flops_per_launch': work,
                            'pipe': pipes.get(name, 'fp32 FFMA (reference numerics are fp32)'),
                            'frac_of_fp32_ffma_peak': ach / 72.0})
                if len(model[name]) > 2:
                    ent['algorithmic_bytes_per_launch'] = model[name][2]
        klist.append(ent)
    return klist, pk, {'rows_avg': rows, 'E_t_avg': et, 'E_m_avg': em, 'E_a_avg': ea}


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the round's ncu --set full capture (profiles/r2_ncu_traffic.json, written by
    tools/ncu_traffic.py from the .ncu-rep), or None."""
    path = os.path.join(ROOT, 'profiles', 'r2_ncu_traffic.json')
    if not os.path.exists(path):
        return None
    try:
        t = json.load(open(path))
        for k, v in t.get('kernels', {}).items():
            if kernel.split(':')[0] in k:
                return v.get('dram_bytes_per_launch')
    except Exception:
        pass
    return None


def timed_device_rollouts(runs, steps, warmup, stream, flush, barrier):
    """load + rollout + read with device-resident inputs / results.  runs: [(decoder, DeviceBatch, hosts, stream)] - one
    entry per engine of the decoder's group (several engines roll their shares of a batch out concurrently, each on its own
    stream).  Timed with CUDA events on `stream`: the engine streams wait for the start event and `stream` waits for an
    event behind every engine's last copy before the end event is recorded."""
    import torch

    def one_step():
        # every engine's rollout is enqueued before the first read (infgen_read waits for the row counts of its engine: the
        # insertion records that travel depend on them)
        for d, db, hosts, _ in runs:
            d.load(db, hosts); d.rollout()
        for d, _, _, _ in runs:
            d.read()

    with torch.cuda.stream(stream):
        for _ in range(max(3, warmup)):
            one_step()
        barrier()
        l0 = sum(d.kernel_launches() for d, _, _, _ in runs)
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _, _, _, st in runs:
                if st is not stream:
                    st.wait_event(e0)
            one_step()
            for _, _, _, st in runs:
                if st is not stream:
                    done = torch.cuda.Event()
                    done.record(st)
                    stream.wait_event(done)
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        for d, _, _, _ in runs:
            d.synchronize()                               # surfaces device-side errors of the device-resident rollouts
        launches = sum(d.kernel_launches() for d, _, _, _ in runs) - l0
    return [a.elapsed_time(b) for a, b in evs], launches


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    # torchrun pins OMP_NUM_THREADS to 1 per rank; the host side of the public call (output assembly: index_put_ over the
    # insertion records of a batch) uses the rank's share of the host cores
    torch.set_num_threads(max(1, min(16, (os.cpu_count() or 1) // max(world, 1))))
    from infgen_b200.weights import make_state_dict
    from infgen_b200.agent_decoder import B200AgentDecoder
    from infgen_b200.host import prepare_scene, HostBatch, DeviceBatch
    from infgen_b200.sharding import shard_scenes, gather_scene_metrics

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    w = workload(args)
    cfg = decoder_config(w)
    sd = make_state_dict(0)
    n_total = w['scenes_per_gpu'] * world
    my_ids = shard_scenes(n_total, rank, world)                      # scene i -> rank i mod world
    scenes = make_scenes(my_ids, cfg)
    maps = [s['map_enc'] for s in scenes]
    steps_per_agent = NH + (w['n_rec'] if w['n_rec'] > 0 else N_STEPS - NH)
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def measure(dec_cfg, scn, mps, ids, steps, warmup, sampler_rank0=False, e2e=True):
        """value (device-resident) and e2e (public call) of one decoder configuration on one list of scenes."""
        kw = {'scenes_per_engine': int(os.environ['INFGEN_SCENES_PER_ENGINE'])} if 'INFGEN_SCENES_PER_ENGINE' in os.environ else {}
        if 'INFGEN_MAX_ENGINES' in os.environ:
            kw['max_engines'] = int(os.environ['INFGEN_MAX_ENGINES'])
        dec = B200AgentDecoder(sd, dec_cfg, device=local_rank, seed=2024, use_cuda_graph=True, **kw)
        # the public call first: it also settles the row capacity (the insertion stage may need capacity reruns)
        for _ in range(2):
            outs = dec.inference_batch(scn, mps, scene_ids=ids)
        # the engines of the call (a batch of more than `scenes_per_engine` scenes is dealt to several, each with its own
        # stream and iteration graph) and the scenes each one served
        groups = dec._groups or [(dec, list(range(len(scn))))]
        runs, hosts, hbs = [], [], []
        for gi, (d, pos) in enumerate(groups):
            hosts_g = list(d._scenes)
            hb = HostBatch(hosts_g, dec_cfg, scene_ids=[ids[i] for i in pos], row_capacity=d._batch.cap)
            st = stream if gi == 0 else torch.cuda.Stream(device=dev)
            d.set_stream(st.cuda_stream)
            with torch.cuda.stream(st):
                db = DeviceBatch(hb, dev)
            runs.append((d, db, hosts_g, st))
            hosts += hosts_g
            hbs.append(hb)
        torch.cuda.synchronize(dev)
        sampler = ClockSampler(local_rank) if sampler_rank0 and rank == 0 else None
        step_ms, launches = timed_device_rollouts(runs, steps, warmup, stream, flush, barrier)
        clocks = sampler.stop() if sampler else None
        for d, _, _, _ in runs:
            d.set_stream(None)
        # (the kernel profile of the roofline leg replays the WHOLE batch on engine 0: per-launch figures at the batch size
        #  the configuration names, whatever the number of engines the timed rollouts were dealt to)
        cap = max(d._batch.cap for d, _ in groups)
        hb_all = hbs[0] if len(groups) == 1 else HostBatch(hosts, dec_cfg, scene_ids=ids, row_capacity=cap)
        res = {'dec': dec, 'hb': hb_all, 'hosts': hosts, 'outs': outs, 'step_ms': step_ms,
               'launches': launches, 'clocks': clocks, 'cap': cap, 'engines': len(groups),
               'scenes_per_engine': [len(pos) for _, pos in groups]}
        if e2e:
            for _ in range(max(1, warmup - 2)):
                dec.inference_batch(scn, mps, scene_ids=ids)
            barrier()
            ts = []
            for _ in range(steps):
                flush.fill_(1)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                res['outs'] = dec.inference_batch(scn, mps, scene_ids=ids)
                ts.append(time.perf_counter() - t0)
            barrier()
            res['e2e_s'] = ts
            gs = dec._groups or [(dec, None)]
            res['h2d'], res['d2h'] = sum(d._batch.h2d_bytes() for d, _ in gs), sum(d._batch.d2h_bytes() for d, _ in gs)
        return res

    m = measure(cfg, scenes, maps, my_ids, args.steps, args.warmup, sampler_rank0=True)
    dec = m['dec']
    agent_steps = sum(h.n_rows for h in m['hosts']) * steps_per_agent
    total_ms = torch.tensor([sum(m['step_ms'])], dtype=torch.float64, device=dev)
    e2e_total = torch.tensor([sum(m['e2e_s'])], dtype=torch.float64, device=dev)
    n_agent_steps = torch.tensor([float(agent_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_agent_steps, op=dist.ReduceOp.SUM)
    total_ms, e2e_total_s, all_agent_steps = float(total_ms.item()), float(e2e_total.item()), float(n_agent_steps.item())
    value = all_agent_steps * args.steps / (total_ms * 1e-3)
    e2e_value = all_agent_steps * args.steps / e2e_total_s

    # ---- final metric gather: one row per scene of the GLOBAL list (the only collective of the path) ------------------
    nh = cfg.num_historical_steps
    local = torch.tensor([[float((o['pred_traj'][:N_AGENTS, -1] - o['pred_traj'][:N_AGENTS, nh]).norm(dim=-1).mean()),
                           float(o['pos_a'].shape[0]), float(o['pos_a'].shape[0] - h.n_rows)]
                          for o, h in zip(m['outs'], m['hosts'])], device=dev)
    full = gather_scene_metrics(my_ids, local, n_total)
    gathered = {'scenes': n_total, 'mean_displacement_m': float(full[:, 0].mean()), 'rows_final_mean': float(full[:, 1].mean()),
                'rows_final_max': int(full[:, 1].max()), 'inserted_agents_total': int(full[:, 2].sum()),
                'collective': 'all_gather (NCCL)' if world > 1 else 'none (1 rank)'}

    # ---- roofline: profiled replay (events around every launch; no graph) ------------------------------------------
    roof, klist, wl_stats = None, [], None
    if rank == 0:
        klist, pk, wl_stats = kernel_profile(dec, m['hb'], m['hosts'], n_prof=2 if w['n_rec'] <= 0 else 1)
        top = next((k for k in klist if 'bound' in k), None)
        if top:
            roof = {'kernel': top['kernel'], 'bound': top['bound'], 'achieved': top['achieved'], 'peak': top['peak'],
                    'unit': top['unit'], 'frac': top['frac'], 'traffic': ncu_traffic(top['kernel']),
                    'peak_source': pk['source'], 'share_of_step': top['share'], 'avg_launch_us': top['avg_us'],
                    'note': ('one 64-agent scene is latency-bound end to end (dependent exchange phases on 64 SMs, DESIGN.md); '
                             'the bandwidth-bound form of the same attention is k_attn in configs2.roofline_kernels')
                    if top['kernel'].startswith('k_layer') else None}
    dec.close()

    # ---- side measurements (rank 0, single GPU, default configuration only) -----------------------------------------
    extra = {}
    if rank == 0 and world == 1 and w['id'] == 1 and not args.no_extra:
        def side(name, fn):
            try:
                extra[name] = fn()
            except Exception as ex:          # the headline line must still be printed
                extra[name] = {'error': repr(ex)[:300]}

        def motion_only():
            mo = measure(decoder_config(w, motion_only=True), scenes, maps, my_ids, max(5, args.steps // 2), 3)
            mo['dec'].close()
            a = sum(h.n_rows for h in mo['hosts']) * steps_per_agent
            return {'workload': 'configs[1] scene with the insertion stage disabled (round-1 headline mode; the reference '
                                'cannot run this mode through InfGen, infgen/model/infgen.py:75-76)',
                    'value': a / (statistics.mean(mo['step_ms']) * 1e-3), 'ms_per_step': statistics.mean(mo['step_ms']),
                    'e2e_value': a / statistics.mean(mo['e2e_s']), 'e2e_ms_per_step': statistics.mean(mo['e2e_s']) * 1e3,
                    'unit': UNIT}

        def batch(cid, n_steps):
            wc = dict(CONFIGS[cid]); wc['id'] = cid
            ccfg = decoder_config(wc)
            ids = list(range(wc['scenes_per_gpu']))
            scn = make_scenes(ids, ccfg)
            mps = [s['map_enc'] for s in scn]
            mb = measure(ccfg, scn, mps, ids, n_steps, 3)
            spa = NH + (wc['n_rec'] if wc['n_rec'] > 0 else N_STEPS - NH)
            a = sum(h.n_rows for h in mb['hosts']) * spa
            kl, _, st = kernel_profile(mb['dec'], mb['hb'], mb['hosts'], n_prof=1)
            mb['dec'].close()
            rows_final = [int(o['pos_a'].shape[0]) for o in mb['outs']]
            return {'workload': f"{wc['name']} on one GPU: {wc['text']}; inputs in HBM for `value`, public call for `e2e_value`",
                    'value': a / (statistics.mean(mb['step_ms']) * 1e-3), 'ms_per_step': statistics.mean(mb['step_ms']),
                    'e2e_value': a / statistics.mean(mb['e2e_s']), 'e2e_ms_per_step': statistics.mean(mb['e2e_s']) * 1e3,
                    'unit': UNIT, 'engines': mb['engines'], 'scenes_per_engine': mb['scenes_per_engine'],
                    'roofline_kernels_note': 'profiled replay of the whole batch on ONE engine' if mb['engines'] > 1 else None,
                    'row_capacity': mb['cap'], 'rows_final_mean': statistics.mean(rows_final),
                    'rows_final_max': max(rows_final), 'launches_per_step': mb['launches'] / n_steps,
                    'workload_stats': st, 'roofline_kernels': kl}

        def map_encoder():
            """Row f1: `InfGenMapDecoder.forward` (runs once per scene right before the decode, infgen_decoder.py:124) on the
            same engine; x_pt stays in HBM for the decode that follows."""
            from infgen_b200.weights import make_map_state_dict
            from infgen_b200.synth import make_map_tokens
            from infgen_b200.map_encoder import load_map_vocab
            msd, traj = make_map_state_dict(0), load_map_vocab()
            datas = []
            for i, s_ in enumerate(scenes):
                pt = make_map_tokens(s_, i)
                d = dict(s_)
                d['pt_token'] = dict(s_['pt_token'])
                d['pt_token'].update({k: pt[k] for k in ('type', 'pl_type', 'token_idx', 'pt_pred_mask', 'pt_valid_mask', 'pt_target_mask')})
                d['pt_token']['light_type'] = pt['polygon_light_type'][pt['polygon']]
                datas.append(d)
            fdec = B200AgentDecoder(sd, cfg, device=local_rank, seed=2024, use_cuda_graph=True, map_state_dict=msd, map_traj_src=traj)
            for _ in range(3):
                fdec.map_encode(datas, want_x=False)
            torch.cuda.synchronize(dev)
            tm = []
            for _ in range(10):
                t0 = time.perf_counter()
                fdec.map_encode(datas, want_x=False)
                tm.append(time.perf_counter() - t0)
            for _ in range(3):
                fdec.inference_batch(datas, None, scene_ids=my_ids)
            te = []
            for _ in range(max(5, args.steps // 2)):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                fdec.inference_batch(datas, None, scene_ids=my_ids)
                te.append(time.perf_counter() - t0)
            fdec.close()
            out = {'workload': 'map encoder of the configs[1] scene (2048 map tokens, radius graph r = 10 m <= 100 neighbours, 3 '
                               'pt2pt AttentionLayers) and the decode fed from it on ONE engine, host tensors in, x_pt kept in HBM',
                   'map_encode_ms': statistics.mean(tm) * 1e3,
                   'e2e_map_plus_decode_ms': statistics.mean(te) * 1e3,
                   'e2e_map_plus_decode_value': agent_steps / statistics.mean(te), 'unit': UNIT}
            try:
                from oracle.map_decoder_oracle import map_encode as cpu_map
                q = dict(datas[0]['pt_token'])
                t0 = time.perf_counter()
                with torch.no_grad():
                    cpu_map(msd, q, traj)
                out['cpu_port_map_encode_ms'] = (time.perf_counter() - t0) * 1e3
            except Exception as ex:
                out['cpu_port_error'] = repr(ex)[:200]
            return out

        def scene_prep():
            """Row f2: `TokenProcessor._tokenize_agent` + `InfGen._fetch_enterings` of the configs[1] scene's raw 10 Hz tracks on
            the engine (infgen_prepare_scene), host tensors in and out, next to the CPU port on the host cores."""
            from infgen_b200.scene_prep import B200ScenePrep
            pdec = B200AgentDecoder(sd, cfg, device=local_rank, seed=2024, use_cuda_graph=False)
            prep = B200ScenePrep(pdec)
            ag = scenes[0]['agent']
            raw = {k: ag[k] for k in ('valid_mask', 'heading', 'position', 'velocity', 'type', 'shape')}
            raw['av_idx'] = ag['av_index']
            mk = lambda: {'agent': {k: v.clone() for k, v in raw.items()}, 'pt_token': {'position': scenes[0]['pt_token']['position']}}
            for _ in range(3):
                prep.tokenize(mk())
            ts = []
            for _ in range(20):
                d = mk()
                t0 = time.perf_counter()
                prep.tokenize(d)
                ts.append(time.perf_counter() - t0)
            # map side (InfGen.match_token_map): 2,048 polylines of 5 m against the 1,024-entry map vocabulary
            from infgen_b200.map_encoder import load_map_vocab
            src = load_map_vocab()
            sp = src[:, torch.linspace(0, src.shape[1] - 1, steps=3).long()].contiguous()
            rng = np.random.default_rng(11)
            P = 2048
            ids, theta = rng.integers(0, sp.shape[0], size=P), rng.uniform(-np.pi, np.pi, size=P).astype(np.float32)
            c_, s_ = np.cos(theta)[:, None], np.sin(theta)[:, None]
            loc = sp.numpy()[ids] + rng.normal(0.0, 0.05, size=(P, 3, 2)).astype(np.float32)
            world_pts = np.stack([loc[..., 0] * c_ - loc[..., 1] * s_, loc[..., 0] * s_ + loc[..., 1] * c_], -1)
            world_pts = (world_pts - world_pts[:, :1] + rng.uniform(-150.0, 150.0, size=(P, 1, 2))).astype(np.float32)
            md = {'map_save': {'traj_pos': torch.from_numpy(world_pts), 'traj_theta': torch.from_numpy(theta),
                               'pl_idx_list': torch.from_numpy(np.sort(rng.integers(0, 64, size=P)).astype(np.float32))},
                  'pt_token': {'side': torch.zeros(P, dtype=torch.uint8), 'num_nodes': P}}
            for _ in range(3):
                prep.match_token_map(md)
            tm = []
            for _ in range(10):
                t0 = time.perf_counter()
                prep.match_token_map(md)
                tm.append(time.perf_counter() - t0)
            pdec.close()
            out = {'workload': 'agent tokenizer + enterings of the configs[1] scene (64 agents x 91 raw steps, 2048 map tokens): '
                               'closed-loop match over 2048 vocabulary boxes per token step, ego-centric grid cells of agents and '
                               'map tokens per column', 'gpu_ms': statistics.mean(ts) * 1e3, 'gpu_ms_min': min(ts) * 1e3,
                   'map_match_workload': 'InfGen.match_token_map: 2,048 map polylines against the 1,024-entry map vocabulary, '
                   'host tensors in and out', 'map_match_gpu_ms': statistics.mean(tm) * 1e3}
            try:
                from oracle.scene_prep_oracle import tokenize_agent, fetch_enterings
                from infgen_b200.synth import load_vocab
                from infgen_b200.grid import PositionGrid
                g = PositionGrid(cfg.grid_range, cfg.grid_interval, cfg.pl2seed_radius, cfg.angle_interval)
                t0 = time.perf_counter()
                tk = tokenize_agent(raw, load_vocab())
                fetch_enterings(tk, scenes[0]['pt_token']['position'], int(raw['av_idx'][0]), g.cells, cfg.pl2seed_radius, cfg.angle_interval)
                out['cpu_port_ms'] = (time.perf_counter() - t0) * 1e3
                from oracle.scene_prep_oracle import match_token_map
                t0 = time.perf_counter()
                match_token_map(world_pts, theta, md['map_save']['pl_idx_list'], md['pt_token']['side'], sp)
                out['map_match_cpu_port_ms'] = (time.perf_counter() - t0) * 1e3
            except Exception as ex:
                out['cpu_port_error'] = repr(ex)[:200]
            return out

        side('motion_only', motion_only)
        side('scene_prep', scene_prep)
        side('map_encoder', map_encoder)
        side('configs2', lambda: batch(2, 5))
        side('configs4_one_gpu', lambda: batch(4, 2))

    # ---- CPU baseline (rank 0, N=1 only): the reference's own implementation, bounded sample -----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref = CpuReference(sd, cfg)
            S = (steps_per_agent - NH) // 5
            total, iters = cpu_measure(ref, scenes[:1], S, 2, 0, 24.0)
            cpu = {'value': N_AGENTS * steps_per_agent * iters / S * 2 / total, 'unit': UNIT, 'cores': ref.cores,
                   'kind': ref.kind, 'sample': f'{iters} of {S} decode iterations of the first scene, 2 repetitions, '
                   f'{total:.1f} s, pro-rated to whole rollouts', 'impl_detail': ref.detail}
        except Exception as ex:
            cpu = {'error': repr(ex)[:300]}

    if rank == 0:
        rows_final = [int(o['pos_a'].shape[0]) for o in m['outs']]
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(3, args.warmup), 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config_block(w, world),
            'impl_detail': {'parallelism': f'scenes sharded round-robin over {world} rank(s), no data-path collective; per GPU '
                            f"{m['engines']} engine(s) x {m['scenes_per_engine']} scenes, one rollout stream and iteration graph "
                            'per engine, rollouts concurrent (roofline_kernels: profiled replay of the whole per-GPU batch on '
                            'one engine)', 'l2': 'flushed between steps (256 MiB write)',
                            'timed_region': 'infgen_load_scenes (device copies, map K/V caches) + prefill + S iterations '
                            '(one CUDA graph per iteration: insertion stage with WHILE / IF conditional nodes + motion '
                            'stage) + result copies', 'row_capacity': m['cap'],
                            'rows_final_rank0': rows_final[:8], 'agent_steps_counted': 'agents of the scenes at the current '
                            'step x raw steps (agents appended by the insertion stage are extra work, not counted)'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': m['h2d'], 'd2h_bytes_per_step': m['d2h'],
                    'ms_per_step': e2e_total_s / args.steps * 1e3,
                    'call': 'B200AgentDecoder.inference_batch(data, map_enc) incl. host setup and output dicts'},
            'gpu_launches': int(m['launches']), 'clocks': m['clocks'], 'roofline': roof, 'roofline_kernels': klist,
            'workload_stats': wl_stats, 'cpu_baseline': cpu, 'final_gather': gathered,
        }
        if world > 1:
            line['weak_scaling_base'] = ('per-GPU work is fixed at %d scenes; the same workload on one GPU is the `configs2` block '
                                         'of the --gpus 1 line (or `--gpus 1 --config 3`)' % w['scenes_per_gpu'])
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
