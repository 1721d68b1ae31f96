"""Generate golden vectors by running the UNMODIFIED reference on CPU (build container only).

    python tests/golden/make_golden.py            # writes tests/golden/case_*.npz

The reference (`/root/reference/infgen/modules/agent_decoder.py:1605 InfGenAgentDecoder.inference`) is imported
through `oracle/shims` (stand-ins for torch_geometric / torch_cluster, whose semantics are documented there).
Inputs are regenerated from seeds by `infgen_b200.synth.make_scene` / `infgen_b200.weights.make_state_dict`
(numpy PCG64 -> bit-identical everywhere), so only the *outputs* are stored.

Cases (see `CASES` in tests/golden/cases.py):
  cfg0_a8    BASELINE.json configs[0]: 8 agents, 11-iteration greedy decode (66-step scene), full logits kept.
  ragged_a24 24 agents, half of them entering/exiting, ego at index 3 with filtered rows before it; the state
             head is live (insertion enabled but the seed head is biased so the reference inserts nobody).
  std_a64    64 agents, 2048 map tokens, 16 iterations greedy (the headline shape).
  insert_a12 12 agents, insertion stage live (reference run with DEBUG=1): one agent inserted per iteration.

    python tests/golden/make_golden.py [case ...]   # default: all cases
"""
import os
import sys
import time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('TQDM_DISABLE', '1')

from tests.golden.cases import CASES, build_case          # noqa: E402
from oracle.ref_runner import run_reference                # noqa: E402


def _stack_ragged(xs):
    """Per-iteration tensors whose row count grows (insertion): pad with NaN to the final row count."""
    n = max(x.shape[0] for x in xs)
    return torch.stack([torch.cat([x, x.new_full((n - x.shape[0], *x.shape[1:]), float('nan'))]) for x in xs])


def main():
    torch.manual_seed(0)
    names = sys.argv[1:] or list(CASES)
    for name in names:
        scene, sd, cfg, spec = build_case(name)
        t0 = time.time()
        if spec.get('debug_force_enter'):
            os.environ['DEBUG'] = '1'                               # agent_decoder.py:1888-1889
        try:
            r = run_reference(scene, sd, cfg)
        finally:
            os.environ.pop('DEBUG', None)
        out, tr = r['out'], r['trace']
        S = len(tr['token_logits'])
        logits = _stack_ragged(tr['token_logits'])                  # [S,A,2048]
        top_v, top_i = logits.nan_to_num(-1e30).topk(8, dim=-1)
        tr['head_in'] = list(_stack_ragged(tr['head_in']))
        tr['state_logits'] = list(_stack_ragged(tr['state_logits']))
        save = {
            'next_token_idx': out['next_token_idx'].numpy(), 'next_state_idx': out['next_state_idx'].numpy(),
            'pos_a': out['pos_a'].numpy(), 'head_a': out['head_a'].numpy(),
            'pred_traj': out['pred_traj'].numpy(), 'pred_head': out['pred_head'].numpy(),
            'pred_state': out['pred_state'].numpy(), 'pred_valid': out['pred_valid'].numpy(),
            'valid_mask': out['valid_mask'].numpy(), 'agent_id': out['agent_id'].numpy(),
            'pred_shape': out['pred_shape'].numpy(), 'eval_shape': out['eval_shape'].numpy(),
            'pred_type': out['pred_type'].numpy(), 'ego_index': np.int64(out['ego_index']),
            'head_in': torch.stack(tr['head_in']).numpy(), 'state_logits': torch.stack(tr['state_logits']).numpy(),
            'top8_logit': top_v.numpy(), 'top8_index': top_i.numpy(),
            'log_message': np.array(out['log_message']),
        }
        if spec.get('full_logits'):
            save['token_logits'] = logits.numpy()
        if not spec['disable_insertion']:
            save['n_rows'] = np.array([x.shape[0] for x in r['trace']['token_logits']])
            for k in ('next_state_prob_seed', 'next_pos_rel_prob_seed', 'grid_agent_occ_seed', 'grid_pt_occ_seed',
                      'grid_agent_occ_gt_seed'):
                save[k] = out[k].numpy()
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f'case_{name}.npz')
        np.savez_compressed(path, **save)
        print(f'{name}: {S} iterations, A={logits.shape[1]}, {time.time() - t0:.1f}s reference CPU, '
              f'{os.path.getsize(path) / 1e6:.2f} MB, log="{out["log_message"]}"')


if __name__ == '__main__':
    main()
