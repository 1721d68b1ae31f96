"""The CPU restatement of the reference map encoder (oracle/map_decoder_oracle.py) against golden
vectors written by the UNMODIFIED reference `InfGenMapDecoder.forward` (tests/golden/make_golden_map.py).  No GPU needed.
This pins the oracle of the NEXT row; the CUDA path for it is not built yet."""
import os
import numpy as np
import pytest
import torch

from tests.golden.make_golden_map import build_map_case, MAP_CASES

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('name', list(MAP_CASES))
def test_map_encoder_oracle_matches_reference_golden(name):
    from oracle.map_decoder_oracle import map_encode
    pt, sd, traj = build_map_case(name)
    z = np.load(os.path.join(GOLD, f'case_{name}.npz'))
    pt = dict(pt)
    pt['light_type'] = pt['polygon_light_type'][pt['polygon']]           # map_decoder.py:85-86
    with torch.no_grad():
        got = map_encode(sd, pt, traj, pl2pl_radius=10.0, max_num_neighbors=100)
    # the pt2pt graph: same edges in the same order (target-major, ascending source)
    assert np.array_equal(got['edge_src'].numpy(), z['edge_src'])
    assert np.array_equal(got['edge_dst'].numpy(), z['edge_dst'])
    np.testing.assert_allclose(got['x_pt'].numpy(), z['x_pt'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(got['map_next_token_prob'].numpy(), z['map_next_token_prob'], rtol=1e-4, atol=1e-5)
    assert np.array_equal(got['map_next_token_idx'].numpy()[:, 0], z['map_next_token_idx'][:, 0])     # arg-max token


def test_radius_graph_truncates_to_the_first_neighbours_by_index():
    """More candidates than max_num_neighbors: the first k by ascending index survive, the self loop is dropped after the
    truncation (torch_cluster semantics as fixed by oracle/shims/cluster.py)."""
    from oracle.map_decoder_oracle import radius_graph_first_k
    pos = torch.zeros(12, 2)
    pos[:, 0] = torch.arange(12) * 0.1                                    # all within r = 10 of each other
    src, dst = radius_graph_first_k(pos, 10.0, 4)
    for t in range(12):
        nb = src[dst == t].tolist()
        want = [i for i in range(5) if i != t]                            # first 4 + 1 candidates, minus the target itself
        assert nb == want, (t, nb, want)
