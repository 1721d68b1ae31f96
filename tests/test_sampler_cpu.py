"""The counter-based inverse-CDF sampler that replaces `torch.multinomial` on the decode path (agent_decoder.py:2162-2163,
2194: softmax -> top-k -> multinomial over the k probabilities).  `torch.multinomial`'s random stream cannot be reproduced
on another device, so parity of sampled rollouts is defined through THIS sampler (shared by the oracle and the CUDA path,
which are compared draw for draw in tests/test_gpu_rollout.py); here: it draws from the same DISTRIBUTION as the
reference's call - chi-square against the exact top-k probabilities and against `torch.multinomial` itself."""
import numpy as np
import torch
from scipy import stats

from oracle.agent_decoder_oracle import sample_topk, uniform01


def _draws(logits_row, k, n, seed):
    logits = logits_row[None].repeat(n, 1)
    return sample_topk(logits, k, seed, scene=3, it=7).numpy()


def test_uniform01_is_uniform():
    u = np.array([uniform01(2024, s, r, t) for s in range(4) for r in range(64) for t in range(40)])
    assert 0.0 <= u.min() and u.max() < 1.0
    counts, _ = np.histogram(u, bins=20, range=(0, 1))
    assert stats.chisquare(counts).pvalue > 1e-3
    # consecutive iterations of one row are not correlated
    a = u.reshape(-1, 40)
    assert abs(np.corrcoef(a[:, :-1].ravel(), a[:, 1:].ravel())[0, 1]) < 0.03


def test_topk_draws_follow_the_reference_distribution():
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(2048, generator=g) * 3.0
    k, n = 5, 20000
    prob = torch.softmax(logits, -1)
    top_p, top_i = torch.topk(prob, k)
    expect = (top_p / top_p.sum()).numpy()
    ours = _draws(logits, k, n, seed=11)
    assert set(np.unique(ours)) <= set(top_i.tolist())
    c_ours = np.array([(ours == int(i)).sum() for i in top_i])
    assert stats.chisquare(c_ours, expect * n).pvalue > 1e-3, (c_ours, expect * n)
    # the reference's own call on the same probabilities
    ref = top_i[torch.multinomial(top_p[None].repeat(n, 1), 1, generator=g)[:, 0]].numpy()
    c_ref = np.array([(ref == int(i)).sum() for i in top_i])
    assert stats.chi2_contingency(np.stack([c_ours, c_ref]))[1] > 1e-3, (c_ours, c_ref)


def test_greedy_is_argmax():
    g = torch.Generator().manual_seed(6)
    logits = torch.randn(37, 2048, generator=g)
    assert torch.equal(sample_topk(logits, 1, 1, 0, 0), logits.argmax(-1))
