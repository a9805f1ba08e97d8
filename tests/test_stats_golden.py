"""CPU: UI-search statistics (RunningTopK, RunningQuantile, tally drivers) against golden
vectors produced by the live reference (oracle/make_golden_stats.py).  These classes are plain
torch (device-agnostic, like the reference's), so parity is checked here without a GPU."""
import os

import numpy as np
import pytest
import torch

from rewriting_b200.utils import runningstats as rs, tally

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'stats_topk_quantile.npz')
QS = [0.001, 0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 0.999]
PROBES = torch.tensor([[-2.0, -0.1, 0.0, 0.3, 5.0]] * 3)


@pytest.fixture(scope='module')
def gold():
    return dict(np.load(GOLD, allow_pickle=True))


@pytest.fixture(scope='module')
def x():
    torch.manual_seed(1234)
    return torch.randn(6000, 3) ** 3


def test_running_topk_matches_reference(gold, x):
    tk = rs.RunningTopK(k=5)
    for i in range(0, 6000, 500):
        tk.add(x[i:i + 500])
    v, idx = tk.result()
    assert tk.size() == 6000
    np.testing.assert_array_equal(v.numpy(), gold['topk_values'])
    np.testing.assert_array_equal(idx.numpy(), gold['topk_index'])
    back = rs.RunningTopK(state=tk.state_dict())                # state round trip
    np.testing.assert_array_equal(back.result()[0].numpy(), gold['topk_values'])
    np.testing.assert_array_equal(back.result()[1].numpy(), gold['topk_index'])
    # explicit sample names, ragged last batch, k larger than a batch
    tk2 = rs.RunningTopK(k=700)
    for i in range(0, 1300, 600):
        tk2.add(x[i:min(i + 600, 1300), 0], index=torch.arange(i, min(i + 600, 1300)) + 10)
    v2, i2 = tk2.result()
    want_v, want_i = x[:1300, 0].topk(700)
    assert torch.equal(v2, want_v) and torch.equal(i2, want_i + 10)


def test_running_quantile_exact_regime_matches_reference(gold, x):
    rq = rs.RunningQuantile(r=1024)
    for i in range(0, 1500, 500):
        rq.add(x[i:i + 500])
    np.testing.assert_allclose(rq.quantiles(QS).numpy(), gold['small_quantiles'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rq.normalize(PROBES).numpy(), gold['small_normalize'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rq.median().numpy(), gold['small_median'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rq.mean().numpy(), gold['small_mean'], rtol=1e-5, atol=1e-6)
    assert rq.quantiles(0.999).shape == (3,)                    # scalar q, as rewriteapp.py:264 uses it
    assert rq.size() == 1500 and rq.minmax().shape == (3, 2)


def test_reference_sketch_state_loads_and_reads_out_identically(gold):
    state = {k[len('big_state.'):]: v for k, v in gold.items() if k.startswith('big_state.')}
    n = int(state.pop('nlevels'))
    data = np.empty(n, dtype=object)
    for j in range(n):
        data[j] = state.pop('data%d' % j)
    state['data'] = data
    rq = rs.RunningQuantile(state=state)
    assert rq.size() == 6000
    np.testing.assert_allclose(rq.quantiles(QS).numpy(), gold['big_quantiles'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rq.percentiles(QS).numpy(), gold['big_percentiles'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rq.normalize(PROBES).numpy(), gold['big_normalize'], rtol=1e-6, atol=1e-7)


def test_exact_quantiles_agree_with_reference_sketch_within_its_resolution(gold, x, tmp_path):
    rq = rs.RunningQuantile(r=1024)
    for i in range(0, 6000, 500):
        rq.add(x[i:i + 500])
    exact = torch.quantile(x.double(), torch.tensor(QS, dtype=torch.float64), dim=0).t()
    mine = rq.quantiles(QS).double()
    # mine vs torch.quantile: same samples, slightly different plotting position (centre of the
    # weight interval vs (n-1) q): identical up to one sample spacing
    ranks_mine = torch.stack([(x[:, d:d + 1] <= mine[d][None, :]).double().mean(0) for d in range(3)])
    assert (ranks_mine - torch.tensor(QS)[None, :].double()).abs().max() < 1e-3
    assert torch.isfinite(exact).all()
    # the reference's randomised sketch estimates the same quantiles: compare in rank space
    ref_q = torch.from_numpy(gold['big_quantiles']).double()
    ranks_ref = torch.stack([(x[:, d:d + 1] <= ref_q[d][None, :]).double().mean(0) for d in range(3)])
    assert (ranks_ref - torch.tensor(QS)[None, :].double()).abs().max() < 5e-3      # r = 1024 sketch
    assert (ranks_ref - ranks_mine).abs().max() < 5e-3
    # the state written here is a one-level sketch in the reference's format
    sd = rq.state_dict()
    assert set(sd) == {'constructor', 'resolution', 'depth', 'buffersize', 'samplerate', 'data',
                       'sizes', 'extremes', 'size', 'batchcount'}
    assert sd['data'][0].shape == (6000, 3) and sd['sizes'][0] >= 6000
    np.savez(tmp_path / 'rq.npz', **sd)
    back = rs.RunningQuantile(state=str(tmp_path / 'rq.npz'))
    assert torch.equal(back.quantiles(QS), rq.quantiles(QS))


def test_tally_topk_and_quantile_driver_and_cache(x, tmp_path):
    calls = []

    def compute(batch):
        calls.append(len(batch))
        return batch.max(dim=1)[0], batch.reshape(-1)[:, None]
    cache = str(tmp_path / 'tq.npz')
    rtk, rq = tally.tally_topk_and_quantile(compute, x[:1000], k=4, batch_size=100, cachefile=cache)
    assert len(calls) == 10 and rtk.size() == 1000 and rq.size() == 3000
    want = x[:1000].max(dim=1)[0].topk(4)
    assert torch.equal(rtk.result()[0], want[0]) and torch.equal(rtk.result()[1], want[1])
    del calls[:]
    rtk2, rq2 = tally.tally_topk_and_quantile(compute, x[:1000], k=4, batch_size=100, cachefile=cache)
    assert not calls                                            # served from the cache
    assert torch.equal(rtk2.result()[1], want[1])
    assert torch.equal(rq2.quantiles([0.5]), rq.quantiles([0.5]))
    rq3 = tally.tally_quantile(lambda b: b, x[:500], batch_size=50)
    assert torch.allclose(rq3.median(), x[:500].median(dim=0)[0], atol=0.05)


def test_running_quantile_bounded_memory_compaction():
    """above `max_retained` values per unit a level is halved like the reference's
    sketch (sorted, every second value promoted with twice the weight): memory stays bounded and
    the quantiles stay within the compaction's rank error of the exact ones."""
    import torch
    from rewriting_b200.utils import runningstats
    g = torch.Generator().manual_seed(3)
    data = torch.randn(40000, 6, generator=g) * torch.tensor([1., 2., 0.5, 3., 1., 10.]) + 0.3
    exact = runningstats.RunningQuantile(max_retained=0)
    small = runningstats.RunningQuantile(max_retained=1024)
    for i in range(0, 40000, 500):
        exact.add(data[i:i + 500])
        small.add(data[i:i + 500])
    kept = sum(v.shape[1] for v, _ in small._levels())
    assert kept <= 6 * 1024 and small.size() == 40000 and len(small._upper) >= 4   # ~1 K per level
    qs = torch.tensor([0.001, 0.01, 0.1, 0.5, 0.9, 0.99, 0.999])
    want = exact.quantiles(qs)
    got = small.quantiles(qs)
    # compare in rank space: where does the estimate fall in the exact distribution?
    rank = exact.normalize(got)
    assert (rank - qs[None, :]).abs().max().item() < 0.01
    assert (got[:, 3] - want[:, 3]).abs().max().item() < 0.05 * 10
    assert abs(small.mean()[5].item() - data[:, 5].mean().item()) < 0.2
    # the bounded sketch round-trips through its state dict (plain arrays per level)
    again = runningstats.RunningQuantile(state=small.state_dict())
    assert torch.allclose(again.quantiles(qs), got)
