"""Rollout writer + running mean/std: oracle identities on CPU, kernel parity on the GPU through the C ABI."""
import math

import pytest
import torch

from oracle import port_policy as pp
from tests.common import assert_close


# ----------------------------------------------------------------------------------------------- oracle (CPU)
def test_oracle_rms_matches_moment_form():
    """Folding one batch into the (mean 0, var 1, count 1) prior has a closed form in the raw moments."""
    torch.manual_seed(0)
    x = (torch.randn(4096, 64) * 3 + 1.5).double()
    s = pp.rms_update(pp.rms_new(64), x.float())
    n = x.shape[0]
    bm, bv = x.mean(0), x.var(0)
    mean = bm * n / (n + 1)
    var = (1.0 + bv * n + bm ** 2 * n / (n + 1)) / (n + 1)
    assert torch.allclose(s["mean"], mean, rtol=1e-6, atol=1e-7) and torch.allclose(s["var"], var, rtol=1e-6) and float(s["count"]) == n + 1
    # folding chunk by chunk converges to the same statistics (not identical: each fold weights an unbiased variance by n)
    c = pp.rms_new(64)
    for chunk in x.float().split(512):
        c = pp.rms_update(c, chunk)
    assert torch.allclose(c["mean"], s["mean"], atol=1e-5) and torch.allclose(c["var"], s["var"], rtol=2e-2)
    y = pp.rms_normalise(s, x.float(), "rl_games")
    assert float(y.abs().max()) <= 5.0 and abs(float(y.mean())) < 0.05
    y2 = pp.rms_normalise(s, x.float(), "skrl")
    assert torch.allclose(y, y2, atol=1e-4)


def test_oracle_neglogp_is_minus_log_prob():
    torch.manual_seed(1)
    mu, ls = torch.randn(257, 12), torch.randn(12) * 0.2
    a = mu + torch.randn(257, 12) * ls.exp()
    ref = -torch.distributions.Normal(mu, ls.exp()).log_prob(a).sum(-1)
    assert torch.allclose(pp.neglogp(a, mu, ls), ref, rtol=1e-5, atol=1e-5)
    assert abs(float(pp.neglogp(mu[:1], mu[:1], torch.zeros(12))) - 6 * math.log(2 * math.pi)) < 1e-5


# ----------------------------------------------------------------------------------------------- kernels (GPU)
def _step_inputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    obs = (torch.randn(n, 64, generator=g) * 2.0 + torch.linspace(-3, 3, 64)).clamp(-5, 5)
    mu = torch.randn(n, 12, generator=g) * 0.5
    value = torch.randn(n, 1, generator=g)
    noise = torch.randn(n, 12, generator=g)
    rew = torch.randn(n, generator=g) * 2
    dones = (torch.rand(n, generator=g) < 0.05).long()
    return obs, mu, value, noise, rew, dones


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 7, 1000, 65536 + 3])
def test_rollout_write_matches_port(n):
    from locomanipulationrl_b200.rollout import RolloutBuffer, RunningMeanStd
    T = 3
    log_std = torch.linspace(-0.5, 0.3, 12)
    buf = RolloutBuffer(T, n, reward_scale=0.01)
    stats = RunningMeanStd(64)
    ref = {k: torch.zeros_like(v, device="cpu") for k, v in buf.tensor_dict.items()}
    ref_stats = pp.rms_new(64)
    actions_out = torch.zeros(n, 12, device="cuda")
    for t in range(T):
        obs, mu, value, noise, rew, dones = _step_inputs(n, 10 * n + t)
        use_noise = t != 1                                     # step 1: deterministic (noise = NULL)
        a_ref = pp.rollout_write(ref, t, obs, mu, value, log_std, noise if use_noise else None, rew, dones, reward_scale=0.01)
        ref_stats = pp.rms_update(ref_stats, obs)
        buf.write(t, obs.cuda(), mu.cuda(), value.cuda(), log_std.cuda(), rew.cuda(), dones.cuda(), noise=noise.cuda() if use_noise else None,
                  actions_out=actions_out, obs_stats=stats)
        assert_close("sampled actions", actions_out, a_ref, rtol=1e-6, atol=1e-7)      # sigma = exp(log_std): <= 1 ulp between libms
    td = buf.tensor_dict
    for k in ("obses", "mus", "values", "rewards", "dones"):
        assert torch.equal(td[k].cpu(), ref[k]), k
    assert_close("actions", td["actions"], ref["actions"], rtol=1e-6, atol=1e-7)
    assert_close("neglogp", td["neglogpacs"], ref["neglogpacs"], rtol=1e-5, atol=1e-5)
    assert_close("running mean", stats.running_mean, ref_stats["mean"], rtol=1e-6, atol=1e-6)
    assert_close("running var", stats.running_var, ref_stats["var"], rtol=2e-6, atol=1e-6)
    assert float(stats.count) == float(ref_stats["count"])
    mean32, inv_std32 = stats.export()
    assert_close("exported mean", mean32, ref_stats["mean"].float(), rtol=1e-6, atol=1e-6)
    assert_close("exported 1/std", inv_std32, 1.0 / torch.sqrt(ref_stats["var"].float() + 1e-5), rtol=2e-6, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,n", [(1, 1), (1, 48 * 8192), (12, 999), (64, 16384), (4, 5)])
def test_rms_update_matches_port(dim, n):
    from locomanipulationrl_b200.rollout import RunningMeanStd, SKRL
    g = torch.Generator().manual_seed(dim * 100 + n)
    stats = RunningMeanStd(dim, variant=SKRL)
    ref = pp.rms_new(dim)
    for k in range(3):
        x = torch.randn(n, dim, generator=g) * (1.0 + k) + 10.0 * k          # moving mean: the shift matters
        ref = pp.rms_update(ref, x)
        stats.update(x.cuda())
    assert_close("running mean", stats.running_mean, ref["mean"], rtol=1e-6, atol=1e-6)
    assert_close("running var", stats.running_var, ref["var"], rtol=5e-6, atol=1e-6)
    assert float(stats.count) == float(ref["count"])
    mean32, inv_std32 = stats.export()
    assert_close("skrl 1/std", inv_std32, 1.0 / (torch.sqrt(ref["var"].float()) + 1e-8), rtol=2e-6, atol=0)


@pytest.mark.gpu
def test_rollout_write_in_kernel_sampling():
    """sample_seed: the kernel draws its own N(0,1) noise.  Check the moments of (a - mu) / sigma, that neglogp is consistent with the
    stored action, reproducibility for a seed, and fresh draws from call to call."""
    from locomanipulationrl_b200.rollout import RolloutBuffer
    n, T = 65536, 2
    obs, mu, value, _, rew, dones = (t.cuda() for t in _step_inputs(n, 1))
    log_std = torch.linspace(-0.5, 0.3, 12).cuda()
    a = torch.zeros(n, 12, device="cuda")

    def run(seed):
        buf = RolloutBuffer(T, n)
        for t in range(T):
            buf.write(t, obs, mu, value, log_std, rew, dones, actions_out=a, sample_seed=seed)
        return buf.tensor_dict

    td = run(7)
    z = (td["actions"] - td["mus"]) / log_std.exp()
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3 and float(z.abs().max()) < 7.0
    assert float(torch.corrcoef(z[0, :, :6].t()).fill_diagonal_(0).abs().max()) < 0.02          # columns independent
    c01 = float((z[0] * z[1]).mean())
    assert abs(c01) < 5e-3, c01                                                                  # fresh draws every call
    ref = pp.neglogp(td["actions"][1].cpu(), td["mus"][1].cpu(), log_std.cpu())
    assert_close("neglogp of the in-kernel sample", td["neglogpacs"][1], ref, rtol=1e-5, atol=1e-5)
    assert torch.equal(run(7)["actions"], td["actions"]) and not torch.equal(run(8)["actions"], td["actions"])
