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
@pytest.mark.parametrize("n,full", [(1, False), (7, True), (1000, False), (65536 + 3, True)])
def test_rollout_write_matches_port(n, full):
    """rl_games play_steps order: slot t gets obs / dones / policy outputs, the reward that came with obs t completes slot t-1,
    finish() closes slot T-1 and hands out GAE's bootstrap.  ``full``: value_bootstrap (time_outs) and normalize_value switched on."""
    from locomanipulationrl_b200.rollout import RolloutBuffer, RunningMeanStd
    T = 3
    log_std = torch.linspace(-0.5, 0.3, 12)
    buf = RolloutBuffer(T, n, reward_scale=0.01)
    stats = RunningMeanStd(64)
    ref = {k: torch.zeros_like(v, device="cpu") for k, v in buf.tensor_dict.items()}
    ref_stats = pp.rms_new(64)
    actions_out = torch.zeros(n, 12, device="cuda")
    vrms, vrms_ref, denorm = None, None, None
    if full:
        vrms = RunningMeanStd(1)
        x = torch.randn(max(n, 64), 1) * 3 + 2
        vrms.update(x.cuda())
        vrms_ref = pp.rms_update(pp.rms_new(1), x)
        denorm = vrms.export_denorm()
        assert_close("denorm export", denorm, torch.stack([vrms_ref["mean"].float()[0], torch.sqrt(vrms_ref["var"].float() + 1e-5)[0]]), rtol=1e-6, atol=1e-7)
    for t in range(T + 1):
        obs, mu, value, noise, rew, dones = _step_inputs(n, 10 * n + t)
        value = value * 4                                       # beyond the +-5 clamp of the de-normaliser now and then
        touts = (torch.rand(n, generator=torch.Generator().manual_seed(t)) < 0.3).long() if full else None
        if t == T:
            lv_ref, ld_ref = pp.rollout_finish(ref, rew, value, dones, reward_scale=0.01, time_outs=touts, gamma=0.97, value_rms=vrms_ref)
            lv, ld = buf.finish(rew.cuda(), value.cuda(), dones.cuda(), time_outs=touts.cuda() if full else None, gamma=0.97, value_denorm=denorm)
            assert_close("last values", lv, lv_ref, rtol=1e-6, atol=1e-6)
            assert torch.equal(ld.cpu(), ld_ref)
            break
        use_noise = t != 1                                     # step 1: deterministic (noise = NULL)
        a_ref = pp.rollout_write(ref, t, obs, mu, value, log_std, noise if use_noise else None, rew, dones, reward_scale=0.01,
                                 time_outs=touts, gamma=0.97, value_rms=vrms_ref)
        ref_stats = pp.rms_update(ref_stats, obs)
        buf.write(t, obs.cuda(), mu.cuda(), value.cuda(), log_std.cuda(), rew.cuda(), dones.cuda(), noise=noise.cuda() if use_noise else None,
                  actions_out=actions_out, obs_stats=stats, time_outs=touts.cuda() if full else None, gamma=0.97, value_denorm=denorm)
        assert_close("sampled actions", actions_out, a_ref, rtol=1e-6, atol=1e-7)      # sigma = exp(log_std): <= 1 ulp between libms
    td = buf.tensor_dict
    for k in ("obses", "mus", "dones"):
        assert torch.equal(td[k].cpu(), ref[k]), k
    if full:
        assert_close("values", td["values"], ref["values"], rtol=1e-6, atol=1e-6)
        assert_close("rewards", td["rewards"], ref["rewards"], rtol=1e-6, atol=1e-6)
    else:
        assert torch.equal(td["values"].cpu(), ref["values"]) and torch.equal(td["rewards"].cpu(), ref["rewards"])
    assert float(ref["rewards"].abs().sum()) > 0 or n == 1
    assert_close("actions", td["actions"], ref["actions"], rtol=1e-6, atol=1e-7)
    assert_close("neglogp", td["neglogpacs"], ref["neglogpacs"], rtol=1e-5, atol=1e-5)
    assert_close("running mean", stats.running_mean, ref_stats["mean"], rtol=1e-6, atol=1e-6)
    assert_close("running var", stats.running_var, ref_stats["var"], rtol=2e-6, atol=1e-6)
    assert float(stats.count) == float(ref_stats["count"])
    mean32, inv_std32 = stats.export()
    assert_close("exported mean", mean32, ref_stats["mean"].float(), rtol=1e-6, atol=1e-6)
    assert_close("exported 1/std", inv_std32, 1.0 / torch.sqrt(ref_stats["var"].float() + 1e-5), rtol=2e-6, atol=0)


def test_oracle_rollout_order_matches_play_steps():
    """The restated writer against a literal transcription of rl_games' play_steps loop driven by a toy env."""
    torch.manual_seed(3)
    T, n = 5, 9
    log_std = torch.zeros(12)
    env_rew = [torch.randn(n) for _ in range(T + 1)]          # env_rew[k] = reward returned by the k-th env.step
    env_obs = [torch.randn(n, 64) for _ in range(T + 1)]      # env_obs[0] = reset observation
    env_done = [(torch.rand(n) < 0.2).long() for _ in range(T + 1)]
    pol = lambda o: (o[:, :12] * 0.1, o[:, 12:13])
    lit = dict(obses=torch.zeros(T, n, 64), rewards=torch.zeros(T, n), values=torch.zeros(T, n), dones=torch.zeros(T, n))
    obs, dones = env_obs[0], env_done[0]
    for k in range(T):                                        # a2c_common.play_steps
        mu, value = pol(obs)
        lit["obses"][k], lit["dones"][k], lit["values"][k] = obs, dones.float(), value.view(-1)
        obs, rewards, dones = env_obs[k + 1], env_rew[k + 1], env_done[k + 1]
        lit["rewards"][k] = rewards * 0.5
    buf = dict(obses=torch.zeros(T, n, 64), actions=torch.zeros(T, n, 12), mus=torch.zeros(T, n, 12), values=torch.zeros(T, n),
               rewards=torch.zeros(T, n), dones=torch.zeros(T, n), neglogpacs=torch.zeros(T, n))
    for t in range(T):                                        # the one-call-per-step form: (obs_t, rew_t, done_t) arrive together
        mu, value = pol(env_obs[t])
        pp.rollout_write(buf, t, env_obs[t], mu, value, log_std, None, env_rew[t], env_done[t], reward_scale=0.5)
    lv, ld = pp.rollout_finish(buf, env_rew[T], pol(env_obs[T])[1], env_done[T], reward_scale=0.5)
    for k in lit:
        assert torch.equal(buf[k], lit[k]), k
    assert torch.equal(lv, pol(env_obs[T])[1].view(-1)) and torch.equal(ld, env_done[T].float())


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
