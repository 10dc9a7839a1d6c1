"""PPO loss head + fused clip/Adam step: the oracle's gradients are torch autograd on the rl_games restatement;
the optimiser is checked against torch.optim.Adam + clip_grad_norm_ themselves."""
import pytest
import torch

from oracle import port_policy as pp
from tests.common import assert_close


def _batch(B, seed, wide=False):
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(B, 12, generator=g) * (0.9 if wide else 0.4)          # wide: some |mu| > 1.1 -> bound loss active
    log_std = torch.randn(12, generator=g) * 0.1
    old_mu = mu + torch.randn(B, 12, generator=g) * 0.05
    old_log_std = log_std + torch.randn(12, generator=g) * 0.02
    actions = old_mu + torch.randn(B, 12, generator=g) * old_log_std.exp()
    old_nlp = pp.neglogp(actions, old_mu, old_log_std) + torch.randn(B, generator=g) * (0.3 if wide else 0.05)   # wide: ratios leave the clip range
    adv = torch.randn(B, generator=g)
    old_v = torch.randn(B, generator=g)
    value = old_v + torch.randn(B, generator=g) * (0.3 if wide else 0.1)
    ret = old_v + torch.randn(B, generator=g) * 0.5
    return mu, value, log_std, actions, old_nlp, adv, old_v, ret, old_mu, old_log_std


def test_oracle_ppo_loss_reduces_to_textbook_cases():
    mu, value, log_std, actions, old_nlp, adv, old_v, ret, old_mu, old_ls = _batch(512, 0)
    nlp = pp.neglogp(actions, mu, log_std)
    loss, info = pp.ppo_loss(mu, value, log_std, actions, nlp, adv, old_v, ret, old_mu, old_ls, critic_coef=0.0, bounds_loss_coef=0.0)
    assert torch.allclose(loss, -adv.mean(), atol=1e-6)                    # ratio = 1 everywhere: surrogate = -adv
    assert float(info["clip_frac"]) == 0.0 and float(info["kl"]) >= 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("B,wide,clip_value", [(1, False, True), (1000, True, True), (32768, True, False), (98304 + 5, False, True)])
def test_ppo_loss_gradients_match_autograd(B, wide, clip_value):
    from locomanipulationrl_b200.ppo import PpoLoss
    mu, value, log_std, actions, old_nlp, adv, old_v, ret, old_mu, old_ls = _batch(B, B, wide)
    cfg = dict(e_clip=0.2, critic_coef=2.0, entropy_coef=0.01, bounds_loss_coef=1e-4, clip_value=clip_value)
    mu_r, v_r, ls_r = mu.clone().requires_grad_(), value.clone().requires_grad_(), log_std.clone().requires_grad_()
    loss, info = pp.ppo_loss(mu_r, v_r, ls_r, actions, old_nlp, adv, old_v, ret, old_mu, old_ls, **cfg)
    loss.backward()
    head = PpoLoss(**cfg)
    c = lambda t: t.cuda().contiguous()
    g_mu, g_v, g_ls, stats = head(c(mu), c(value), c(log_std), c(actions), c(old_nlp), c(adv), c(old_v), c(ret), c(old_mu), c(old_ls))
    scale = 1.0 / B
    assert_close("dL/dmu", g_mu, mu_r.grad, rtol=2e-4, atol=2e-6 * scale + 1e-9)
    assert_close("dL/dvalue", g_v, v_r.grad, rtol=2e-4, atol=2e-6 * scale + 1e-9)
    assert_close("dL/dlog_std", g_ls, ls_r.grad, rtol=5e-4, atol=2e-6)
    s = stats.cpu()
    for i, k in enumerate(("a_loss", "c_loss", "b_loss", "entropy", "kl")):
        assert abs(float(s[i]) - float(info[k])) <= 2e-5 + 2e-4 * abs(float(info[k])), (k, float(s[i]), float(info[k]))
    assert abs(float(s[5]) - float(loss.detach())) <= 2e-5 + 2e-4 * abs(float(loss.detach()))
    assert abs(float(s[6]) - float(info["clip_frac"])) <= 2.0 / B + 1e-6                     # a ratio on the clip boundary may flip
    if wide:
        assert float(info["clip_frac"]) > 0.05 and float(info["b_loss"]) > 0.0                # the clipped branches were exercised


@pytest.mark.gpu
@pytest.mark.parametrize("n,max_norm,wd", [(58637, 1.0, 0.0), (10178, 0.0, 0.0), (1000, 0.5, 0.01)])
def test_adam_step_matches_torch(n, max_norm, wd):
    from locomanipulationrl_b200.ppo import FlatAdam
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    params = p0.clone().cuda()
    ours = FlatAdam(params, lr=3e-4, eps=1e-8, weight_decay=wd, max_grad_norm=max_norm)
    for step in range(5):
        grad = torch.randn(n, generator=g) * (3.0 if step % 2 else 0.01)
        ref.grad = (grad * 0.5).clone()                           # grad_scale = 0.5: "two ranks summed"
        if max_norm > 0:
            total = torch.nn.utils.clip_grad_norm_([ref], max_norm)
        else:
            total = ref.grad.norm()
        opt.step()
        gd = grad.clone().cuda()
        ours.step(gd, grad_scale=0.5)
        assert_close("grad norm", ours.grad_norm, total.view(1), rtol=1e-5, atol=1e-7)
        assert_close("clipped grad", gd, ref.grad, rtol=1e-5, atol=1e-9)
        assert_close("params", params, ref.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_minibatch_update_through_the_mlp():
    """One full minibatch update: torch forward/backward of the MLP + lm_ppo_loss + lm_adam_step vs the all-torch update."""
    from locomanipulationrl_b200.policy.mlp import MlpActorCritic
    from locomanipulationrl_b200.ppo import FlatAdam, PpoLoss
    B = 8192
    torch.manual_seed(3)
    model = MlpActorCritic(device="cuda:0")
    ref_model = MlpActorCritic(device="cuda:0")
    ref_model.load_state_dict(model.state_dict())
    obs = torch.randn(B, 64, device="cuda")
    _, _, _, actions, old_nlp, adv, old_v, ret, old_mu, old_ls = (t.cuda() for t in _batch(B, 9, wide=True))
    cfg = dict(e_clip=0.2, critic_coef=2.0, entropy_coef=0.0, bounds_loss_coef=1e-4, clip_value=True)
    # ---- all-torch reference update
    ref_params = [p for p in ref_model.parameters()]
    opt = torch.optim.Adam(ref_params, lr=3e-4, eps=1e-8)
    mu, value, log_std = ref_model.reference_forward(obs)
    loss, _ = pp.ppo_loss(mu, value.view(-1), log_std, actions, old_nlp, adv, old_v, ret, old_mu, old_ls, **cfg)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
    opt.step()
    # ---- ours: flat arena, loss head kernel, fused clip + Adam
    params = [p for p in model.parameters()]
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    flat_grad = torch.zeros_like(flat)
    off = 0
    for p in params:                                            # parameters and their .grad become views of the two arenas
        k = p.numel()
        p.data = flat[off:off + k].view_as(p)
        p.grad = flat_grad[off:off + k].view_as(p)
        off += k
    mu, value, log_std = model.reference_forward(obs)
    head = PpoLoss(**cfg)
    g_mu, g_v, g_ls, _ = head(mu.detach().contiguous(), value.detach().view(-1).contiguous(), log_std.detach().contiguous(), actions, old_nlp,
                              adv, old_v, ret, old_mu, old_ls)
    torch.autograd.backward([mu, value.view(-1), log_std], [g_mu, g_v, g_ls])
    FlatAdam(flat, lr=3e-4, eps=1e-8, max_grad_norm=1.0).step(flat_grad)
    for (name, p), q in zip(model.named_parameters(), ref_model.parameters()):
        assert_close(f"updated {name}", p, q, rtol=1e-4, atol=2e-6)


@pytest.mark.gpu
def test_minibatch_update_through_the_gnn():
    """The GNN policy trains through its torch forward (autograd) + lm_ppo_loss + lm_adam_step; the same update all in torch."""
    from locomanipulationrl_b200.policy.gnn import GnnActorCritic
    from locomanipulationrl_b200.ppo import FlatAdam, PpoLoss
    B = 4096
    torch.manual_seed(4)
    model, ref_model = GnnActorCritic(device="cuda:0"), GnnActorCritic(device="cuda:0")
    ref_model.load_state_dict(model.state_dict())
    obs = torch.randn(B, 64, device="cuda")
    _, _, _, actions, old_nlp, adv, old_v, ret, old_mu, old_ls = (t.cuda() for t in _batch(B, 21, wide=True))
    cfg = dict(e_clip=0.2, critic_coef=2.0, entropy_coef=0.0, bounds_loss_coef=1e-4, clip_value=True)
    ref_params = list(ref_model.parameters())
    opt = torch.optim.Adam(ref_params, lr=3e-4, eps=1e-8)
    mu, value, log_std = ref_model.reference_forward(obs)
    loss, _ = pp.ppo_loss(mu, value.view(-1), log_std, actions, old_nlp, adv, old_v, ret, old_mu, old_ls, **cfg)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
    opt.step()
    params = list(model.parameters())
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    flat_grad = torch.zeros_like(flat)
    off = 0
    for p in params:
        k = p.numel()
        p.data = flat[off:off + k].view_as(p)
        p.grad = flat_grad[off:off + k].view_as(p)
        off += k
    mu, value, log_std = model.reference_forward(obs)
    g_mu, g_v, g_ls, _ = PpoLoss(**cfg)(mu.detach().contiguous(), value.detach().view(-1).contiguous(), log_std.detach().contiguous(), actions,
                                        old_nlp, adv, old_v, ret, old_mu, old_ls)
    torch.autograd.backward([mu, value.view(-1), log_std], [g_mu, g_v, g_ls])
    FlatAdam(flat, lr=3e-4, eps=1e-8, max_grad_norm=1.0).step(flat_grad)
    for (name, p), q in zip(model.named_parameters(), ref_model.parameters()):
        assert_close(f"updated {name}", p, q, rtol=1e-4, atol=2e-6)
    # and the kernels' inference forward tracks the updated parameters
    mu_k, value_k, _ = GnnActorCritic.forward(model, obs)
    mu_t, value_t, _ = model.reference_forward(obs)
    assert_close("bf16 kernel after the update", mu_k, mu_t.detach(), rtol=0, atol=1e-2)
