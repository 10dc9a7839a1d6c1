"""GPU parity tests of the policy-side kernels (MLP tcgen05, GNN, GAE, PD) through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import port_policy
from tests.common import GOLDEN_DIR, assert_close

pytestmark = pytest.mark.gpu


def _mlp(seed=0):
    from locomanipulationrl_b200.policy.mlp import MlpActorCritic
    torch.manual_seed(seed)
    return MlpActorCritic(device="cuda:0")


@pytest.mark.parametrize("n", [1, 100, 128, 16384, 65536 + 37])
def test_mlp_forward_matches_fp32_reference(n):
    """bf16 operands / fp32 accumulate vs stock fp32 torch: logits within 1e-2 (north_star tolerance)."""
    m = _mlp()
    g = torch.Generator().manual_seed(n)
    obs = (torch.randn(n, 64, generator=g) * 1.5).clamp(-5, 5).cuda()
    mu, value, log_std = m(obs)
    with torch.no_grad():
        mu_ref, value_ref, _ = m.reference_forward(obs)
    assert mu.shape == (n, 12) and value.shape == (n, 1) and log_std.shape == (12,)
    e1 = assert_close("mu", mu, mu_ref, rtol=0, atol=1e-2)
    e2 = assert_close("value", value, value_ref, rtol=0, atol=1e-2)
    print(f"n={n} max |mu err| {e1:.2e}, |value err| {e2:.2e}")
    # and against an emulation that rounds operands to bf16 exactly where the kernel does: tight agreement
    with torch.no_grad():
        h = obs.bfloat16().float()
        for i in (0, 2, 4):
            h = torch.nn.functional.elu(h @ m.net[i].weight.bfloat16().float().t() + m.net[i].bias).bfloat16().float()
        mu_e = h @ m.mean_layer.weight.bfloat16().float().t() + m.mean_layer.bias
        v_e = h @ m.value_layer.weight.bfloat16().float().t() + m.value_layer.bias
    assert_close("mu vs bf16 emulation", mu, mu_e, rtol=0, atol=2e-3)
    assert_close("value vs bf16 emulation", value, v_e, rtol=0, atol=2e-3)


@pytest.mark.parametrize("n", [100, 16384, 65536 + 37])
def test_mlp_inference_kernel_matches_fp32_reference_and_its_bf16_emulation(n):
    """The inference instantiation (torch.no_grad: hidden activations stored as bf16(log2(e) h), biases pre-scaled, heads un-scaled in
    their epilogue): logits within 1e-2 of stock fp32 torch, within 2e-3 of an emulation that rounds to bf16 exactly where the kernel
    does, and within 4e-3 of the training forward (same weights, the plain formulation)."""
    m = _mlp()
    g = torch.Generator().manual_seed(7 * n + 1)
    obs = (torch.randn(n, 64, generator=g) * 1.5).clamp(-5, 5).cuda()
    s = 1.4426950408889634
    with torch.no_grad():
        mu, value, _ = m(obs)
        mu_ref, value_ref, _ = m.reference_forward(obs)
        h = (obs * s).bfloat16().float()
        for i in (0, 2, 4):
            t = h @ m.net[i].weight.bfloat16().float().t() + s * m.net[i].bias
            h = (s * torch.nn.functional.elu(t / s)).bfloat16().float()
        mu_e = (h @ m.mean_layer.weight.bfloat16().float().t()) / s + m.mean_layer.bias
        v_e = (h @ m.value_layer.weight.bfloat16().float().t()) / s + m.value_layer.bias
    mu_t, value_t, _ = m(obs)                       # grad enabled: the training forward
    assert_close("mu", mu, mu_ref, rtol=0, atol=1e-2)
    assert_close("value", value, value_ref, rtol=0, atol=1e-2)
    assert_close("mu vs scaled bf16 emulation", mu, mu_e, rtol=0, atol=2e-3)
    assert_close("value vs scaled bf16 emulation", value, v_e, rtol=0, atol=2e-3)
    assert_close("mu inference vs training forward", mu, mu_t.detach(), rtol=0, atol=4e-3)
    assert_close("value inference vs training forward", value, value_t.detach(), rtol=0, atol=4e-3)


def test_mlp_normalisation_and_repack():
    m = _mlp(1)
    obs = torch.randn(4096, 64, device="cuda") * 3 + 0.5
    mean = torch.randn(64, device="cuda") * 0.3
    var = torch.rand(64, device="cuda") + 0.5
    inv_std = 1.0 / torch.sqrt(var + 1e-5)              # rl_games RunningMeanStd form
    mu, value, _ = m(obs, norm_mean=mean, norm_inv_std=inv_std, norm_clip=5.0)
    x = torch.clamp((obs - mean) * inv_std, -5.0, 5.0)
    with torch.no_grad():
        mu_ref, value_ref, _ = m.reference_forward(x)
    assert_close("mu (normalised)", mu, mu_ref, rtol=0, atol=1e-2)
    assert_close("value (normalised)", value, value_ref, rtol=0, atol=1e-2)
    with torch.no_grad():
        m.mean_layer.bias.add_(1.0)                      # parameter update must trigger a re-pack
    mu2, _, _ = m(obs, norm_mean=mean, norm_inv_std=inv_std)
    assert_close("mu after update", mu2, mu_ref + 1.0, rtol=0, atol=1e-2)


def _gnn_from_golden(precision="fp32"):
    from locomanipulationrl_b200.policy.gnn import GnnActorCritic
    g = np.load(os.path.join(GOLDEN_DIR, "gnn_reference.npz"))
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p/")}
    m = GnnActorCritic(device="cuda:0", precision=precision)
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and set(missing.missing_keys) <= {"log_std_parameter"}
    return m, g, sd


def test_gnn_matches_reference_golden():
    m, g, _ = _gnn_from_golden()
    mu, value, _ = m(torch.from_numpy(g["obs"]).cuda())
    assert_close("gnn mu vs reference", mu, g["mu"], rtol=1e-5, atol=2e-6)
    assert_close("gnn value vs reference", value, g["value"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("n", [1, 39, 40, 41, 8192])
def test_gnn_matches_port(n):
    m, _, sd = _gnn_from_golden()
    obs = torch.randn(n, 64, generator=torch.Generator().manual_seed(n)) * 1.5
    mu, value, _ = m(obs.cuda())
    mu_ref, value_ref = port_policy.gnn_forward(sd, obs)
    assert_close("gnn mu", mu, mu_ref, rtol=1e-5, atol=2e-6)
    assert_close("gnn value", value, value_ref, rtol=1e-5, atol=2e-6)


def test_gnn_bf16_matches_reference_golden():
    """tcgen05 path vs the reference GraphNet's own outputs: bf16 policy logits within 1e-2 (north_star)."""
    m, g, _ = _gnn_from_golden("bf16")
    mu, value, _ = m(torch.from_numpy(g["obs"]).cuda())
    assert_close("gnn bf16 mu vs reference", mu, g["mu"], rtol=0, atol=1e-2)
    assert_close("gnn bf16 value vs reference", value, g["value"], rtol=0, atol=1e-2)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 8192, 65536 + 37])
def test_gnn_bf16_matches_port(n):
    if n % 2:
        m, _, sd = _gnn_from_golden("bf16")
    else:                                                  # fresh nn.Linear initialisation, scaled up: larger activations
        from locomanipulationrl_b200.policy.gnn import GnnActorCritic
        torch.manual_seed(n)
        m = GnnActorCritic(device="cuda:0", precision="bf16")
        with torch.no_grad():
            for prm in m.parameters():
                prm.mul_(1.25)
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    obs = torch.randn(n, 64, generator=torch.Generator().manual_seed(n)) * 1.5
    mu, value, _ = m(obs.cuda())
    mu_ref, value_ref = port_policy.gnn_forward(sd, obs)
    e1 = assert_close("gnn bf16 mu", mu, mu_ref, rtol=0, atol=1e-2)
    e2 = assert_close("gnn bf16 value", value, value_ref, rtol=0, atol=1e-2)
    print(f"n={n} max |mu err| {e1:.2e}, |value err| {e2:.2e}")
    # against an emulation that rounds weights and node features to bf16 exactly where the kernel does: tight agreement
    mu_e, value_e = port_policy.gnn_forward(sd, obs, emulate_bf16=True)
    assert_close("gnn mu vs bf16 emulation", mu, mu_e, rtol=0, atol=2e-3)
    assert_close("gnn value vs bf16 emulation", value, value_e, rtol=0, atol=2e-3)
    # fp32 and bf16 kernels share one handle / one pack
    m.precision = "fp32"
    mu32, value32, _ = m(obs.cuda())
    assert_close("gnn fp32 (same handle) mu", mu32, mu_ref, rtol=1e-5, atol=2e-6)
    assert_close("gnn fp32 (same handle) value", value32, value_ref, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("T,N", [(16, 65536), (48, 8192), (5, 333), (1, 7)])
def test_gae_variants(T, N):
    from locomanipulationrl_b200 import gae
    g = torch.Generator().manual_seed(T * 1000 + N)
    r = torch.randn(T, N, generator=g)
    v = torch.randn(T, N, generator=g)
    d = (torch.rand(T, N, generator=g) < 0.05).float()
    lv = torch.randn(N, generator=g)
    ld = (torch.rand(N, generator=g) < 0.05).float()
    a_ref, ret_ref = port_policy.gae_rlgames(r, v, d, lv, ld)
    a, ret = gae.compute_gae(r.cuda(), v.cuda(), d.cuda(), lv.cuda(), ld.cuda(), variant=gae.RLGAMES)
    assert_close("rl_games adv", a, a_ref, rtol=1e-5, atol=1e-6)
    assert_close("rl_games returns", ret, ret_ref, rtol=1e-5, atol=1e-6)
    s_ref, sret_ref = port_policy.gae_skrl(r, v, d.bool(), lv)
    s, sret = gae.compute_gae(r.cuda(), v.cuda(), d.cuda(), lv.cuda(), variant=gae.SKRL)
    assert_close("skrl adv", s, s_ref, rtol=1e-5, atol=1e-6)
    assert_close("skrl returns", sret, sret_ref, rtol=1e-5, atol=1e-6)
    # env-major warp-scan variant: same recurrence, different association -> round-off level agreement
    a2, ret2 = gae.compute_gae(r.t().contiguous().cuda(), v.t().contiguous().cuda(), d.t().contiguous().cuda(), lv.cuda(), ld.cuda(),
                               variant=gae.RLGAMES, env_major=True)
    assert_close("env-major adv", a2.t(), a_ref, rtol=1e-4, atol=1e-5)
    assert_close("env-major ret", ret2.t(), ret_ref, rtol=1e-4, atol=1e-5)


def test_pd_kernels():
    import ctypes as C
    from locomanipulationrl_b200 import _lib
    from locomanipulationrl_b200.configs import TaskConfig
    lib = _lib.load()
    cfg = TaskConfig(name="x", kind=0, vertical=False)
    n = 4097
    g = torch.Generator().manual_seed(1)
    q = torch.randn(n, 20, generator=g); qd = torch.randn(n, 20, generator=g) * 3
    tg = torch.randn(n, 12, generator=g) * 2
    tsum = torch.randn(n, 12, generator=g)
    ref = port_policy.pd_torque(tg, q[:, :12], qd[:, :12])
    qc, qdc, tgc, tsc = q.cuda(), qd.cuda(), tg.cuda(), tsum.cuda()
    tau = torch.empty(n, 12, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.lm_pd_substep(qc.data_ptr(), qdc.data_ptr(), tgc.data_ptr(), tau.data_ptr(), tsc.data_ptr(), None, n, 4.5, 0.2, 1.5, s))
    assert torch.equal(tau.cpu(), ref)
    assert torch.equal(tsc.cpu(), tsum + ref)
    a = torch.rand(n, 12, generator=g) * 2 - 1
    se = torch.tensor(cfg.pd_init_se).repeat(n, 1) + torch.randn(n, 12, generator=g) * 0.3
    lo, hi = torch.tensor(cfg.pd_min_se), torch.tensor(cfg.pd_max_se)
    for zero in (True, False):
        se_ref, tg_ref = port_policy.pd_integrate(a, se, lo, hi, 0.1, zero)
        sec, out = se.clone().cuda(), torch.empty(n, 12, device="cuda")
        lo_c = (C.c_float * 12)(*cfg.pd_min_se); hi_c = (C.c_float * 12)(*cfg.pd_max_se)
        _lib.check(lib.lm_pd_integrate_targets(a.cuda().data_ptr(), sec.data_ptr(), out.data_ptr(), lo_c, hi_c, n, 0.1, int(zero), s))
        assert torch.equal(sec.cpu(), se_ref) and torch.equal(out.cpu(), tg_ref)
