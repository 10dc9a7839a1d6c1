"""tcgen05 pieces of the PPO update: tile blobs and the weight-gradient GEMMs, against torch on the same bf16 values."""
import pytest
import torch

from tests.common import assert_close

pytestmark = pytest.mark.gpu

DIMS = dict(a0=64, a1=256, a2=128, a3=64, dz1=256, dz2=128, dz3=64, dh=16)


def _blob(lib, x):
    from locomanipulationrl_b200 import _lib
    rows, d = x.shape
    buf = torch.empty(lib.lm_mlp_blob_bytes(rows, d), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.lm_mlp_blob_pack(x.contiguous().data_ptr(), buf.data_ptr(), rows, d, s), "pack")
    return buf


@pytest.mark.parametrize("rows,d", [(1, 8), (128, 64), (300, 256), (1000, 16)])
def test_blob_round_trip(rows, d):
    from locomanipulationrl_b200 import _lib
    lib = _lib.load()
    x = torch.randn(rows, d, device="cuda")
    out = torch.empty_like(x)
    b = _blob(lib, x)
    _lib.check(lib.lm_mlp_blob_unpack(b.data_ptr(), out.data_ptr(), rows, d, torch.cuda.current_stream().cuda_stream), "unpack")
    assert torch.equal(out, x.bfloat16().float())


@pytest.mark.parametrize("rows", [128, 130, 128 * 7, 128 * 300 + 5])
def test_weight_gradients_match_torch(rows):
    """dW = dZ^T A and db = dZ^T 1 from the MN-major re-read of the K-major blobs vs fp32 torch on the bf16-rounded values."""
    from locomanipulationrl_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(rows)
    t = {k: torch.randn(rows, d, device="cuda", generator=g) * (0.5 if k.startswith("a") else 0.05) for k, d in DIMS.items()}
    t["dh"][:, 13:] = 0.0
    blobs = {k: _blob(lib, v) for k, v in t.items()}
    r = {k: v.bfloat16().float() for k, v in t.items()}
    shapes = dict(w1=(256, 64), b1=(256,), w2=(128, 256), b2=(128,), w3=(64, 128), b3=(64,), w_mu=(12, 64), b_mu=(12,), w_value=(1, 64), b_value=(1,))
    out = {k: torch.full(s, float("nan"), device="cuda") for k, s in shapes.items()}
    ctas = lib.lm_device_sm_count(0)
    ws = torch.empty(ctas * lib.lm_mlp_wgrad_workspace_floats(), device="cuda")
    rc = lib.lm_mlp_wgrad(*[blobs[k].data_ptr() for k in ("dz1", "dz2", "dz3", "dh", "a0", "a1", "a2", "a3")], rows, ws.data_ptr(), ctas,
                          *[out[k].data_ptr() for k in ("w1", "b1", "w2", "b2", "w3", "b3", "w_mu", "b_mu", "w_value", "b_value")],
                          torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "lm_mlp_wgrad")
    ref = dict(w1=r["dz1"].t() @ r["a0"], b1=r["dz1"].sum(0), w2=r["dz2"].t() @ r["a1"], b2=r["dz2"].sum(0), w3=r["dz3"].t() @ r["a2"], b3=r["dz3"].sum(0),
               w_mu=r["dh"][:, :12].t() @ r["a3"], b_mu=r["dh"][:, :12].sum(0), w_value=r["dh"][:, 12:13].t() @ r["a3"], b_value=r["dh"][:, 12].sum().view(1))
    for k in shapes:
        scale = float(ref[k].abs().max()) + 1e-6
        assert_close(k, out[k], ref[k], rtol=1e-4, atol=2e-5 * scale + 1e-6)


@pytest.mark.parametrize("B", [128, 1000, 32768 + 17])
def test_ppo_gradients_match_autograd(B):
    """Training forward + loss head + backward + weight gradients on tcgen05 vs torch autograd of the fp32 model on the restated loss.
    bf16 operands (activations, pre-activation gradients, weights): tolerance 3 % of each tensor's largest gradient."""
    from locomanipulationrl_b200.policy.mlp import MlpActorCritic
    from locomanipulationrl_b200.ppo import MlpPpoGradients
    from oracle import port_policy as pp
    from tests.test_ppo import _batch
    torch.manual_seed(B)
    model = MlpActorCritic(device="cuda:0")
    with torch.no_grad():
        model.log_std_parameter.copy_(torch.randn(12) * 0.1)
    obs = torch.randn(B, 64, device="cuda").clamp(-5, 5)
    _, _, _, actions, old_nlp, adv, old_v, ret, old_mu, old_ls = (t.cuda() for t in _batch(B, B + 1, wide=True))
    cfg = dict(e_clip=0.2, critic_coef=2.0, entropy_coef=0.0, bounds_loss_coef=1e-4, clip_value=True)
    # reference: fp32 autograd
    mu, value, log_std = model.reference_forward(obs)
    loss, _ = pp.ppo_loss(mu, value.view(-1), log_std, actions, old_nlp, adv, old_v, ret, old_mu, old_ls, **cfg)
    loss.backward()
    ref = {k: p.grad.clone() for k, p in model.named_parameters()}
    for p in model.parameters():
        p.grad = None
    stats = MlpPpoGradients(model, max_batch=B, **cfg)(obs, actions, old_nlp, adv, old_v, ret, old_mu, old_ls)
    assert abs(float(stats[5]) - float(loss.detach())) <= 2e-2 * abs(float(loss.detach())) + 1e-3
    for k, p in model.named_parameters():
        scale = float(ref[k].abs().max()) + 1e-12
        err = float((p.grad - ref[k]).abs().max())
        assert err <= 3e-2 * scale, f"{k}: max err {err:.3e} vs scale {scale:.3e}"
        cos = torch.nn.functional.cosine_similarity(p.grad.flatten(), ref[k].flatten(), dim=0)
        assert float(cos) > 0.999, f"{k}: cosine {float(cos):.5f}"


def test_whole_update_step_with_normalisation_and_adam():
    """Observation normalisation in the training forward, gradients on tcgen05, fused clip + Adam on the flat arena: the parameter
    step points the same way as the all-torch fp32 update (Adam divides by |g|, so compare the step direction, not its entries)."""
    from locomanipulationrl_b200.policy.mlp import MlpActorCritic
    from locomanipulationrl_b200.ppo import FlatAdam, MlpPpoGradients
    from oracle import port_policy as pp
    from tests.test_ppo import _batch
    B = 16384
    torch.manual_seed(5)
    model, ref_model = MlpActorCritic(device="cuda:0"), MlpActorCritic(device="cuda:0")
    ref_model.load_state_dict(model.state_dict())
    obs = torch.randn(B, 64, device="cuda") * 3 + 0.5
    mean, inv_std = torch.randn(64, device="cuda") * 0.3, 1.0 / torch.sqrt(torch.rand(64, device="cuda") + 0.5)
    _, _, _, actions, old_nlp, adv, old_v, ret, old_mu, old_ls = (t.cuda() for t in _batch(B, 77, wide=True))
    cfg = dict(e_clip=0.2, critic_coef=2.0, entropy_coef=0.0, bounds_loss_coef=1e-4, clip_value=True)
    # all-torch reference
    ref_params = list(ref_model.parameters())
    before = torch.cat([p.detach().reshape(-1) for p in ref_params]).clone()
    opt = torch.optim.Adam(ref_params, lr=3e-4, eps=1e-8)
    mu, value, log_std = ref_model.reference_forward(torch.clamp((obs - mean) * inv_std, -5.0, 5.0))
    loss, _ = pp.ppo_loss(mu, value.view(-1), log_std, actions, old_nlp, adv, old_v, ret, old_mu, old_ls, **cfg)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
    opt.step()
    step_ref = torch.cat([p.detach().reshape(-1) for p in ref_params]) - before
    # ours: parameters and gradients as views of two flat arenas
    params = list(model.parameters())
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    flat_grad = torch.zeros_like(flat)
    off = 0
    for p in params:
        k = p.numel()
        p.data = flat[off:off + k].view_as(p)
        p.grad = flat_grad[off:off + k].view_as(p)
        off += k
    MlpPpoGradients(model, max_batch=B, **cfg)(obs, actions, old_nlp, adv, old_v, ret, old_mu, old_ls, norm_mean=mean, norm_inv_std=inv_std)
    g_ref = torch.cat([p.grad.reshape(-1) for p in ref_params])          # clipped reference gradient
    opt_ours = FlatAdam(flat, lr=3e-4, eps=1e-8, max_grad_norm=1.0)
    opt_ours.step(flat_grad)
    cos_g = torch.nn.functional.cosine_similarity(flat_grad, g_ref, dim=0)
    assert float(cos_g) > 0.999, float(cos_g)
    step = flat - before
    big = g_ref.abs() > 0.05 * g_ref.abs().max()                         # entries whose sign is not decided by bf16 noise
    assert torch.equal(torch.sign(step[big]), torch.sign(step_ref[big]))
    assert float((step - step_ref).abs().max()) <= 2.05 * 3e-4


def test_ppo_updater_epoch_tracks_the_all_torch_loop():
    """PpoUpdater (GAE, advantage normalisation, contiguous minibatches, tcgen05 gradients, clip + Adam) against the same loop in
    stock torch on a copy of the model: after one mini-epoch the parameter steps point the same way and the value loss went down."""
    from locomanipulationrl_b200.policy.mlp import MlpActorCritic
    from locomanipulationrl_b200.ppo import PpoUpdater
    from locomanipulationrl_b200.rollout import RolloutBuffer
    from oracle import port_policy as pp
    T, N, MB = 8, 2048, 4096
    torch.manual_seed(11)
    model, ref_model = MlpActorCritic(device="cuda:0"), MlpActorCritic(device="cuda:0")
    ref_model.load_state_dict(model.state_dict())
    buf = RolloutBuffer(T, N, reward_scale=1.0)
    g = torch.Generator(device="cuda").manual_seed(3)
    log_std = model.log_std_parameter.detach().clone()
    with torch.no_grad():
        for t in range(T):                                   # a synthetic rollout under the current policy
            obs = torch.randn(N, 64, device="cuda", generator=g).clamp(-5, 5)
            mu, value, _ = model.reference_forward(obs)
            buf.write(t, obs, mu.contiguous(), value.contiguous(), log_std, torch.randn(N, device="cuda", generator=g),
                      (torch.rand(N, device="cuda", generator=g) < 0.02).long(), noise=torch.randn(N, 12, device="cuda", generator=g))
    last_v = torch.randn(N, device="cuda", generator=g)
    last_d = torch.zeros(N, device="cuda")
    td = {k: v.clone() for k, v in buf.tensor_dict.items()}
    before = torch.cat([p.detach().reshape(-1) for p in ref_model.parameters()]).clone()
    # ---- all-torch loop
    adv, ret = pp.gae_rlgames(td["rewards"].cpu(), td["values"].cpu(), td["dones"].cpu(), last_v.cpu(), last_d.cpu())
    adv, ret = adv.cuda(), ret.cuda()
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    fl = lambda x: x.reshape(T * N, *x.shape[2:])
    ref_params = list(ref_model.parameters())
    opt = torch.optim.Adam(ref_params, lr=3e-4, eps=1e-8)
    c_first = c_last = None
    for i in range(T * N // MB):
        sl = slice(i * MB, (i + 1) * MB)
        mu, value, ls = ref_model.reference_forward(fl(td["obses"])[sl])
        loss, info = pp.ppo_loss(mu, value.view(-1), ls, fl(td["actions"])[sl], fl(td["neglogpacs"])[sl], fl(adv)[sl], fl(td["values"])[sl], fl(ret)[sl],
                                 fl(td["mus"])[sl], log_std)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
        opt.step()
    step_ref = torch.cat([p.detach().reshape(-1) for p in ref_params]) - before
    # ---- ours
    up = PpoUpdater(model, buf, minibatch_size=MB, mini_epochs=1, lr=3e-4, kl_threshold=None)
    stats = up.update(last_v, last_d)
    step = up.flat - before
    cos = torch.nn.functional.cosine_similarity(step, step_ref, dim=0)
    assert float(cos) > 0.97, float(cos)
    assert torch.isfinite(stats).all()
    assert_close("GAE returns inside the updater", up.returns, ret, rtol=1e-5, atol=1e-5)
