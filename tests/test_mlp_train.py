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


def _synthetic_rollout(model, buf, seed, dev="cuda:0", value_denorm=None):
    """Fills ``buf`` with a rollout under the current policy in rl_games' play_steps order (T writes + finish)."""
    T, N = buf.horizon, buf.num_envs
    g = torch.Generator(device=dev).manual_seed(seed)
    log_std = model.log_std_parameter.detach().clone()
    with torch.no_grad():
        for t in range(T + 1):
            obs = torch.randn(N, 64, device=dev, generator=g).clamp(-5, 5)
            mu, value, _ = model.reference_forward(obs)
            rew = torch.randn(N, device=dev, generator=g)
            dones = (torch.rand(N, device=dev, generator=g) < 0.02).long()
            if t == T:
                last_v, last_d = buf.finish(rew, value.contiguous(), dones, value_denorm=value_denorm)
            else:
                buf.write(t, obs, mu.contiguous(), value.contiguous(), log_std, rew, dones, noise=torch.randn(N, 12, device=dev, generator=g),
                          value_denorm=value_denorm)
    return last_v.clone(), last_d.clone()


def _torch_train_epoch(ref_model, td, last_v, last_d, T, N, MB, mini_epochs, lr, kl_threshold, normalize_value, value_rms=None):
    """rl_games train_epoch in stock torch + the oracle restatements (fp32): prepare_dataset, mini-epochs, legacy adaptive schedule."""
    from oracle import port_policy as pp
    adv, ret = pp.gae_rlgames(td["rewards"].cpu(), td["values"].cpu(), td["dones"].cpu(), last_v.cpu(), last_d.cpu())
    values = td["values"].cpu()
    if normalize_value:
        value_rms = pp.rms_update(value_rms, values.reshape(-1, 1))
        values = pp.rms_normalise(value_rms, values)
        value_rms = pp.rms_update(value_rms, ret.reshape(-1, 1))
        ret_n = pp.rms_normalise(value_rms, ret)
    else:
        ret_n = ret
    adv = pp.normalize_advantages(adv)
    fl = lambda x: x.reshape(T * N, *x.shape[2:]).cuda()
    obs, actions, nlp = fl(td["obses"]), fl(td["actions"]), fl(td["neglogpacs"])
    mus = fl(td["mus"]).clone()
    advs, vals, rets = fl(adv), fl(values), fl(ret_n)
    nmb = T * N // MB
    old_ls = ref_model.log_std_parameter.detach().clone().expand(nmb, 12).clone()
    ref_params = list(ref_model.parameters())
    opt = torch.optim.Adam(ref_params, lr=lr, eps=1e-8)
    kls, lrs = [], []
    for _ in range(mini_epochs):
        for i in range(nmb):
            sl = slice(i * MB, (i + 1) * MB)
            mu, value, ls = ref_model.reference_forward(obs[sl])
            loss, info = pp.ppo_loss(mu, value.view(-1), ls, actions[sl], nlp[sl], advs[sl], vals[sl], rets[sl], mus[sl], old_ls[i])
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
            opt.step()
            mus[sl] = mu.detach()                                  # dataset.update_mu_sigma
            old_ls[i] = ls.detach()
            kls.append(float(info["kl"]))
            if kl_threshold is not None:
                lr = pp.adaptive_lr(lr, kls[-1], kl_threshold)
                for gr in opt.param_groups:
                    gr["lr"] = lr
            lrs.append(lr)
    return ret, adv, value_rms, kls, lrs


@pytest.mark.parametrize("normalize_value", [False, True])
def test_ppo_updater_tracks_the_all_torch_train_epoch(normalize_value):
    """PpoUpdater with the reference yaml's settings (5 mini-epochs, adaptive legacy schedule at kl_threshold 0.008, update_mu_sigma,
    normalize_value) against rl_games' train_epoch in stock torch on a copy of the model: same learning-rate trajectory, parameter
    steps pointing the same way, and — the stale-weights check — the kernels' forward after the update agrees with the UPDATED
    fp32 masters."""
    from locomanipulationrl_b200.policy.mlp import MlpActorCritic
    from locomanipulationrl_b200.ppo import PpoUpdater
    from oracle import port_policy as pp
    T, N, MB, EPOCHS, LR = 8, 2048, 4096, 5, 3e-3                  # a rate high enough that the schedule has to act
    torch.manual_seed(11)
    model, ref_model = MlpActorCritic(device="cuda:0"), MlpActorCritic(device="cuda:0")
    ref_model.load_state_dict(model.state_dict())
    from locomanipulationrl_b200.rollout import RolloutBuffer
    buf = RolloutBuffer(T, N, reward_scale=1.0)
    up = PpoUpdater(model, buf, minibatch_size=MB, mini_epochs=EPOCHS, lr=LR, normalize_value=normalize_value)
    last_v, last_d = _synthetic_rollout(model, buf, 3, value_denorm=up.value_denorm)
    td = {k: v.clone() for k, v in buf.tensor_dict.items()}
    before = torch.cat([p.detach().reshape(-1) for p in ref_model.parameters()]).clone()
    probe = torch.randn(4096, 64, device="cuda").clamp(-5, 5)
    mu_before = model(probe)[0].clone()
    ret, adv, vrms, kls, lrs = _torch_train_epoch(ref_model, td, last_v, last_d, T, N, MB, EPOCHS, LR, 0.008, normalize_value,
                                                  pp.rms_new(1) if normalize_value else None)
    stats = up.update(last_v, last_d)
    assert torch.isfinite(stats).all()
    # prepare_dataset
    if normalize_value:
        assert_close("normalised returns", up.returns, pp.rms_normalise(vrms, ret), rtol=1e-5, atol=1e-5)
        assert_close("value normaliser mean", up.value_rms.running_mean, vrms["mean"], rtol=1e-6, atol=1e-6)
        assert_close("value normaliser var", up.value_rms.running_var, vrms["var"], rtol=1e-5, atol=1e-6)
        assert_close("value_denorm", up.value_denorm, torch.stack([vrms["mean"].float()[0], torch.sqrt(vrms["var"].float() + 1e-5)[0]]), rtol=1e-5, atol=1e-6)
    else:
        assert_close("GAE returns inside the updater", up.returns, ret, rtol=1e-5, atol=1e-5)
    assert_close("normalised advantages", up.advantages, adv, rtol=1e-4, atol=1e-5)
    # schedule: the decisions are only comparable while no reference KL sits at a threshold
    margin = min(min(abs(k - 0.016) / 0.016, abs(k - 0.004) / 0.004) for k in kls)
    assert len(set(lrs)) > 1, "the test rollout must make the schedule act"
    if margin > 0.15:
        assert abs(up.opt.lr / lrs[-1] - 1.0) < 1e-5, (up.opt.lr, lrs[-1], kls)
    else:
        assert 0.4 < up.opt.lr / lrs[-1] < 2.5, (up.opt.lr, lrs[-1])
    # parameters
    step_ref = torch.cat([p.detach().reshape(-1) for p in ref_model.parameters()]) - before
    step = up.flat - before
    cos = torch.nn.functional.cosine_similarity(step, step_ref, dim=0)
    assert float(cos) > 0.9, float(cos)
    # the acting policy is the updated one: kernel forward (bf16 blobs) vs the fp32 masters the optimiser wrote
    mu_k, v_k, _ = model(probe)
    with torch.no_grad():
        mu_m, v_m, _ = model.reference_forward(probe)
    assert_close("mu after the update", mu_k, mu_m, rtol=0, atol=1e-2)
    assert_close("value after the update", v_k, v_m, rtol=0, atol=1.5e-2)
    moved = float((mu_m - mu_before).abs().max())
    assert moved > 0.05, f"the update must move the policy for this check to mean anything ({moved})"


def test_mlp_forward_is_differentiable_through_the_kernels():
    """rl_games calc_gradients / skrl _update run their own torch loss on the module's outputs: under enable_grad the forward is an
    autograd.Function whose backward is lm_mlp_backward + lm_mlp_wgrad.  Compared with autograd through the fp32 restatement."""
    from locomanipulationrl_b200.policy.adapters import SkrlShared
    torch.manual_seed(5)
    shared = SkrlShared(device="cuda:0")
    model = shared.core
    obs = torch.randn(3000, 64, device="cuda").clamp(-5, 5)
    target_mu, target_v = torch.randn(3000, 12, device="cuda"), torch.randn(3000, 1, device="cuda")

    def loss_of(mu, value, ls):
        return ((mu - target_mu) ** 2).mean() + 0.5 * ((value - target_v) ** 2).mean() + 0.01 * ls.sum()

    mu, ls, _ = shared.compute({"states": obs}, "policy")
    value, _ = shared.compute({"states": obs}, "value")
    assert mu.requires_grad and value.requires_grad
    # two differentiable calls share one saved-activation workspace: take each backward right after its forward
    (mu_g,) = torch.autograd.grad(((mu - target_mu) ** 2).mean(), [model.mean_layer.weight], retain_graph=False, allow_unused=True)
    model.zero_grad()
    mu, value, ls = model(obs)
    loss_of(mu, value, ls).backward()
    got = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.zero_grad()
    mu_r, value_r, ls_r = model.reference_forward(obs)
    loss_of(mu_r, value_r, ls_r).backward()
    for k, p in model.named_parameters():
        ref = p.grad
        scale = float(ref.abs().max()) + 1e-12
        assert float((got[k] - ref).abs().max()) <= 0.04 * scale + 1e-6, (k, float((got[k] - ref).abs().max()), scale)
    g_all = torch.cat([got[k].reshape(-1) for k, _ in model.named_parameters()])
    r_all = torch.cat([p.grad.reshape(-1) for _, p in model.named_parameters()])
    assert float(torch.nn.functional.cosine_similarity(g_all, r_all, dim=0)) > 0.999
    # inference path: stale-cache check — the task rewrites obs_buf in place, the second step must not return the first step's heads
    with torch.no_grad():
        m1, _, _ = shared.compute({"states": obs}, "policy")
        v1, _ = shared.compute({"states": obs}, "value")
        obs.copy_(torch.randn_like(obs).clamp(-5, 5).contiguous())            # torch op: version bump
        m2, _, _ = shared.compute({"states": obs}, "policy")
        assert float((m1 - m2).abs().max()) > 1e-3
        import ctypes
        torch.cuda.synchronize()
        lib = model._lib                                                        # a kernel writes through data_ptr: NO version bump
        lib.lm_clamp_copy(obs.data_ptr(), obs.data_ptr(), obs.numel(), 0.01, torch.cuda.current_stream().cuda_stream)
        m3, _, _ = shared.compute({"states": obs}, "policy")
        assert float((m3 - m2).abs().max()) > 1e-3


# ----------------------------------------------------------------------------------------------- GNN policy
@pytest.mark.parametrize("n", [5, 64, 3000])
def test_gnn_backward_matches_autograd(n):
    """lm_gnn_backward (forward recomputed in fp32, limb graph walked backwards, shared-memory weight-gradient accumulation) against
    torch autograd through GnnActorCritic.reference_forward for an arbitrary upstream gradient."""
    import ctypes as C
    from locomanipulationrl_b200 import _lib
    from locomanipulationrl_b200.policy.gnn import GnnActorCritic
    torch.manual_seed(7)
    model = GnnActorCritic(device="cuda:0", precision="fp32")
    with torch.no_grad():
        for q in model.parameters():
            q.mul_(1.6)                                            # livelier activations than the default initialisation
    obs = (torch.randn(n, 64, device="cuda") * 1.5).clamp(-5, 5)
    g_mu, g_v = torch.randn(n, 12, device="cuda") / n, torch.randn(n, device="cuda") / n
    mu, value, _ = model.reference_forward(obs)
    (mu * g_mu).sum().add((value.view(-1) * g_v).sum()).backward()
    ref = [q.grad.clone() for q in model.ordered_parameters()]
    model.zero_grad()
    model.pack()
    lib = _lib.load()
    ctas = lib.lm_device_sm_count(0)
    ws = torch.empty(ctas * lib.lm_gnn_grad_workspace_floats(), device="cuda")
    got = [torch.full_like(q, 7.0) for q in model.ordered_parameters()]      # the kernel WRITES the gradients
    arr = (C.c_void_p * 20)(*[t.data_ptr() for t in got])
    _lib.check(lib.lm_gnn_backward(model._handle, obs.data_ptr(), g_mu.data_ptr(), g_v.data_ptr(), n, ws.data_ptr(), ctas, arr, 20,
                                   torch.cuda.current_stream().cuda_stream), "lm_gnn_backward")
    names = [k for k, _ in model.named_parameters() if k != "log_std_parameter"]
    for k, g, r in zip(names, got, ref):
        scale = float(r.abs().max()) + 1e-9
        err = float((g - r).abs().max())
        assert err <= 2e-4 * scale + 1e-7, (k, err, scale)


def test_ppo_updater_gnn_tracks_autograd():
    """PpoUpdater on the GNN policy (kernel forward, loss head, lm_gnn_backward, clip + Adam, re-pack) against the same minibatch loop
    through torch autograd on a copy of the model."""
    from locomanipulationrl_b200.policy.gnn import GnnActorCritic
    from locomanipulationrl_b200.ppo import PpoUpdater
    from locomanipulationrl_b200.rollout import RolloutBuffer
    from oracle import port_policy as pp
    T, N, MB = 4, 1024, 2048
    torch.manual_seed(2)
    model, ref_model = GnnActorCritic(device="cuda:0", precision="fp32"), GnnActorCritic(device="cuda:0", precision="fp32")
    ref_model.load_state_dict(model.state_dict())
    buf = RolloutBuffer(T, N)
    up = PpoUpdater(model, buf, minibatch_size=MB, mini_epochs=2, lr=1e-3, kl_threshold=None)
    last_v, last_d = _synthetic_rollout(model, buf, 5)
    td = {k: v.clone() for k, v in buf.tensor_dict.items()}
    before = torch.cat([p.detach().reshape(-1) for p in ref_model.parameters()]).clone()
    adv, ret = pp.gae_rlgames(td["rewards"].cpu(), td["values"].cpu(), td["dones"].cpu(), last_v.cpu(), last_d.cpu())
    adv = pp.normalize_advantages(adv)
    fl = lambda x: x.reshape(T * N, *x.shape[2:]).cuda()
    mus = fl(td["mus"]).clone()
    ref_params = list(ref_model.parameters())
    opt = torch.optim.Adam(ref_params, lr=1e-3, eps=1e-8)
    old_ls = ref_model.log_std_parameter.detach().clone().expand(T * N // MB, 12).clone()
    for _ in range(2):
        for i in range(T * N // MB):
            sl = slice(i * MB, (i + 1) * MB)
            mu, value, ls = ref_model.reference_forward(fl(td["obses"])[sl])
            loss, _ = pp.ppo_loss(mu, value.view(-1), ls, fl(td["actions"])[sl], fl(td["neglogpacs"])[sl], fl(adv)[sl], fl(td["values"])[sl], fl(ret)[sl],
                                  mus[sl], old_ls[i])
            opt.zero_grad(); loss.backward()
            torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
            opt.step()
            mus[sl] = mu.detach(); old_ls[i] = ls.detach()
    up.update(last_v, last_d)
    step_ref = torch.cat([p.detach().reshape(-1) for p in ref_params]) - before
    step = up.flat - before
    cos = float(torch.nn.functional.cosine_similarity(step, step_ref, dim=0))
    assert cos > 0.995, cos
    probe = torch.randn(256, 64, device="cuda").clamp(-5, 5)
    with torch.no_grad():
        mu_k, mu_m = model(probe)[0], model.reference_forward(probe)[0]
    assert_close("GNN acts with the updated weights", mu_k, mu_m, rtol=1e-4, atol=1e-5)
