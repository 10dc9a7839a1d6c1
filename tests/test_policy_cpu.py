"""Host-side policy code that runs without a GPU: the differentiable torch forward of the GNN module (the training path)."""
import os

import numpy as np
import torch

from tests.common import GOLDEN_DIR


def test_gnn_training_forward_matches_reference_golden_and_is_differentiable():
    from locomanipulationrl_b200.policy.gnn import GnnActorCritic
    g = np.load(os.path.join(GOLDEN_DIR, "gnn_reference.npz"))
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p/")}
    m = GnnActorCritic(device="cpu")
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and set(res.missing_keys) <= {"log_std_parameter"}
    mu, value, log_std = m.reference_forward(torch.from_numpy(g["obs"]))
    assert np.allclose(mu.detach().numpy(), g["mu"], rtol=1e-5, atol=2e-6) and np.allclose(value.detach().numpy(), g["value"], rtol=1e-5, atol=2e-6)
    (mu.square().sum() + value.square().sum() + log_std.sum()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
