"""Observation / action noise: the port is pinned to the reference Randomizer's own outputs (golden fixture);
the kernel is checked against the port through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import port_policy as pp
from oracle.make_golden import NOISE_CASES, NOISE_N, NOISE_STEPS, noise_inputs
from tests.common import GOLDEN_DIR, assert_close


def _draws(params, zs, us, t):
    """(reset rows, interval rows) in the form the spec's distribution consumes."""
    def pick(spec, which):
        return zs[t][which] if spec["distribution"] in ("gaussian", "normal") else us[t][which]
    dr = pick(params["on_reset"], 0) if "on_reset" in params else None
    di = pick(params["on_interval"], 1) if "on_interval" in params else None
    return dr, di


def _compact(rows, mask):
    """rows of the affected envs in ascending env order -> the [k, dim] tensor the reference draws (full-size buffer, first k rows)."""
    return rows


@pytest.mark.parametrize("case", list(NOISE_CASES))
def test_port_matches_reference_randomizer(case):
    g = np.load(os.path.join(GOLDEN_DIR, "noise_reference.npz"))
    kind, dim, params = NOISE_CASES[case]
    port = pp.NoisePort(params, NOISE_N, dim)
    bufs, resets, zs, us = noise_inputs(case)
    for t in range(NOISE_STEPS):
        dr, di = _draws(params, zs, us, t)
        out = port.apply(bufs[t].clone(), resets[t], dr, di)
        assert np.array_equal(out.numpy(), g[f"{case}/s{t}_out"]), (case, t)
        assert np.array_equal(port.counter.numpy().astype(np.int16), g[f"{case}/s{t}_counter"])
        assert np.array_equal(port.correlated.numpy(), g[f"{case}/s{t}_correlated"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(NOISE_CASES))
@pytest.mark.parametrize("n", [NOISE_N, 5000])
def test_kernel_matches_port(case, n):
    from locomanipulationrl_b200.noise import NoiseRandomizer
    kind, dim, params = NOISE_CASES[case]
    port = pp.NoisePort(params, n, dim)
    rz = NoiseRandomizer(params, n, dim)
    bufs, resets, zs, us = noise_inputs(case, n=n, steps=6)
    loguniform = any(s["distribution"].startswith("log") for s in params.values())
    for t in range(6):
        dr, di = _draws(params, zs, us, t)
        ref = port.apply(bufs[t].clone(), resets[t], dr, di)
        buf = bufs[t].clone().cuda()
        rz.apply(buf, resets[t].cuda(), draws_reset=None if dr is None else dr.cuda(), draws_interval=None if di is None else di.cuda(), compact_draws=True)
        if loguniform:                                       # exp(): <= 2 ulp between torch-CPU (SLEEF) and CUDA expf
            assert_close(f"{case} out s{t}", buf, ref, rtol=1e-6, atol=1e-7)
        else:
            assert torch.equal(buf.cpu(), ref), (case, t, float((buf.cpu() - ref).abs().max()))
            assert torch.equal(rz.correlated_noise.cpu(), port.correlated)
        assert torch.equal(rz.counter.cpu(), port.counter)


@pytest.mark.gpu
def test_kernel_internal_generator_statistics():
    """Without supplied draws the kernel uses its own counter-based generator: check moments and step-to-step independence."""
    from locomanipulationrl_b200.noise import NoiseRandomizer
    params = {"on_reset": {"operation": "additive", "distribution": "gaussian", "distribution_parameters": [0.5, 0.1]},
              "on_interval": {"frequency_interval": 1, "operation": "additive", "distribution": "uniform", "distribution_parameters": [-1.0, 1.0]}}
    n, dim = 65536, 64
    rz = NoiseRandomizer(params, n, dim, seed=123)
    reset = torch.ones(n, dtype=torch.long, device="cuda")
    a = torch.zeros(n, dim, device="cuda")
    rz.apply(a, reset)
    corr = rz.correlated_noise.clone()
    assert abs(float(corr.mean()) - 0.5) < 1e-3 and abs(float(corr.std()) - 0.1) < 1e-3
    u1 = a - corr
    assert abs(float(u1.mean())) < 5e-3 and abs(float(u1.std()) - (2.0 / 12 ** 0.5)) < 5e-3 and float(u1.min()) >= -1.0 and float(u1.max()) <= 1.0
    b = torch.zeros(n, dim, device="cuda")
    rz.apply(b, torch.zeros(n, dtype=torch.long, device="cuda"))
    assert torch.equal(rz.correlated_noise, corr)                          # no resets: the correlated noise stays
    u2 = b - corr
    c = float((u1 * u2).mean() / (u1.std() * u2.std()))
    assert abs(c) < 5e-3, c                                                # fresh draws every step
    # per-channel: columns are not copies of each other
    assert float(torch.corrcoef(u1[:, :8].t()).fill_diagonal_(0).abs().max()) < 0.02
