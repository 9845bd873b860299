"""CPU: host-side logic of the PPO update's flat buckets (no kernels are launched)."""
import torch

import _util  # noqa: F401


def test_flat_params_cover_the_module_and_alias_it():
    from uaisa_ppo.policy import rnn_ac
    from uaisa_ppo.trainer import FlatParams

    torch.manual_seed(0)
    ac = rnn_ac()
    before = {k: v.clone() for k, v in ac.state_dict().items()}
    F = FlatParams(ac, torch.device("cpu"))
    assert F.n == 680991 == sum(p.numel() for p in ac.parameters())          # SURVEY.md section 8a row a-P1
    assert all(torch.equal(v, before[k]) for k, v in ac.state_dict().items())
    # optimiser ranges: Adam(ac.v.parameters()) = v_net + reader = [0, v_hi); Adam(ac.pi.parameters()) = reader + net_out + log_std = [pi_lo, n)
    assert F.v_hi == sum(p.numel() for p in ac.v.parameters()) and F.n - F.pi_lo == sum(p.numel() for p in ac.pi.parameters())
    # parameters and gradients are views of the buckets
    F.p.zero_()
    assert all(float(p.abs().sum()) == 0.0 for p in ac.parameters())
    F.g[:F.n].fill_(2.0)
    assert all(bool((p.grad == 2.0).all()) for p in ac.parameters())
    x = torch.randn(5, 102); x[:, 21:] = 0
    cnt = torch.ones(5, dtype=torch.uint8)
    F.g.zero_()
    _, v, _ = ac.evaluate(x, cnt, None, heads="v")
    v.sum().backward()                                                         # autograd accumulates IN PLACE into the bucket
    assert float(F.g[:F.v_hi].abs().sum()) > 0 and float(F.g[F.v_hi:F.n].abs().sum()) == 0.0
    sd = F.torch_state_dict(type("O", (), {"step": torch.zeros(1, dtype=torch.int32), "m": torch.zeros(F.n), "v": torch.zeros(F.n), "lr": 1e-3, "betas": (0.9, 0.999), "eps": 1e-8})(), 0, F.v_hi)
    assert len(sd["state"]) == len(list(ac.v.parameters())) and sd["param_groups"][0]["lr"] == 1e-3
