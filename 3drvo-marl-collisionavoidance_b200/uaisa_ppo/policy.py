"""Batched actor-critic of the reference (train/policy/policy_rnn_ac.py) over the dense observation layout.

Same architecture and the same ``state_dict`` keys as ``rnn_ac`` (policy_rnn_ac.py:31-72), so the shipped
checkpoints (train/model_save/*_check_point_0.pt, ``model_state``) load with ``strict=True``:
shared bi-GRU reader over the neighbour rows (9 -> 256, final hidden states of both directions SUMMED,
:116-122) -> concat the 12 self values -> LayerNorm(268) -> actor MLP 268-256-256-3 (ReLU, ReLU, Tanh) with a
learned ``log_std`` / critic MLP 268-256-256-1.

What changes is the batching: the reference calls ``ac.step`` once per drone per step with a ragged vector
(multi_ppo.py:193-196); here all A agents go through at once from ``obs[A, 12+9nm]`` + ``count[A]`` exactly as
the env kernel writes them.  A drone without neighbours feeds ONE all-zero row (ir_gym.py:211-212), which is
what the zero padding already is, so step 0 of the GRU runs densely for everybody and only the (few) agents
with two or more rows go through the sequential part.

The rollout forward is the fused tcgen05 kernel (uaisa_ppo/fused.py); in the update the two MLPs run through _MlpTc (csrc/uaisa_mlp.cu, tcgen05)
when `tc_mlp` is set, through torch (cuBLAS) otherwise.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions.normal import Normal


def mlp(sizes, activation, output_activation=nn.Identity):   # policy_rnn_ac.py:10-17
    layers = []
    for j in range(len(sizes) - 1):
        act = activation if j < len(sizes) - 2 else output_activation
        layers += [nn.Linear(sizes[j], sizes[j + 1]), act()]
    return nn.Sequential(*layers)


class _ReaderStep0(torch.autograd.Function):
    """The reader for samples that all have <= 1 neighbour row as ONE forward and ONE backward kernel (csrc/uaisa_train.cu: uaisa_reader_fwd /
    uaisa_reader_bwd): bi-GRU from h0 = 0, direction sum, concat, LayerNorm -- instead of ~60 element-wise / reduction launches over
    [batch x 768] tensors.  fp32, same values and gradients as the torch ops below (tests/test_ppo_gpu.py).  `index` gathers rows of `obs` on the fly."""

    @staticmethod
    def forward(ctx, obs, index, *params):
        import ctypes as C
        from uaisa_env import _capi
        L = _capi.lib()
        B = int(index.shape[0]) if index is not None else int(obs.shape[0])
        feat = torch.empty((B, 268), dtype=torch.float32, device=obs.device)
        stats = torch.empty((B, 2), dtype=torch.float32, device=obs.device)
        hsum = torch.empty((B, 256), dtype=torch.float32, device=obs.device)      # h_fwd + h_bwd before the LayerNorm: spares the backward one gate evaluation
        ptrs = (C.c_void_p * 8)(*[p.data_ptr() for p in params])
        s = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
        rc = L.uaisa_reader_fwd(C.c_void_p(obs.data_ptr()), None if index is None else C.c_void_p(index.data_ptr()), B, int(obs.shape[1]), ptrs,
                                C.c_void_p(feat.data_ptr()), C.c_void_p(stats.data_ptr()), C.c_void_p(hsum.data_ptr()), s)
        if rc:
            raise RuntimeError("uaisa_reader_fwd failed (%d)" % rc)
        ctx.save_for_backward(obs, index, stats, hsum, *params)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        import ctypes as C
        from uaisa_env import _capi
        L = _capi.lib()
        obs, index, stats, hsum, *params = ctx.saved_tensors
        B = int(stats.shape[0])
        dfeat = dfeat.contiguous()
        grads = [torch.zeros_like(p) for p in params]
        ptrs = (C.c_void_p * 8)(*[p.data_ptr() for p in params])
        gptrs = (C.c_void_p * 8)(*[g.data_ptr() for g in grads])
        s = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
        rc = L.uaisa_reader_bwd(C.c_void_p(obs.data_ptr()), None if index is None else C.c_void_p(index.data_ptr()), B, int(obs.shape[1]), ptrs,
                                C.c_void_p(dfeat.data_ptr()), C.c_void_p(stats.data_ptr()), C.c_void_p(hsum.data_ptr()), gptrs, s)
        if rc:
            raise RuntimeError("uaisa_reader_bwd failed (%d)" % rc)
        return (None, None, *grads)


class _MlpTc(torch.autograd.Function):
    """One 268-256-256-nout MLP head (ReLU, ReLU, linear output; the actor's Tanh stays outside) on the tcgen05 tensor cores: forward, input
    gradient and weight gradients are launches of ONE GEMM kernel (csrc/uaisa_mlp.cu: TF32 operands, fp32 accumulation in TMEM, fused bias /
    ReLU / output-layer / ReLU-mask epilogues, split-K vector atomics for the weight gradients) plus two small CUDA-core kernels -- instead of
    torch autograd's cuBLAS SIMT fp32 GEMMs (policy_rnn_ac.py:206-257 forward, multi_ppo.py:379-404 backward through autograd)."""

    SPLITS = 74     # split-K CTAs per 128-row tile of a weight gradient whose CTA tile is 288 wide (one CTA per SM: x 2 row tiles = 148 CTAs)
    SPLITS_256 = 148   # ... whose CTA tile is 256 wide (two CTAs per SM); measured at K = 65 536: 55.8 us instead of 62.0 (profiles/gemm_splits.py)

    @staticmethod
    def _lib():
        import ctypes as C
        from uaisa_env import _capi
        return C, _capi.lib()

    @staticmethod
    def _gemm(L, C, s, A, lda, a_mn, Bm, ldb, b_mn, Cm, ldc, M, N, K, splits, epi, bias=None, mask=None, ldm=0, W3=None, b3=None, Y=None, nout=0):
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        rc = L.uaisa_mlp_gemm(p(A), lda, a_mn, p(Bm), ldb, b_mn, p(Cm), ldc, M, N, K, splits, epi, p(bias), p(mask), ldm, p(W3), p(b3), p(Y), nout, s)
        if rc:
            raise RuntimeError("uaisa_mlp_gemm failed (%d)" % rc)

    @staticmethod
    def forward(ctx, feat, W1, b1, W2, b2, W3, b3):
        C, L = _MlpTc._lib()
        B, nout, dev = int(feat.shape[0]), int(W3.shape[0]), feat.device
        s = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        H1 = torch.empty((B, 256), dtype=torch.float32, device=dev)
        H2 = torch.empty((B, 256), dtype=torch.float32, device=dev)
        y = torch.empty((B, nout), dtype=torch.float32, device=dev)
        _MlpTc._gemm(L, C, s, feat, 268, 0, W1, 268, 0, H1, 256, B, 256, 268, 1, 1, bias=b1)                                  # H1 = relu(feat W1^T + b1)
        _MlpTc._gemm(L, C, s, H1, 256, 0, W2, 256, 0, H2, 256, B, 256, 256, 1, 2, bias=b2, W3=W3, b3=b3, Y=y, nout=nout)      # H2 = relu(H1 W2^T + b2), y = H2 W3^T + b3
        ctx.save_for_backward(feat, W1, W2, W3, H1, H2)
        return y

    @staticmethod
    def backward(ctx, dy):
        C, L = _MlpTc._lib()
        feat, W1, W2, W3, H1, H2 = ctx.saved_tensors
        B, nout, dev = int(feat.shape[0]), int(W3.shape[0]), feat.device
        s = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        dy = dy.contiguous()
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)  # noqa: E731
        dW1, db1, dW2, db2, dW3, db3 = z(256, 268), z(256), z(256, 256), z(256), z(nout, 256), z(nout)
        dZ2 = torch.empty((B, 256), dtype=torch.float32, device=dev)
        dZ1 = torch.empty((B, 256), dtype=torch.float32, device=dev)
        if L.uaisa_mlp_out_bwd(p(dy), p(H2), p(W3), B, nout, p(dZ2), p(dW3), p(db3), p(db2), s):                              # dZ2, db2, dW3, db3
            raise RuntimeError("uaisa_mlp_out_bwd failed")
        sp, sp2 = max(1, min(_MlpTc.SPLITS, (B + 31) // 32)), max(1, min(_MlpTc.SPLITS_256, (B + 31) // 32))
        _MlpTc._gemm(L, C, s, dZ2, 256, 1, H1, 256, 1, dW2, 256, 256, 256, B, sp2, 4)                                          # dW2 = dZ2^T H1       (K = batch)
        _MlpTc._gemm(L, C, s, dZ2, 256, 0, W2, 256, 1, dZ1, 256, B, 256, 256, 1, 3, mask=H1, ldm=256)                         # dZ1 = (dZ2 W2) * (H1 > 0)
        if L.uaisa_colsum(p(dZ1), B, 256, p(db1), s):                                                                         # db1
            raise RuntimeError("uaisa_colsum failed")
        _MlpTc._gemm(L, C, s, dZ1, 256, 1, feat, 268, 1, dW1, 268, 256, 268, B, sp, 4)                                        # dW1 = dZ1^T feat
        dfeat = None
        if ctx.needs_input_grad[0]:
            dfeat = torch.empty((B, 268), dtype=torch.float32, device=dev)
            _MlpTc._gemm(L, C, s, dZ1, 256, 0, W1, 268, 1, dfeat, 268, B, 268, 256, 1, 0)                                     # dfeat = dZ1 W1
        return dfeat, dW1, db1, dW2, db2, dW3, db3


def _tc_mlp_ok(net, feat):
    """The tcgen05 MLP path applies to the reference architecture only: Linear(268,256) ReLU Linear(256,256) ReLU Linear(256,nout<=3) + output activation."""
    lin = [m for m in net if isinstance(m, nn.Linear)]
    return (feat.is_cuda and feat.dtype == torch.float32 and feat.is_contiguous() and feat.shape[1] == 268 and len(lin) == 3 and len(net) == 6
            and isinstance(net[1], nn.ReLU) and isinstance(net[3], nn.ReLU)
            and tuple(lin[0].weight.shape) == (256, 268) and tuple(lin[1].weight.shape) == (256, 256) and lin[2].weight.shape[1] == 256 and lin[2].weight.shape[0] <= 3
            and all(q.is_contiguous() for l in lin for q in (l.weight, l.bias)))


def _run_mlp(net, feat, tc: bool):
    if tc and _tc_mlp_ok(net, feat):
        l0, l1, l2 = net[0], net[2], net[4]
        return net[5](_MlpTc.apply(feat, l0.weight, l0.bias, l1.weight, l1.bias, l2.weight, l2.bias))
    return net(feat)


class rnn_Reader(nn.Module):   # policy_rnn_ac.py:75-168 (mode 'biGRU')
    def __init__(self, state_dim=12, input_dim=9, hidden_dim=256):
        super().__init__()
        self.state_dim, self.input_dim, self.hidden_dim = state_dim, input_dim, hidden_dim
        self.rnn_net = nn.GRU(input_dim, hidden_dim, batch_first=True, bidirectional=True)
        self.ln = nn.LayerNorm(state_dim + hidden_dim)

    def _cell(self, x, h, rev: bool):
        sfx = "_reverse" if rev else ""
        w_ih, w_hh = getattr(self.rnn_net, "weight_ih_l0" + sfx), getattr(self.rnn_net, "weight_hh_l0" + sfx)
        b_ih, b_hh = getattr(self.rnn_net, "bias_ih_l0" + sfx), getattr(self.rnn_net, "bias_hh_l0" + sfx)
        gi, gh = F.linear(x, w_ih, b_ih), F.linear(h, w_hh, b_hh)
        i_r, i_z, i_n = gi.chunk(3, -1)
        h_r, h_z, h_n = gh.chunk(3, -1)
        r, z = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        return (1 - z) * n + z * h

    def _cell0(self, x, rev: bool):
        """GRU cell from h = 0: W_hh h vanishes, so gh is just b_hh (same value and same gradients, 60 % fewer FLOPs)."""
        sfx = "_reverse" if rev else ""
        w_ih, b_ih, b_hh = getattr(self.rnn_net, "weight_ih_l0" + sfx), getattr(self.rnn_net, "bias_ih_l0" + sfx), getattr(self.rnn_net, "bias_hh_l0" + sfx)
        i_r, i_z, i_n = F.linear(x, w_ih, b_ih).chunk(3, -1)
        h_r, h_z, h_n = b_hh.chunk(3, -1)
        r, z = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
        return (1 - z) * torch.tanh(i_n + r * h_n)

    fused_step0 = True       # CUDA + multi_cap < 0: the single-step reader runs as one forward / one backward kernel (_ReaderStep0)

    def _step0_params(self):
        g = self.rnn_net
        return (g.weight_ih_l0, g.bias_ih_l0, g.bias_hh_l0, g.weight_ih_l0_reverse, g.bias_ih_l0_reverse, g.bias_hh_l0_reverse, self.ln.weight, self.ln.bias)

    def forward(self, obs: torch.Tensor, count: torch.Tensor, multi_cap: int = 0, index=None) -> torch.Tensor:
        """obs [B, 12+9nm] float32, count [B] (rows valid; 0 means one zero row) -> LayerNorm(cat(self, h_f + h_b)) [B, 268].
        `index` (int64 [b]): evaluate rows obs[index] / count[index] (the fused single-step path gathers them inside the kernel).

        multi_cap = 0: agents with >= 2 rows are compacted with ``nonzero`` (exact for any input, one host sync).
        multi_cap = K > 0: static shapes and no host sync (CUDA-graph capturable): the K agents with the most rows go
        through the sequential part; exact as long as at most K agents have >= 2 rows -- ``self.overflow`` counts the
        agents that did not fit so the caller can check it once per epoch."""
        if (multi_cap < 0 and self.fused_step0 and obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and self.hidden_dim == 256
                and self.state_dim == 12 and self.input_dim == 9 and all(p.is_contiguous() for p in self._step0_params())):
            return _ReaderStep0.apply(obs, index, *self._step0_params())
        if index is not None:
            obs, count = obs[index], count[index]
        B = obs.shape[0]
        nm = (obs.shape[1] - self.state_dim) // self.input_dim
        self_state = obs[:, : self.state_dim]
        rows = obs[:, self.state_dim: self.state_dim + nm * self.input_dim].reshape(B, nm, self.input_dim)
        x0 = rows[:, 0] if nm > 0 else obs.new_zeros((B, self.input_dim))
        hsum = self._cell0(x0, False) + self._cell0(x0, True)                 # exact for every agent with <= 1 row
        if nm > 1 and multi_cap >= 0:   # multi_cap < 0: the caller guarantees that nobody has >= 2 rows
            if multi_cap > 0:
                K = min(int(multi_cap), B)
                c_top, multi = torch.topk(count.to(torch.int32), K)           # static shape, no sync
                c_m, lmax = c_top.to(torch.long), nm
                if getattr(self, "overflow", None) is not None:               # device counter owned by the caller (graph-safe: in place)
                    self.overflow += ((count > 1).sum() - (c_top > 1).sum()).to(torch.int64)
            else:
                multi = torch.nonzero(count > 1).squeeze(-1)                   # (one host sync; few agents)
                c_m = count[multi].to(torch.long)
                lmax = int(c_m.max()) if multi.numel() > 0 else 0
            if lmax > 0:
                r_m = rows[multi]
                hf = obs.new_zeros((multi.shape[0], self.hidden_dim)); hb = hf.clone()
                c_eff = c_m.clamp(min=1)                                       # an agent without rows feeds one zero row
                for l in range(lmax):                                          # forward direction: x_0 .. x_{len-1}
                    m = (l < c_eff).unsqueeze(-1)
                    hf = torch.where(m, self._cell(r_m[:, l], hf, False), hf)
                for l in range(lmax - 1, -1, -1):                              # reverse direction: x_{len-1} .. x_0
                    m = (l < c_eff).unsqueeze(-1)
                    hb = torch.where(m, self._cell(r_m[:, l], hb, True), hb)
                hsum = hsum.index_copy(0, multi, hf + hb)
        return self.ln(torch.cat((self_state, hsum), dim=1))


class GaussianActor(nn.Module):   # policy_rnn_ac.py:191-235
    def __init__(self, obs_dim, act_dim, hidden_sizes, activation, output_activation, rnn_reader):
        super().__init__()
        self.rnn_reader = rnn_reader
        self.net_out = mlp([obs_dim] + list(hidden_sizes) + [act_dim], activation, output_activation)
        self.log_std = nn.Parameter(torch.as_tensor(-1 * np.ones(act_dim, dtype=np.float32)))

    tc_mlp = False      # MLP forward / backward through csrc/uaisa_mlp.cu (tcgen05, TF32 operands); set by PPOTrainer for the update

    def distribution_from_features(self, feat, std_factor=1.0):
        mu = _run_mlp(self.net_out, feat, self.tc_mlp)
        std = torch.clamp(std_factor * torch.exp(self.log_std) + 1e-6, min=1e-4, max=10.0)
        return Normal(mu, std, validate_args=False)   # validation would synchronise (and break graph capture)


class Critic(nn.Module):   # policy_rnn_ac.py:238-257
    def __init__(self, obs_dim, hidden_sizes, activation, output_activation, rnn_reader):
        super().__init__()
        self.v_net = mlp([obs_dim] + list(hidden_sizes) + [1], activation, output_activation)
        self.rnn_reader = rnn_reader

    tc_mlp = False


class rnn_ac(nn.Module):   # policy_rnn_ac.py:31-72, defaults of train_process.py:37-48
    def __init__(self, state_dim=12, rnn_input_dim=9, rnn_hidden_dim=256, hidden_sizes_ac=(256, 256), hidden_sizes_v=(256, 256),
                 act_dim=3, activation=nn.ReLU, output_activation=nn.Tanh, output_activation_v=nn.Identity):
        super().__init__()
        obs_dim = rnn_hidden_dim + state_dim
        rnn = rnn_Reader(state_dim, rnn_input_dim, rnn_hidden_dim)
        self.pi = GaussianActor(obs_dim, act_dim, hidden_sizes_ac, activation, output_activation, rnn)
        self.v = Critic(obs_dim, hidden_sizes_v, activation, output_activation_v, rnn)     # the reader is SHARED (:46-54)

    def features(self, obs, count, multi_cap=0, index=None):
        return self.pi.rnn_reader(obs, count, multi_cap, index)

    def evaluate(self, obs, count, act=None, std_factor=1.0, multi_cap=0, heads="both", index=None):
        """-> (Normal, value[B], logp[B] or None); one reader pass feeds both heads.  `heads` = "pi" / "v" skips the head an update
        iteration does not use (its forward would be thrown away): the skipped entry comes back as None."""
        feat = self.features(obs, count, multi_cap, index)
        dist = self.pi.distribution_from_features(feat, std_factor) if heads != "v" else None
        v = torch.squeeze(_run_mlp(self.v.v_net, feat, self.v.tc_mlp), -1) if heads != "pi" else None
        logp = dist.log_prob(act).sum(-1) if (act is not None and dist is not None) else None
        return dist, v, logp

    @torch.no_grad()
    def step(self, obs, count, std_factor=1.0, multi_cap=0):
        """ac.step for all agents (policy_rnn_ac.py:57-69): sampled action increment, value, log-prob (CUDA tensors)."""
        dist, v, _ = self.evaluate(obs, count, None, std_factor, multi_cap)
        # Normal.sample() = torch.normal(mu, std), which checks std >= 0 with a host synchronisation; mu + std * eps is the
        # same distribution without it (needed inside a CUDA graph)
        a = dist.mean + dist.stddev * torch.randn_like(dist.mean)
        return a, v, dist.log_prob(a).sum(-1)

    @torch.no_grad()
    def act(self, obs, count, std_factor=1.0):
        return self.step(obs, count, std_factor)[0]
