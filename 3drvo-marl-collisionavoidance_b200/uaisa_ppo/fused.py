"""Fused rollout forward of rnn_ac through libuaisa_b200.so.

Default (``kernel="full"``, csrc/uaisa_policy_fwd.cu): ONE kernel for the whole ``ac.step`` forward (policy_rnn_ac.py:57-69) -- the masked bi-GRU
recurrence over each agent's neighbour rows, LayerNorm(268) and both MLP heads on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulation in
TMEM), weights fetched as pre-swizzled tiles by the bulk-copy engine.  No torch branch, no host synchronisation, any count in 0..nm.
``kernel="split"`` keeps the round-1 pair (csrc/uaisa_policy.cu: K6a step-0 features on the CUDA cores + K6b tcgen05 heads, agents with >= 2 rows
patched from the torch reader) for A/B runs.  Inference only (the update goes through torch autograd); call :meth:`refresh` after every optimiser step."""
from __future__ import annotations

import ctypes as C

import torch

from uaisa_env import _capi
from .policy import rnn_ac

KPAD = 288


class FusedPolicy:
    def __init__(self, ac: rnn_ac, device, kernel: str = "full"):
        self.ac, self.dev = ac, torch.device(device)
        if self.dev.type == "cuda" and self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self.L = _capi.lib()
        if kernel not in ("full", "split"):
            raise ValueError("kernel must be 'full' or 'split'")
        self.kernel = kernel
        rd = ac.pi.rnn_reader
        shapes = (tuple(rd.rnn_net.weight_ih_l0.shape), tuple(rd.rnn_net.weight_hh_l0.shape), tuple(rd.ln.weight.shape), tuple(ac.pi.net_out[0].weight.shape),
                  tuple(ac.pi.net_out[2].weight.shape), tuple(ac.pi.net_out[4].weight.shape), tuple(ac.v.v_net[0].weight.shape), tuple(ac.v.v_net[4].weight.shape))
        if shapes != ((768, 9), (768, 256), (268,), (256, 268), (256, 256), (3, 256), (256, 268), (1, 256)) or not rd.rnn_net.bidirectional:
            raise ValueError("the fused kernels are built for the reference architecture (bi-GRU 9->256, LayerNorm 268, MLPs 268-256-256-3/1); got %r" % (shapes,))
        self._scratch = None
        self.stats = torch.zeros(2, dtype=torch.int32, device=self.dev)      # {tiles that ran the recurrence, max steps}
        self.refresh()

    def _params22(self):
        g, ln, pi, v = self.ac.pi.rnn_reader.rnn_net, self.ac.pi.rnn_reader.ln, self.ac.pi.net_out, self.ac.v.v_net
        ps = [g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0, g.weight_ih_l0_reverse, g.weight_hh_l0_reverse, g.bias_ih_l0_reverse, g.bias_hh_l0_reverse,
              ln.weight, ln.bias, pi[0].weight, pi[0].bias, pi[2].weight, pi[2].bias, pi[4].weight, pi[4].bias,
              v[0].weight, v[0].bias, v[2].weight, v[2].bias, v[4].weight, v[4].bias]
        for t in ps:
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.dev:
                raise ValueError("fused forward needs contiguous fp32 parameters on %s" % self.dev)
        return ps

    @torch.no_grad()
    def refresh(self):
        """(Re)pack the module's weights into the kernels' layouts.  Buffers are allocated once and updated in place, so
        device addresses stay fixed (a CUDA graph that captured the fused forward stays valid)."""
        if self.kernel == "full":
            if getattr(self, "wpack", None) is None:
                self.wpack = torch.zeros(int(self.L.uaisa_policy_pack_floats()), dtype=torch.float32, device=self.dev)
            ps = self._params22()
            ptrs = (C.c_void_p * 22)(*[t.data_ptr() for t in ps])
            rc = self.L.uaisa_policy_pack(ptrs, C.c_void_p(self.wpack.data_ptr()), self._stream())
            if rc:
                raise _capi.UaisaError("uaisa_policy_pack failed (%d)" % rc)
            return
        g, ln = self.ac.pi.rnn_reader.rnn_net, self.ac.pi.rnn_reader.ln
        H = 256
        wih, bias = [], []
        for sfx in ("", "_reverse"):
            w_ih, b_ih, b_hh = getattr(g, "weight_ih_l0" + sfx), getattr(g, "bias_ih_l0" + sfx), getattr(g, "bias_hh_l0" + sfx)
            wih.append(w_ih.t())                                                # [9][768]
            bias.append(torch.stack([b_ih[:H] + b_hh[:H], b_ih[H:2 * H] + b_hh[H:2 * H], b_ih[2 * H:], b_hh[2 * H:]]))   # [4][256]
        new = [torch.stack(wih).float(), torch.stack(bias).float(), ln.weight.detach().float(), ln.bias.detach().float()]
        for net in (self.ac.pi.net_out, self.ac.v.v_net):
            w1 = torch.zeros((256, KPAD), device=self.dev); w1[:, :268] = net[0].weight
            new += [w1, net[0].bias.detach().float(), net[2].weight.detach().float(), net[2].bias.detach().float(),
                    net[4].weight.detach().float(), net[4].bias.detach().float()]
        if getattr(self, "_bufs", None) is None:
            self._bufs = [t.contiguous().to(self.dev).clone() for t in new]
        else:
            for b, t in zip(self._bufs, new):
                b.copy_(t)
        self.wih_t, self.bias, self.ln_w, self.ln_b = self._bufs[:4]
        self.ws = self._bufs[4:]
        self.wptr = (C.c_void_p * 12)(*[w.data_ptr() for w in self.ws])

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    # ---- kernel = "full" ---------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, obs: torch.Tensor, count: torch.Tensor, want_features: bool = False, out_v: torch.Tensor = None):
        """-> (mu [A,3], v [A], LayerNorm features [A,268] or None); `out_v` (contiguous fp32 [A]): where the critic's values are written."""
        A = obs.shape[0]
        if obs.dtype != torch.float32 or not obs.is_contiguous() or count.dtype != torch.uint8 or count.numel() != A:
            raise ValueError("obs must be contiguous float32 [A, 12+9nm] and count uint8 [A]")
        tiles = (A + 127) // 128
        if self._scratch is None or self._scratch.numel() < tiles * 128 * 256:
            self._scratch = torch.empty(tiles * 128 * 256, dtype=torch.float32, device=self.dev)
        mu = torch.empty((A, 3), dtype=torch.float32, device=self.dev)
        v = self._out(out_v, (A,)) if out_v is not None else torch.empty(A, dtype=torch.float32, device=self.dev)
        feat = torch.empty((A, 268), dtype=torch.float32, device=self.dev) if want_features else None
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        rc = self.L.uaisa_policy_forward(p(obs), p(count), A, obs.shape[1], p(self.wpack), p(mu), p(v), p(feat), p(self._scratch), p(self.stats), self._stream())
        if rc:
            raise _capi.UaisaError("uaisa_policy_forward failed (%d)" % rc)
        return mu, v, feat

    def _out(self, t, shape):
        if t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != tuple(shape) or t.device != self.dev:
            raise ValueError("output buffer must be contiguous float32 %r on %s" % (tuple(shape), self.dev))
        return t

    @torch.no_grad()
    def sample(self, mu: torch.Tensor, std_factor: float = 1.0, out_logp: torch.Tensor = None, eps: torch.Tensor = None):
        """a = mu + std eps and its log-probability (policy_rnn_ac.py:57-69) in one launch (uaisa_policy_sample) -> (a [A,3], logp [A])."""
        A = mu.shape[0]
        if eps is None:
            eps = torch.randn_like(mu)
        a = torch.empty_like(mu)
        logp = self._out(out_logp, (A,)) if out_logp is not None else torch.empty(A, dtype=torch.float32, device=self.dev)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        ls = self.ac.pi.log_std
        if ls.dtype != torch.float32 or not ls.is_contiguous() or ls.numel() != 3:
            raise ValueError("log_std must be contiguous float32 [3]")
        rc = self.L.uaisa_policy_sample(p(mu), p(eps), p(ls), float(std_factor), A, p(a), p(logp), self._stream())
        if rc:
            raise _capi.UaisaError("uaisa_policy_sample failed (%d)" % rc)
        return a, logp

    # ---- kernel = "split" --------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def features(self, obs: torch.Tensor) -> torch.Tensor:
        A = obs.shape[0]
        feat = torch.empty((A, KPAD), dtype=torch.float32, device=self.dev)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        rc = self.L.uaisa_policy_features(p(obs), A, obs.shape[1], p(self.wih_t), p(self.bias), p(self.ln_w), p(self.ln_b), p(feat), self._stream())
        if rc:
            raise _capi.UaisaError("uaisa_policy_features failed (%d)" % rc)
        return feat

    @torch.no_grad()
    def heads(self, feat: torch.Tensor):
        A = feat.shape[0]
        mu = torch.empty((A, 3), dtype=torch.float32, device=self.dev)
        v = torch.empty(A, dtype=torch.float32, device=self.dev)
        rc = self.L.uaisa_policy_heads(C.c_void_p(feat.data_ptr()), A, self.wptr, C.c_void_p(mu.data_ptr()), C.c_void_p(v.data_ptr()), self._stream())
        if rc:
            raise _capi.UaisaError("uaisa_policy_heads failed (%d)" % rc)
        return mu, v

    def _split_mu_v(self, obs, count, multi_cap):
        feat = self.features(obs)
        rd = self.ac.pi.rnn_reader
        if multi_cap < 0:
            pass                      # the caller has checked that no agent has >= 2 rows: K6a is exact for everybody
        elif multi_cap > 0:
            K = min(int(multi_cap), obs.shape[0])
            c_top, idx = torch.topk(count.to(torch.int32), K)
            if getattr(rd, "overflow", None) is not None:
                rd.overflow += ((count > 1).sum() - (c_top > 1).sum()).to(torch.int64)
            ov = getattr(rd, "overflow", None); rd.overflow = None          # the inner call sees exactly K agents: nothing can overflow
            f_t = rd(obs[idx], c_top.to(count.dtype), multi_cap=K)                # static shapes, no host sync (graph-capturable)
            rd.overflow = ov
            feat[:, :268].index_copy_(0, idx, torch.where((c_top > 1).unsqueeze(-1), f_t, feat[idx, :268]))
        else:
            idx = torch.nonzero(count > 1).squeeze(-1)
            if idx.numel():
                feat[idx, :268] = rd(obs[idx], count[idx])
        return self.heads(feat)

    # -----------------------------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, obs, count, std_factor=1.0, multi_cap=0, out_v=None, out_logp=None):
        """Drop-in for rnn_ac.step: (a, v, logp).  (`multi_cap` only matters for kernel = "split"; `out_v` / `out_logp`: rollout-buffer rows the values and
        log-probabilities are written into directly.)"""
        if self.kernel == "full":
            mu, v, _ = self.forward(obs, count, out_v=out_v)
            a, logp = self.sample(mu, std_factor, out_logp=out_logp)
            return a, v, logp
        mu, v = self._split_mu_v(obs, count, multi_cap)
        std = torch.clamp(std_factor * torch.exp(self.ac.pi.log_std) + 1e-6, min=1e-4, max=10.0)
        a = mu + std * torch.randn_like(mu)
        logp = (-((a - mu) ** 2) / (2 * std * std) - torch.log(std) - 0.9189385332046727).sum(-1)
        if out_v is not None:
            out_v.copy_(v); v = out_v
        if out_logp is not None:
            out_logp.copy_(logp); logp = out_logp
        return a, v, logp
