"""Fused rollout forward of rnn_ac through libuaisa_b200.so (csrc/uaisa_policy.cu): K6a features (bi-GRU step 0 + LayerNorm,
CUDA cores) and K6b heads (actor / critic MLPs on tcgen05 tensor cores, TF32 inputs, fp32 accumulation in TMEM).
Inference only (the update goes through torch autograd); call :meth:`refresh` after every optimiser step."""
from __future__ import annotations

import ctypes as C

import torch

from uaisa_env import _capi
from .policy import rnn_ac

KPAD = 288


class FusedPolicy:
    def __init__(self, ac: rnn_ac, device):
        self.ac, self.dev = ac, torch.device(device)
        self.L = _capi.lib()
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        """(Re)pack the module's weights into the kernels' layouts.  Buffers are allocated once and updated in place, so
        device addresses stay fixed (a CUDA graph that captured the fused forward stays valid)."""
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

    @torch.no_grad()
    def step(self, obs, count, std_factor=1.0, multi_cap=0):
        """Drop-in for rnn_ac.step: (a, v, logp).  Agents with >= 2 neighbour rows get their feature row from the torch reader."""
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
        mu, v = self.heads(feat)
        std = torch.clamp(std_factor * torch.exp(self.ac.pi.log_std) + 1e-6, min=1e-4, max=10.0)
        a = mu + std * torch.randn_like(mu)
        logp = (-((a - mu) ** 2) / (2 * std * std) - torch.log(std) - 0.9189385332046727).sum(-1)
        return a, v, logp
