"""-m gpu: the tcgen05 GEMM of the PPO update (csrc/uaisa_mlp.cu) against torch fp32 matmuls -- every operand orientation, epilogue, ragged edge and
the split-K path -- and the MLP head built from it (forward + all gradients) against torch autograd on the same module.
Tolerance: TF32 operands (10-bit mantissa, truncated) with fp32 accumulation: relative error ~ 2^-10 per product, ~1e-3 of the result's scale."""
import ctypes as C

import numpy as np
import pytest
import torch

import _util  # noqa: F401
from uaisa_env import _capi
from uaisa_ppo.policy import _MlpTc, rnn_ac

pytestmark = pytest.mark.gpu


def _gemm(A, a_mn, B, b_mn, Cm, M, N, K, splits=1, epi=0, **kw):
    L = _capi.lib()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _MlpTc._gemm(L, C, s, A, A.stride(0), a_mn, B, B.stride(0), b_mn, Cm, 0 if Cm is None else Cm.stride(0), M, N, K, splits, epi, **kw)
    torch.cuda.synchronize()


def _scale_err(got, ref):
    return float((got - ref).abs().max() / ref.abs().max().clamp(min=1e-6))


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (300, 256, 268), (1000, 268, 256), (77, 40, 33), (256, 268, 5000), (4133, 256, 268)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_gemm_tf32_every_orientation(M, N, K, a_mn, b_mn):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + a_mn * 2 + b_mn)
    A = torch.randn((M, K), device="cuda", generator=g)
    B = torch.randn((N, K), device="cuda", generator=g)
    ref = A.double() @ B.double().t()
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    out = torch.full((M, N), 7.0, device="cuda")
    _gemm(As, a_mn, Bs, b_mn, out, M, N, K)
    assert _scale_err(out.double(), ref) < 2e-3, _scale_err(out.double(), ref)
    if K >= 64:          # split-K with atomics, on top of a non-zero C
        out2 = torch.ones((M, N), device="cuda")
        _gemm(As, a_mn, Bs, b_mn, out2, M, N, K, splits=5, epi=4)
        assert _scale_err(out2.double() - 1.0, ref) < 2e-3


def test_gemm_tf32_unaligned_operands_take_the_scalar_path():
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 200, 256, 268
    bufA = torch.randn(M * K + 1, device="cuda", generator=g); bufB = torch.randn(N * K + 3, device="cuda", generator=g)
    A, B = bufA[1:].view(M, K), bufB[3:].view(N, K)            # 4- and 12-byte offsets: not 16-byte aligned (a parameter inside the flat bucket)
    out = torch.empty((M, N), device="cuda")
    _gemm(A, 0, B, 0, out, M, N, K)
    assert _scale_err(out.double(), A.double() @ B.double().t()) < 2e-3


def test_gemm_epilogues():
    g = torch.Generator(device="cuda").manual_seed(11)
    M, N, K = 700, 256, 268
    A = torch.randn((M, K), device="cuda", generator=g); B = torch.randn((N, K), device="cuda", generator=g) * 0.1
    bias = torch.randn(N, device="cuda", generator=g)
    pre = (A.double() @ B.double().t() + bias.double())
    out = torch.empty((M, N), device="cuda")
    _gemm(A, 0, B, 0, out, M, N, K, epi=1, bias=bias)
    assert _scale_err(out.double(), pre.clamp(min=0)) < 2e-3
    for nout in (1, 3):
        W3 = torch.randn((nout, N), device="cuda", generator=g) * 0.1; b3 = torch.randn(nout, device="cuda", generator=g)
        out2 = torch.empty((M, N), device="cuda"); Y = torch.empty((M, nout), device="cuda")
        _gemm(A, 0, B, 0, out2, M, N, K, epi=2, bias=bias, W3=W3, b3=b3, Y=Y, nout=nout)
        assert torch.equal(out2, out)
        assert _scale_err(Y.double(), out.double() @ W3.double().t() + b3.double()) < 1e-4      # the output layer is fp32 FMA on the CUDA cores
    mask = torch.randn((M, N), device="cuda", generator=g)
    out3 = torch.empty((M, N), device="cuda")
    _gemm(A, 0, B, 0, out3, M, N, K, epi=3, mask=mask, ldm=N)
    plain = torch.empty((M, N), device="cuda"); _gemm(A, 0, B, 0, plain, M, N, K)
    assert torch.equal(out3, torch.where(mask > 0, plain, torch.zeros_like(plain)))


def _tf32(x):
    """What kind::tf32 reads from an fp32 operand: the low 13 mantissa bits are ignored (truncation)."""
    return (x.contiguous().view(torch.int32) & -8192).view(torch.float32).double()


def _emulated(feat, dy, net):
    """The head's forward / backward with every tensor-core GEMM restated in fp64 on TF32-truncated operands (so that ReLU masks agree with the
    kernel's and the comparison can be tight); the output layer and the element-wise parts are fp32 FMA in the kernels -> fp64 here."""
    lin = [m for m in net if isinstance(m, torch.nn.Linear)]
    W1, b1, W2, b2, W3, b3 = (t.detach() for l in lin for t in (l.weight, l.bias))
    H1 = (_tf32(feat) @ _tf32(W1).t() + b1.double()).clamp(min=0).float()
    H2 = (_tf32(H1) @ _tf32(W2).t() + b2.double()).clamp(min=0).float()
    y = H2.double() @ W3.double().t() + b3.double()
    dZ2 = ((dy.double() @ W3.double()) * (H2 > 0)).float()
    dW3, db3, db2 = dy.double().t() @ H2.double(), dy.double().sum(0), dZ2.double().sum(0)
    dW2 = _tf32(dZ2).t() @ _tf32(H1)
    dZ1 = ((_tf32(dZ2) @ _tf32(W2)) * (H1 > 0)).float()
    db1 = dZ1.double().sum(0)
    dW1 = _tf32(dZ1).t() @ _tf32(feat)
    dfeat = _tf32(dZ1) @ _tf32(W1)
    return [y, dfeat, dW1, db1, dW2, db2, dW3, db3]


@pytest.mark.parametrize("B,nout", [(1, 3), (300, 1), (4133, 3), (65536, 1)])
def test_mlp_head_forward_and_gradients_match_torch_autograd(B, nout):
    torch.manual_seed(B + nout)
    ac = rnn_ac().cuda()
    net = ac.pi.net_out if nout == 3 else ac.v.v_net
    feat = torch.randn((B, 268), device="cuda")
    dy = torch.randn((B, nout), device="cuda") / B
    res = {}
    for tc in (False, True):
        f = feat.clone().requires_grad_(True)
        for q in net.parameters():
            q.grad = None
        lin = [m for m in net if isinstance(m, torch.nn.Linear)]
        y = _MlpTc.apply(f, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias) if tc else net[4](net[3](net[2](net[1](net[0](f)))))
        (y * dy).sum().backward()
        res[tc] = [y.detach().clone(), f.grad.clone()] + [q.grad.clone() for q in net.parameters()]
    names = ["y", "dfeat", "dW1", "db1", "dW2", "db2", "dW3", "db3"]
    emu = _emulated(feat, dy, net)
    for name, a, e, b in zip(names, res[True], emu, res[False]):
        assert a.shape == b.shape == e.shape, name
        # (1) against the TF32-operand restatement: same operand precision, so what is left is accumulation order, intermediate values whose last
        #     fp32 bit moves them across a TF32 truncation boundary, and the one-in-a-million hidden unit whose pre-activation lies within the
        #     accumulation noise of zero: its ReLU mask flips and moves that sample's whole gradient row by ~1/16 (a few units per 10^6: ~3e-3 in
        #     Frobenius norm whatever the batch size)
        fro_e = float((a.double() - e).norm() / e.norm().clamp(min=1e-30))
        assert fro_e < (1e-3 if name == "y" else 1.5e-2), (name, fro_e)
        if name == "y":
            assert _scale_err(a.double(), e) < 5e-4, _scale_err(a.double(), e)
        # (2) against torch fp32 autograd: TF32 operand error, plus a few ReLU masks that flip for pre-activations within 1e-3 of zero
        fro = float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))
        assert fro < 5e-2 and _scale_err(a.double(), b.double()) < 0.15, (name, fro, _scale_err(a.double(), b.double()))
    assert _scale_err(res[True][0].double(), res[False][0].double()) < 4e-3     # the forward has no masks to flip downstream


def test_update_iteration_with_the_tc_mlp_follows_the_fp32_update():
    """A few PPO iterations with the tcgen05 MLPs stay close to the same iterations through torch fp32 (same buffer, same seeds)."""
    from uaisa_env.scenario import scenario_c1
    from uaisa_env.vec_env import VecMdin
    from uaisa_ppo.trainer import PPOConfig, PPOTrainer
    outs = []
    for tc in (False, True):
        env = VecMdin(scenario_c1(), 64, neighbors_num=10)
        tr = PPOTrainer(env, cfg=PPOConfig(steps_per_epoch=16, train_pi_iters=3, train_v_iters=3, minibatch=2048, pi_lr=1e-3, vf_lr=1e-3, tc_mlp=tc), seed=3)
        p0 = tr.flat.p.clone()
        tr.rollout()
        st = tr.update()
        outs.append((tr.flat.p.clone() - p0, st["loss_v"], st["kl"]))
        assert ("tcgen05" in tr.update_path) == tc
        tr.close(); env.close()
    d0, d1 = outs[0][0], outs[1][0]
    assert float(d0.abs().max()) > 0
    cos = float((d0 * d1).sum() / (d0.norm() * d1.norm()))
    assert cos > 0.98, cos          # Adam normalises the step, so compare directions: the parameter updates point the same way
    assert abs(outs[0][1] - outs[1][1]) < 2e-2 * max(1.0, abs(outs[0][1]))
