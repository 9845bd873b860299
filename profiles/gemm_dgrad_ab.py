"""A/B of the input-gradient GEMMs of the update (dZ1 = (dZ2 W2) * mask, dfeat = dZ1 W1): W read as an MN-major B operand (as it lies) vs a pre-transposed
K-major copy; and of the forward GEMM for reference.  python profiles/gemm_dgrad_ab.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
import torch, ctypes as C
from uaisa_env import _capi
from uaisa_ppo.policy import _MlpTc
L = _capi.lib()
B = 65536
dZ = torch.randn(B, 256, device="cuda"); H = torch.randn(B, 256, device="cuda"); X = torch.randn(B, 268, device="cuda")
W2 = torch.randn(256, 256, device="cuda"); W1 = torch.randn(256, 268, device="cuda"); b = torch.randn(256, device="cuda")
W2T, W1T = W2.t().contiguous(), W1.t().contiguous()
out = torch.empty(B, 256, device="cuda"); out2 = torch.empty(B, 268, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def t(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-60s %.1f us" % (name, e0.elapsed_time(e1) / 20 * 1e3))
t("forward H2 = relu(H1 W2^T + b)           (both K-major)", lambda: _MlpTc._gemm(L, C, s, H, 256, 0, W2, 256, 0, out, 256, B, 256, 256, 1, 1, bias=b))
t("dgrad dZ1 = (dZ2 W2) * mask   B = W2 as it lies (MN-major)", lambda: _MlpTc._gemm(L, C, s, dZ, 256, 0, W2, 256, 1, out, 256, B, 256, 256, 1, 3, mask=H, ldm=256))
t("dgrad dZ1 = (dZ2 W2) * mask   B = W2^T copy (K-major)", lambda: _MlpTc._gemm(L, C, s, dZ, 256, 0, W2T, 256, 0, out, 256, B, 256, 256, 1, 3, mask=H, ldm=256))
t("dgrad dfeat = dZ1 W1          B = W1 as it lies (MN-major)", lambda: _MlpTc._gemm(L, C, s, dZ, 256, 0, W1, 268, 1, out2, 268, B, 268, 256, 1, 0))
t("dgrad dfeat = dZ1 W1          B = W1^T copy (K-major)", lambda: _MlpTc._gemm(L, C, s, dZ, 256, 0, W1T, 256, 0, out2, 268, B, 268, 256, 1, 0))
t("plain store, both K-major", lambda: _MlpTc._gemm(L, C, s, dZ, 256, 0, W2T, 256, 0, out, 256, B, 256, 256, 1, 0))
