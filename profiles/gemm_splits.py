import sys
sys.path[:0] = ["/root/repo", "/root/repo/3drvo-marl-collisionavoidance_b200", "/root/repo/tests"]
import torch, ctypes as C
from uaisa_env import _capi
from uaisa_ppo.policy import _MlpTc
L = _capi.lib()
B = 65536
dZ = torch.randn(B, 256, device="cuda"); H = torch.randn(B, 256, device="cuda"); X = torch.randn(B, 268, device="cuda")
for N, Bm, ldb in ((256, H, 256), (268, X, 268)):
    for sp in (37, 74, 111, 148, 296):
        out = torch.zeros(256, N, device="cuda")
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            _MlpTc._gemm(L, C, s, dZ, 256, 1, Bm, ldb, 1, out, N, 256, N, B, sp, 4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _MlpTc._gemm(L, C, s, dZ, 256, 1, Bm, ldb, 1, out, N, 256, N, B, sp, 4)
        e1.record(); torch.cuda.synchronize()
        print("wgrad N=%d splits=%d: %.1f us" % (N, sp, e0.elapsed_time(e1) / 20 * 1e3))
