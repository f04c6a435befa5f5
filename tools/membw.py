"""Write-only / copy bandwidth ceilings on this GPU (context for the env-step roofline: that kernel is ~98% writes)."""
import torch

def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

nbytes = 2244 * (1 << 20)
x = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
y = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
ms = t(lambda: x.fill_(1)); print("fill_ u8    %.1f GB/s" % (nbytes / ms / 1e6))
xf = x.view(torch.float32)
ms = t(lambda: xf.fill_(1.0)); print("fill_ f32   %.1f GB/s" % (nbytes / ms / 1e6))
ms = t(lambda: x.zero_()); print("zero_ (memset) %.1f GB/s" % (nbytes / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy (r+w)  %.1f GB/s" % (2 * nbytes / ms / 1e6))
