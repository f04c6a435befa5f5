import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onitama_alphazero_b200.net import ConvResNet
torch.manual_seed(0)
def t(label, fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); a = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); ms = (time.perf_counter() - a) / reps * 1e3
    print("%-40s %.3f ms  %.2f M evals/s" % (label, ms, x.shape[0] / ms / 1e3), flush=True)
for B in (4096, 16384):
    x = (torch.rand(B, 21, 5, 5, device="cuda") > 0.7).float()
    m = ConvResNet(64, 21, 3).cuda().eval()
    with torch.no_grad():
        t("B=%d default" % B, lambda: m(x))
        torch.backends.cudnn.benchmark = True
        t("B=%d cudnn.benchmark" % B, lambda: m(x))
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            m(x); torch.cuda.synchronize()
            with torch.cuda.graph(g): y = m(x)
        t("B=%d cuda graph" % B, lambda: g.replay())
        mc = ConvResNet(64, 21, 3).cuda().eval().to(memory_format=torch.channels_last)
        xc = x.contiguous(memory_format=torch.channels_last)
        t("B=%d channels_last" % B, lambda: mc(xc))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            t("B=%d channels_last bf16 autocast" % B, lambda: mc(xc))
        torch.backends.cudnn.benchmark = False
