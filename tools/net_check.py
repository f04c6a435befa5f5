"""GPU check of the fused network kernel against the PyTorch restatement (f32, TF32 off): python tools/net_check.py [n] [blocks]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import onitama_alphazero_b200 as onb
from onitama_alphazero_b200 import _lib as L
from onitama_alphazero_b200.net import ConvResNet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else "f16"   # f16 | tf32 | f32 (split operands, f32-faithful)
tf32 = prec == "tf32"
torch.manual_seed(7)
model = ConvResNet(64, 21, blocks)
with torch.no_grad():  # lively activations and non-trivial BatchNorm statistics so that every path is exercised
    for m in model.modules():
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
            m.bias.normal_(0, 0.1)
        if isinstance(m, torch.nn.Linear):
            m.weight.normal_(0, 2.0 / m.in_features ** 0.5)
            m.bias.normal_(0, 0.3)
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.uniform_(0.7, 1.3)
            m.bias.normal_(0.1, 0.2)
model.eval()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
with onb.Context(n, seed=3, mcts_max_sims=4) as ctx:
    ctx.reset()
    for s in range(7):
        ctx.step_random(s, auto_reset=True, out_flags=onb.OUT_PLANES)
    planes = ctx.tensor(L.BUF_PLANES).clone()
    ctx.net_load(model, precision=prec)
    ctx.net_forward(L.BUF_PLANES)
    ctx.sync()
    pol = ctx.tensor(L.BUF_POLICY).clone().cpu().numpy().reshape(n, 50)
    val = ctx.tensor(L.BUF_VALUE).clone().cpu().numpy().reshape(n)
    with torch.no_grad():
        p_ref, v_ref = model.cuda()(planes.reshape(n, 21, 5, 5))
    p_ref = p_ref.reshape(n, 50).cpu().numpy(); v_ref = v_ref.reshape(n).cpu().numpy()
    dp = np.abs(pol - p_ref); dv = np.abs(val - v_ref)
    print(prec, "n", n, "blocks", blocks, "max|dp|", dp.max(), "mean|dp|", dp.mean(), "max|dv|", dv.max(), "mean|dv|", dv.mean())
    print("policy sums", pol.sum(1).min(), pol.sum(1).max(), "ref p range", p_ref.min(), p_ref.max(), "v range", v_ref.min(), v_ref.max())
    print("variation across boards: policy std max %.4f, value std %.4f" % (p_ref.std(0).max(), v_ref.std()))
    bad = np.argwhere(dp > 5e-3)
    print("boards with |dp|>5e-3:", len(set(bad[:, 0].tolist())), "first", bad[:5].tolist())
    ts = ctx.torch_stream()
    with torch.cuda.stream(ts):
        for rep in range(2):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(ts)
            for _ in range(10):
                ctx.net_forward(L.BUF_PLANES)
            e1.record(ts); ts.synchronize()
            ms = e0.elapsed_time(e1) / 10
        print("fused net: %.3f ms per %d positions = %.2fM pos/s, %.1f TFLOP/s" % (ms, n, n / ms / 1e3, n * 11.7e6 / ms / 1e9))
