import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import onitama_alphazero_b200 as onb
n, sims = 1 << 14, 400
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
ctx = onb.Context(n, seed=1, stream=s.cuda_stream, mcts_max_sims=sims, planes=False)
ctx.reset()
for i in range(8): ctx.step_random(i)
st = ctx.get_states(); st["result"] = 0; ctx.set_states(st)
def t(label, reps=5):
    for _ in range(2):
        ctx.mcts_begin(2.0, sims); ctx.mcts_run(onb.EVAL_UNIFORM, sims); ctx.mcts_finish(to_host=False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(reps):
        ctx.mcts_begin(2.0, sims); ctx.mcts_run(onb.EVAL_UNIFORM, sims); ctx.mcts_finish(to_host=False)
    b.record(s); b.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-20s %.2f ms  %.3g sims/s" % (label, ms, n * sims / ms * 1e3), flush=True)
t("eval mode")
ctx.mcts_set_noise(True, 0.25, 0.03, 5)
t("train mode (noise)")
