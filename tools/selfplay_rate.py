"""Lockstep self_play vs self_play_continuous with the on-device network: completed games and samples per second.
python tools/selfplay_rate.py [slots] [sims] [games_factor]"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import onitama_alphazero_b200 as onb
from test_net_cpu import lively_model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 100
factor = int(sys.argv[3]) if len(sys.argv) > 3 else 3
with onb.Context(n, seed=1, mcts_max_sims=sims) as ctx:
    ctx.net_load(lively_model(3))
    onb.self_play(ctx, 2.0, sims, max_plies=4, evaluator=onb.EVAL_NET)   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a = onb.self_play(ctx, 2.0, sims, evaluator=onb.EVAL_NET)
    torch.cuda.synchronize()
    ta = time.perf_counter() - t0
    t0 = time.perf_counter()
    b = onb.self_play_continuous(ctx, 2.0, sims, n_games=factor * n, evaluator=onb.EVAL_NET)
    torch.cuda.synchronize()
    tb = time.perf_counter() - t0
    ctx.self_play_native(2.0, sims, n // 4, evaluator=onb.EVAL_NET)            # warm-up (buffer allocation)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cnat = ctx.self_play_native(2.0, sims, factor * n, evaluator=onb.EVAL_NET)
    torch.cuda.synchronize()
    tc = time.perf_counter() - t0
ma, mb = a["planes"].shape[0], b["planes"].shape[0]
print("lockstep   : %d games, %d samples (%.1f plies/game) in %.2f s -> %.0f games/s, %.0f samples/s" % (n, ma, ma / n, ta, n / ta, ma / ta))
print("continuous : %d games, %d samples (%.1f plies/game) in %.2f s -> %.0f games/s, %.0f samples/s" % (b["games"], mb, mb / b["games"], tb, b["games"] / tb, mb / tb))
mc = cnat["planes"].shape[0]
print("native     : %d games, %d samples (%.1f plies/game) in %.2f s -> %.0f games/s, %.0f samples/s" % (cnat["games"], mc, mc / cnat["games"], tc, cnat["games"] / tc, mc / tc))
