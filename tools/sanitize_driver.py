"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck), sized to finish quickly."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import onitama_alphazero_b200 as onb

n = 333
with onb.Context(n, seed=3, mcts_max_sims=40) as ctx:
    ctx.reset()
    for step in range(12):
        ctx.step_random(step, auto_reset=True, out_flags=onb.OUT_MASKS | onb.OUT_PLANES | onb.OUT_ACTIONS)
    ctx.step_random(12, policy=onb.POLICY_AGENT)
    ctx.choose_random(13)
    ctx.step(None, out_flags=onb.OUT_PLANES)
    ctx.legal_moves(); ctx.legal_masks(); ctx.encode()
    s = ctx.get_states(); ctx.set_states(s)
    ctx.search(2.0, 40)
    ctx.search(2.0, 40, evaluator=onb.EVAL_HASH)
    ctx.search(2.0, 20, fused=False)
    ctx.search(2.0, 20, evaluator=onb.EVAL_HASH, fused=False)
    ctx.mcts_set_noise(True, 0.25, 0.03, 1)
    ctx.search(2.0, 30)
    ctx.search(2.0, 10, fused=False)
    ctx.mcts_set_noise(False)
    ctx.mcts_play_best(out_flags=onb.OUT_MASKS)
    ctx.mcts_dump_tree(5)
    ctx.playout(max_plies=40)
    nodes, wins, zero = ctx.perft(onb.start_states([[1, 2, 0, 3, 11], [4, 3, 1, 0, 2]]), 4)
    assert nodes[0].tolist() == [10, 90, 954, 11132]
print("sanitize driver ok")
