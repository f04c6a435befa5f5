import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import onitama_alphazero_b200 as onb
decks = [[a, b, c, d, e] for a in range(16) for b in range(a + 1, 16) for c in range(16) if c not in (a, b)
         for d in range(c + 1, 16) if d not in (a, b) for e in range(16) if e not in (a, b, c, d)]
decks = np.array(decks, dtype=np.uint8)
rs = np.random.RandomState(0)
with onb.Context(8, planes=False) as ctx:
    for n, depth in ((1, 6), (64, 6), (1024, 6), (8192, 6), (131040, 4), (131040, 5)):
        sel = decks[rs.choice(len(decks), n, replace=False)] if n < len(decks) else decks
        roots = onb.start_states(sel)
        t0 = time.perf_counter()
        nodes, wins, zero = ctx.perft(roots, depth)
        dt = time.perf_counter() - t0
        tot = int(nodes.sum())
        print("deals %6d depth %d: %.3f s, %d nodes total, %.3e nodes/s, leaves(d)=%d" % (n, depth, dt, tot, tot / dt, int(nodes[:, -1].sum())), flush=True)
