#!/usr/bin/env python3
"""One AlphaZero iteration on the GPU hot path, shaped like the reference's train() (alphazero-training/src/train.rs:158-412):

    self-play (train-mode noise, network-guided PUCT, all games in lockstep)  ->  replay ring buffer
    ->  a few SGD minibatches with alphaloss  ->  arena: the trained network against the `Random` agent.

Everything game- and search-related runs in libonb.so; the network, optimiser and bookkeeping are plain PyTorch / Python
(out of scope of the hot path). Usage: python examples/selfplay_train_loop.py [--games 256] [--sims 64] [--iters 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import onitama_alphazero_b200 as onb
from onitama_alphazero_b200.net import ConvResNet, alphaloss, make_evaluator


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=256)
    ap.add_argument("--sims", type=int, default=64)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--max-plies", type=int, default=40)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--sgd-steps", type=int, default=20)
    ap.add_argument("--eval-games", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)

    torch.manual_seed(args.seed)
    model = ConvResNet(64, 21, 3).cuda()
    opt = torch.optim.SGD(model.parameters(), lr=1e-2, weight_decay=1e-4)  # train.rs:181-186
    replay = onb.ReplayBuffer(200_000, device="cuda")
    log = []
    for it in range(args.iters):
        # ---- self-play (train.rs:218-245): train-mode root noise, the current network as evaluator
        net = make_evaluator(model)
        with onb.Context(args.games, seed=args.seed + it, mcts_max_sims=args.sims) as ctx:
            ctx.mcts_set_noise(True, 0.25, 0.03, args.seed + 1000 * it)
            data = onb.self_play(ctx, 2.0, args.sims, max_plies=args.max_plies, net=net)
        replay.add(data["planes"], data["pi"], data["z"])
        # ---- training (train.rs:264-339)
        model.train()
        losses = []
        for _ in range(args.sgd_steps):
            x, pi, z = replay.sample(args.batch)
            p, v = model(x)
            vl, pl = alphaloss(v, p, pi, z)
            opt.zero_grad()
            (vl + pl).backward()
            opt.step()
            losses.append((float(vl), float(pl)))
        model.eval()
        # ---- evaluation (evaluator.rs:195-239): the network-guided search against the Random agent, colours alternate
        net = make_evaluator(model)
        a_is_red = (np.arange(args.eval_games) % 2) == 0
        counter = {"i": 0}
        with onb.Context(args.eval_games, seed=10_000 + it, mcts_max_sims=args.sims, planes=False) as ctx:
            ctx.reset()

            def az(cx):
                cx.search_device(2.0, args.sims, net=net)
                cx.tensor(onb.BUF_ACTIONS).copy_(cx.tensor(onb.BUF_BEST))

            def rnd(cx):
                cx.choose_random(counter["i"], policy=onb.POLICY_AGENT)
                counter["i"] += 1

            w, l, d = onb.fight(ctx, az, rnd, a_is_red, max_plies=150)
            stats = onb.fight_statistics(ctx.last_fight_results, a_is_red)
        log.append(dict(iteration=it, samples=int(data["planes"].shape[0]), replay=replay.size, value_loss=losses[-1][0],
                        policy_loss=losses[-1][1], wins=w, losses=l, draws=d, elo=stats.rating_a))
        print(log[-1], flush=True)
    return log


if __name__ == "__main__":
    main()
