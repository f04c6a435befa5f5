"""Host-side bookkeeping that mirrors the reference (no GPU): Elo update and FightStatistics (elo_rating.rs:53-70,
evaluator.rs:38-110)."""
import math

from onitama_alphazero_b200.selfplay import EloRating, FightStatistics, fight_statistics


def test_elo_change():
    ra, rb = EloRating.elo_change(800.0, 800.0, True)
    assert (ra, rb) == (816.0, 784.0)
    ra, rb = EloRating.elo_change(1000.0, 800.0, False)
    ea = 1.0 / (1.0 + 10.0 ** (2.5e-3 * -200.0))
    assert math.isclose(ra, 1000.0 - 32.0 * ea) and math.isclose(rb, 800.0 + 32.0 * ea)
    assert math.isclose(ra + rb, 1800.0)


def test_fight_statistics_fold():
    # games: A is Red, Red wins; A is Blue, Red wins; A is Red, draw; A is Blue, Blue wins
    st = fight_statistics([1, 1, 0, 2], [True, False, True, False])
    assert st.general == dict(wins=2, loses=1, draws=1)
    assert st.color[0] == dict(wins=1, loses=0, draws=1) and st.color[1] == dict(wins=1, loses=1, draws=0)
    assert st.winrate == 0.5 and st.color_winrate == [0.5, 0.5]
    assert len(st.rating_change_history) == 4 and st.rating_change_history[2][0] == st.rating_change_history[2][1]
    assert math.isclose(st.rating_a + st.rating_b, 1600.0)
    s2 = FightStatistics(900.0, 700.0)
    s2.update(2, 0)
    assert s2.general["loses"] == 1 and s2.rating_a < 900.0


def test_replay_ring_buffer_and_trainer_feed():
    import torch
    from onitama_alphazero_b200.selfplay import ReplayBuffer
    from onitama_alphazero_b200.net import ConvResNet, alphaloss
    rb = ReplayBuffer(10)
    mk = lambda lo, hi: (torch.arange(lo, hi).float().reshape(-1, 1, 1, 1).expand(-1, 21, 5, 5).clone(),
                         torch.full((hi - lo, 2, 25), 1.0 / 50.0), torch.arange(lo, hi).float())
    rb.add(*mk(0, 6))
    assert rb.size == 6 and rb.head == 6
    rb.add(*mk(6, 13))   # wraps: keeps samples 3..12
    assert rb.size == 10 and rb.head == 3
    assert sorted(rb.z.tolist()) == [float(v) for v in range(3, 13)]
    assert (rb.planes[:, 0, 0, 0] == rb.z).all()            # planes and z stay aligned through the wrap
    rb.add(*mk(100, 125))  # more than the capacity at once: the newest 10 survive
    assert sorted(rb.z.tolist()) == [float(v) for v in range(115, 125)]
    g = torch.Generator().manual_seed(0)
    x, pi, z = rb.sample(4, generator=g)
    assert x.shape == (4, 21, 5, 5) and pi.shape == (4, 2, 25) and z.shape == (4, 1) and len(set(z.flatten().tolist())) == 4
    # one SGD step of the reference's loss on a minibatch (train.rs:300-325): the loss decreases on the same batch
    torch.manual_seed(0)
    net = ConvResNet(16, 21, 1)
    opt = torch.optim.SGD(net.parameters(), lr=0.05)
    zt = torch.tanh(z / 100.0)
    losses = []
    for _ in range(5):
        p, v = net(x)
        vl, pl = alphaloss(v, p, pi, zt)
        opt.zero_grad(); (vl + pl).backward(); opt.step()
        losses.append(float(vl + pl))
    assert losses[-1] < losses[0]
