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
