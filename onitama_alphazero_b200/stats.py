"""The reference's on-disk training log (alphazero-training/src/stats.rs:14-80): `Stats` collects per-iteration losses, the four arena
results of `Evaluator::pit` (PitStatistics of evaluator.rs:112-118) and the self-play volume, and `save()` writes them as pretty JSON
into `./loss_stats/loss_<YYYYmmdd_HHMMSS>/loss_stats_<ts>.json` with serde's field names, so that the reference's own tooling (and a
`serde_json::from_str::<Stats>`) reads what this framework writes. Host bookkeeping: no device code."""
import datetime
import json
import os


def fight_statistics_dict(st):
    """FightStatistics (evaluator.rs:38-50) as serde serialises it. Accepts selfplay.FightStatistics or the dict Context.fight_stats returns."""
    if isinstance(st, dict):
        general, color, hist = st["general"], st["color"], st.get("rating_change_history", [])
        winrate, cw, ra, rb = st["winrate"], st["color_winrate"], st["rating_a"], st["rating_b"]
    else:
        general, color, hist = st.general, st.color, st.rating_change_history
        winrate, cw, ra, rb = st.winrate, st.color_winrate, st.rating_a, st.rating_b
    wld = lambda d: {"wins": int(d["wins"]), "loses": int(d["loses"]), "draws": int(d["draws"])}
    return {"general": wld(general), "winrate": float(winrate), "color": [wld(color[0]), wld(color[1])],
            "color_winrate": [float(cw[0]), float(cw[1])], "rating_a": float(ra), "rating_b": float(rb),
            "rating_change_history": [{"before_a": float(h[0]), "after_a": float(h[1]), "before_b": float(h[2]), "after_b": float(h[3])}
                                      for h in hist]}


def empty_fight_statistics(rating_a=0.0, rating_b=0.0):
    """FightStatistics::default (an arena that was not run this iteration)"""
    z = {"wins": 0, "loses": 0, "draws": 0}
    return {"general": dict(z), "winrate": 0.0, "color": [dict(z), dict(z)], "color_winrate": [0.0, 0.0], "rating_a": rating_a,
            "rating_b": rating_b, "rating_change_history": []}


class Stats:
    """stats.rs:14-80"""

    def __init__(self, root="./loss_stats", now=None):
        now = now or datetime.datetime.now()
        self.iteration, self.loss, self.policy_loss, self.value_loss = [], [], [], []
        self.was_best_change, self.fight_statistics, self.games_played = [], [], []
        self.dir = os.path.join(root, "loss_%s" % now.strftime("%Y%m%d_%H%M%S"))   # stats.rs:27-31

    def push_games_played(self, games_amnt, positions_amnt):                        # stats.rs:45-50
        self.games_played.append({"games_amnt": int(games_amnt), "positions_retrieved": int(positions_amnt)})

    def push(self, epoch, loss, value_loss, policy_loss):                           # stats.rs:52-57
        self.iteration.append(int(epoch)); self.loss.append(float(loss))
        self.policy_loss.append(float(policy_loss)); self.value_loss.append(float(value_loss))

    def push_fight(self, was_best_change, self_fight=None, random_fight=None, alphabeta_fight=None, mcts_fight=None):  # stats.rs:59-62
        f = lambda s: empty_fight_statistics() if s is None else fight_statistics_dict(s)
        self.was_best_change.append(bool(was_best_change))
        self.fight_statistics.append({"self_fight": f(self_fight), "random_fight": f(random_fight), "alphabeta_fight": f(alphabeta_fight),
                                      "mcts_fight": f(mcts_fight)})

    def to_dict(self):
        return {"iteration": self.iteration, "loss": self.loss, "policy_loss": self.policy_loss, "value_loss": self.value_loss,
                "was_best_change": self.was_best_change, "fight_statistics": self.fight_statistics, "games_played": self.games_played,
                "dir": self.dir}

    def get_filename(self, now=None):                                               # stats.rs:64-68 (the reference's format string, typo included)
        now = now or datetime.datetime.now()
        return "loss_stats_%s.json" % now.strftime("%Y%m%y_%H%M%S")

    def save(self, now=None):                                                       # stats.rs:70-80: serde_json::to_string_pretty
        os.makedirs(self.dir, exist_ok=True)
        path = os.path.join(self.dir, self.get_filename(now))
        with open(path, "w") as f:
            json.dump(self.to_dict(), f, indent=2)
        return path

    @classmethod
    def load(cls, path):
        d = json.load(open(path))
        s = cls.__new__(cls)
        for k in ("iteration", "loss", "policy_loss", "value_loss", "was_best_change", "fight_statistics", "games_played", "dir"):
            setattr(s, k, d[k])
        return s
