"""Pins oracle/onb_oracle.cpp (the CPU restatement) against
  (1) the reference's own known-answer tests, transcribed into tests/golden/golden.json with file:line cites,
  (2) an independent pure-Python restatement (tests/golden/gen_golden.py), and
  (3) the SURVEY.md Appendix-A values.
CPU only.
"""
import json
import math
import os
import struct
import zlib

import numpy as np
import pytest

import oracle_lib as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
REF = G["reference_tests"]
NAMES = [c["name"] for c in G["cards"]]


def sq(rc):
    return rc[0] * 5 + rc[1]


# ------------------------------------------------------------------ bit layout (common/mod.rs:82-134)
def test_bit_layout_from_2d():
    for (rc, expect) in REF["from_2d_to_bitboard"]:
        assert (0x80000000 >> sq(rc)) == expect
    g = O.new_games(1, deck=[4, 3, 1, 0, 2])
    # start position constants (state.rs:24-45)
    assert list(g["pawns"][0]) == REF["start_position"]["pawns"]
    assert list(g["kings"][0]) == REF["start_position"]["kings"]


def test_get_bit_order_through_the_encoder():
    """common/mod.rs:82-92 (get_bit, i = 0 is the MSB) pinned through get_bit_array / create_tensor_from_state"""
    gb = REF["get_bit"]
    g = O.new_games(1, deck=[4, 3, 1, 0, 2])
    g["pawns"][0][0] = gb["bits"] & 0xFFFFFF80
    plane0 = O.encode(g)[0, 0].reshape(25)
    assert plane0.tolist() == [float(b) for b in gb["expected"][:25]]


# ------------------------------------------------------------------ attack maps (card.rs:476-604)
def test_attack_maps_match_golden_and_survey():
    att = O.attack_maps()
    assert att.tolist() == G["attack_maps"]
    crc = "%08x" % zlib.crc32(struct.pack("<800I", *att.tolist()))
    assert crc == G["attack_maps_crc32"] == G["survey_appendix_a"]["attack_maps_crc32"]
    assert int(att.astype(np.uint64).sum()) == G["survey_appendix_a"]["attack_maps_sum"]
    assert sum(bin(int(x)).count("1") for x in att) == G["survey_appendix_a"]["attack_maps_popcount"]


def test_card_colour_stamps():
    for i, c in enumerate(G["cards"]):
        assert O.lib().orc_card_color(i) == c["color"]


# ------------------------------------------------------------------ opening move sets (state.rs:420-492)
@pytest.mark.parametrize("case", REF["opening_moves"], ids=lambda c: "%s-%s-card%d" % (c["cite"], c["color"], c["card"]))
def test_opening_moves(case):
    g = O.new_games(1, deck=case["deck"])
    got = sorted((O.decode_action(a)["frm"], O.decode_action(a)["to"], O.decode_action(a)["piece"])
                 for a in O.gen_moves_card(g, case["color"], case["card"]))
    exp = sorted((sq(m[0]), sq(m[1]), m[2]) for m in case["moves"])
    assert got == exp


# ------------------------------------------------------------------ make_move (state.rs:495-816)
@pytest.mark.parametrize("case", REF["make_move"], ids=lambda c: c["cite"])
def test_make_move_reference_cases(case):
    g = O.new_games(1, deck=case["deck"])
    for k, v in case["set"].items():
        g[k[:-1]][0][int(k[-1])] = v
    g["side"][0] = case["color"]
    frm, to, piece = case["mov"]
    action = to | (frm << 5) | (case["card_idx"] << 10) | (piece << 12)
    r = O.make_move(g, action)
    assert O.RESULT_NAMES[r] == case["result"]
    for (field, color, n, bit) in case["bits"]:
        assert (int(g[field][0][color]) >> (31 - n)) & 1 == bit
    assert int(g["cards"][0][4]) == case["neutral"]  # card rotation (deck.rs:87-90)
    for (field, color, value) in case.get("equals", []):
        assert int(g[field][0][color]) == value
    for (field, color) in case.get("nonzero", []):
        assert int(g[field][0][color]) > 0
    assert int(g["side"][0]) == 1 - case["color"]


# ------------------------------------------------------------------ no-legal-move positions (state.rs:819-889)
@pytest.mark.parametrize("case", REF["no_moves"], ids=lambda c: c["cite"])
def test_no_legal_moves(case):
    g = O.make_state(case["deck"], pawns=case["pawns"], kings=case["kings"], side=case["color"])
    for card in case["cards"]:
        assert len(O.gen_moves_card(g, case["color"], card)) == 0
    if len(case["cards"]) == 2:
        assert len(O.gen_moves(g)) == 0
        assert O.legal_masks(g).tolist() == [[0, 0]]


# ------------------------------------------------------------------ expansion order (ai/mcts/mcts_arena.rs:403-457)
def _fmt(g_cards, a):
    d = O.decode_action(a)
    return "%s %s-%s" % (NAMES[g_cards[d["card_idx"]]], O.sq_name(d["frm"]), O.sq_name(d["to"]))


def test_expand_order_strings():
    case = REF["expand_order"]
    g = O.new_games(1, deck=case["deck"])
    cards = list(g["cards"][0])
    assert [_fmt(cards, a) for a in O.gen_moves(g, 0)] == case["root"]
    # the reference expands child 1 with the ROOT state and the child's colour (Blue)
    assert [_fmt(cards, a) for a in O.gen_moves(g, 1)] == case["child_of_first_with_root_state"]


# ------------------------------------------------------------------ perft vs independent restatement and survey
@pytest.mark.parametrize("case", G["perft"], ids=lambda c: ",".join(map(str, c["deck"])))
def test_perft_matches_python_restatement(case):
    g = O.new_games(1, deck=case["deck"])
    assert int(g["side"][0]) == case["first"]
    depth = len(case["nodes"])
    nodes, wins, zero = O.perft(g, depth)
    assert nodes.tolist() == case["nodes"]
    assert wins.tolist() == case["wins"]
    assert zero.sum() == 0


def test_perft_depth6_matches_survey_appendix():
    for key, v in G["survey_appendix_a"]["perft"].items():
        deck = [int(x) for x in key.split(",")]
        nodes, wins, zero = O.perft(O.new_games(1, deck=deck), 6)
        leaves = (nodes - wins).tolist()
        assert leaves == v["leaves"]
        assert np.cumsum(wins).tolist() == v["cum_wins"]
        assert zero.sum() == 0


# ------------------------------------------------------------------ RNG, deals, playouts (project-defined RNG, two restatements)
def test_rng_known_answers():
    for r in G["rng"]:
        assert O.lib().orc_rand_u32(r["seed"], r["game"], r["step"], r["draw"]) == r["value"]
    for d in G["deals"]:
        out = np.zeros(5, dtype=np.uint8)
        O.lib().orc_deal(d["seed"], d["game"], d["epoch"], out.ctypes.data)
        assert out.tolist() == d["deck"]


def _check_playouts(rows, seed, deck):
    for row in rows:
        g, plies, trace, total = O.playout_games(1, seed, game0=row["game"], deck=deck)
        assert int(plies[0]) == row["plies"]
        assert str(int(trace[0])) == row["trace"]
        assert g["pawns"][0].tolist() == row["pawns"] and g["kings"][0].tolist() == row["kings"]
        assert g["cards"][0].tolist() == row["cards"]
        assert int(g["side"][0]) == row["side"]
        assert int(g["result"][0]) == {0: 1, 1: 2, -1: 0}[row["winner"]]


def test_random_playouts_bit_exact_vs_python():
    _check_playouts(G["playouts"], 2024, None)
    _check_playouts(G["playouts_fixed_deck"], 5, [1, 2, 0, 3, 11])


def test_playout_equals_lockstep_stepping():
    n = 64
    fin, plies, trace, total = O.playout_games(n, 99)
    g = O.new_games(n, seed=99)
    for step in range(int(plies.max())):
        O.env_step_random(g, 99, step)
    assert g.tobytes() == fin.tobytes()


# ------------------------------------------------------------------ encoder (common.rs:26-80)
def test_encoder_vs_python():
    for row in G["encode"]:
        g = O.make_state(row["cards"], pawns=row["pawns"], kings=row["kings"], side=row["side"])
        g["cards"][0] = row["cards"]
        planes = O.encode(g).reshape(-1)
        assert np.flatnonzero(planes == 1.0).tolist() == row["ones"]
        assert set(np.unique(planes)) <= {0.0, 1.0}


def test_encoder_structure():
    g = O.new_games(8, seed=3)
    for s in range(5):
        O.env_step_random(g, 3, s)
    p = O.encode(g)
    for i in range(8):
        side = int(g["side"][i])
        own = g["cards"][i][2 * side: 2 * side + 2]
        card_planes = [k for k in range(16) if p[i, 4 + k].all()]
        assert sorted(card_planes) == sorted(own.tolist())
        assert p[i, 20].all() == (side == 1) and p[i, 20].any() == (side == 1)
        assert p[i, 0].sum() == bin(int(g["pawns"][i][0])).count("1")


# ------------------------------------------------------------------ PUCT arena
@pytest.mark.parametrize("case", G["puct"], ids=lambda c: "%s-c%.3f-%d" % (",".join(map(str, c["deck"])), c["c_puct"], c["sims"]))
def test_puct_vs_python_restatement(case):
    g = O.new_games(1, deck=case["deck"])
    r = O.mcts_search(g, case["c_puct"], case["sims"], evaluator=0, dump=True)
    t = r["tree"]
    kids = list(range(int(t["first_child"][0]), int(t["first_child"][0]) + int(t["n_child"][0])))
    assert t["visits"][kids].tolist() == case["visits"]
    assert r["n_nodes"] == case["n_nodes"]
    assert r["best"] == case["best"]
    assert r["root_q"] == case["root_q"]  # bit-exact f64
    assert t["winrate"][kids].tolist() == case["child_q"]
    assert t["prior"][kids].tolist() == case["child_prior"]
    assert abs(r["mean_depth"] - case["mean_depth"]) < 1e-12
    assert r["pass_seen"] == 0


def test_puct_vs_survey_appendix():
    for case in G["survey_appendix_a"]["puct"]:
        c = {"sqrt2": math.sqrt(2.0), "2.0": 2.0, "5.0": 5.0}[case["c"]]
        r = O.mcts_search(O.new_games(1, deck=case["deck"]), c, case["sims"], dump=True)
        t = r["tree"]
        kids = slice(int(t["first_child"][0]), int(t["first_child"][0]) + int(t["n_child"][0]))
        assert t["visits"][kids].tolist() == case["visits"]
        if "nodes" in case:
            assert r["n_nodes"] == case["nodes"]


def test_puct_tree_invariants():
    g = O.new_games(1, seed=11, game0=5)
    for s in range(6):
        O.env_step_random(g, 11, s, game0=5)
    r = O.mcts_search(g, 2.0, 300, evaluator=1, dump=True)
    t = r["tree"]
    assert t["visits"][0] == 300
    for i in range(r["n_nodes"]):
        if t["n_child"][i]:
            kids = slice(int(t["first_child"][i]), int(t["first_child"][i]) + int(t["n_child"][i]))
            # every playout through an expanded, non-terminal node continues into exactly one child
            assert t["visits"][kids].sum() == t["visits"][i] - 1
            assert (t["parent"][kids] == i).all()
    assert abs(r["pi"].sum() - 1.0) < 1e-6


def test_pass_nodes_defined():
    # state.rs:852-889: Blue to move has no legal move; the reference would panic on the second playout
    case = REF["no_moves"][1]
    g = O.make_state(case["deck"], pawns=case["pawns"], kings=case["kings"], side=case["color"])
    r = O.mcts_search(g, 1.5, 50, dump=True)
    assert r["pass_seen"] == 1
    t = r["tree"]
    assert t["n_child"][0] == 2 and (t["flags"][1:3] & 4).all()
    assert O.decode_action(r["best"])["is_pass"] == 1


# ------------------------------------------------------------------ plain UCT with rollouts (ai/mcts/mcts_arena.rs)
def _bb(r, c):
    return 0x80000000 >> (r * 5 + c)


# the reference's own known-answer tests for the `Mcts` agent (ai/mcts/mcts_arena.rs:459-554): position, parameters, expected move
PLAIN_MCTS_CASES = [
    dict(cite="mcts_arena.rs:459-486 test_best_move_win", deck=[3, 2, 0, 1, 11], side=1, min_visits=5, c=math.sqrt(2.0),
         set={"kings0": _bb(1, 3)}, expect=dict(frm=1, to=8, piece=0, card_idx=3)),
    dict(cite="mcts_arena.rs:488-521 test_no_way_to_hide_for_blue", deck=[12, 8, 3, 11, 1], side=1, min_visits=5, c=2.0,
         set={"kings1": _bb(0, 4), "pawns1": 0, "pawns0": _bb(0, 3) | _bb(1, 4)}, expect=dict(frm=4, to=2, piece=1, card_idx=2)),
    dict(cite="mcts_arena.rs:523-553 test_worst_case_capture_blue", deck=[8, 7, 6, 9, 11], side=1, min_visits=1, c=1.0,
         set={"kings1": _bb(1, 2), "pawns0": _bb(2, 3) | _bb(3, 2)}, expect=dict(frm=7, to=2, piece=1, card_idx=3)),
]


def plain_mcts_root(case):
    g = O.new_games(1, deck=case["deck"])
    for k, v in case["set"].items():
        g[k[:-1]][0][int(k[-1])] = v
    g["side"][0] = case["side"]
    return g


@pytest.mark.parametrize("case", PLAIN_MCTS_CASES, ids=lambda c: c["cite"].split()[-1])
def test_plain_mcts_reference_known_answers(case):
    """5000 playouts as in the reference's tests; the expected move must come out for every RNG stream tried (the reference's
    thread_rng is unseeded, so its tests assert exactly this robustness)."""
    g = plain_mcts_root(case)
    for seed in (1, 2, 3):
        r = O.uct_search_batch(g, case["c"], case["min_visits"], 5000, seed=seed)
        d = O.decode_action(int(r["best"][0]))
        assert {k: d[k] for k in ("frm", "to", "piece", "card_idx")} == case["expect"], (seed, d)
        assert int(r["child_visits"][0].sum()) <= 5000 and r["pass_seen"][0] == 0


def test_plain_mcts_structure():
    """min_node_visits gates expansion (mcts_arena.rs:116-121): the root is expanded DURING playout min + 2 (which still rolls out
    from the root itself), so the children share playouts - (min + 2) visits; unvisited children score +inf and the LAST one wins
    the tie, so the first visits run from the last child back."""
    g = O.new_games(1, deck=[1, 2, 0, 3, 11])
    for min_v, sims in ((5, 7), (5, 12), (0, 5), (5, 400)):
        r = O.uct_search_batch(g, math.sqrt(2.0), min_v, sims, seed=9)
        v = r["child_visits"][0][:10]
        assert int(v.sum()) == max(0, sims - (min_v + 2))
        k = int(v.sum())
        if 0 < k <= 10:
            assert v.tolist() == [0] * (10 - k) + [1] * k
    two = np.concatenate([g, g])
    a = O.uct_search_batch(two, 1.4, 5, 300, seed=4, game0=10)
    assert not np.array_equal(a["child_visits"][0], a["child_visits"][1])     # different game ids -> different rollouts
    b = O.uct_search_batch(two[1:], 1.4, 5, 300, seed=4, game0=11)
    assert np.array_equal(a["child_visits"][1], b["child_visits"][0])           # keyed by the global game id
