#!/usr/bin/env python3
"""Generate tests/golden/golden.json.

Run in the BUILD container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/gen_golden.py

The Rust reference cannot be executed here (no cargo/rustc), so this script is an INDEPENDENT
pure-Python restatement used to cross-check oracle/onb_oracle.cpp. It deliberately uses a different
formulation from the oracle: the card hex constants are re-parsed from the reference's
onitama-game/src/game/card.rs at run time and decoded into (d_row, d_col) offsets, and positions are
sets of squares, not shifted bitboards. Everything the reference pins with a known-answer test is also
written into the JSON verbatim from the reference's test sources (cited per entry) so the tests can
check the oracle and the CUDA path against the reference's own expectations.

Cited reference files (relative to the reference root):
  onitama-game/src/game/card.rs, state.rs, deck.rs, common/mod.rs, ai/mcts/mcts_arena.rs,
  alphazero-training/src/alphazero_mcts/mcts_arena.rs, alphazero-training/src/common.rs
"""
import json
import math
import os
import re
import struct
import sys
import zlib

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")

RED, BLUE = 0, 1
PAWN, KING = 0, 1
NAMES = ["Tiger", "Dragon", "Frog", "Rabbit", "Crab", "Elephant", "Goose", "Rooster", "Monkey",
         "Mantis", "Crane", "Horse", "Ox", "Boar", "Eel", "Cobra"]  # card.rs:471-474


# ----------------------------------------------------------------------------- parse card.rs
def parse_cards():
    src = open(os.path.join(REF, "onitama-game/src/game/card.rs")).read()
    blocks = re.findall(r"pub const ([A-Z]+): Card = Card \{(.*?)\n\};", src, re.S)
    raw = {}
    for name, body in blocks:
        # strip comments
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        body = re.sub(r"//[^\n]*", "", body)
        pos = re.search(r"positions:\s*([^,]+),", body).group(1).strip()
        mir = re.search(r"mirror:\s*([^,]+),", body).group(1).strip()
        col = re.search(r"player_color:\s*PlayerColor::(\w+)", body).group(1)
        idx = int(re.search(r"index:\s*(\d+)", body).group(1))
        raw[name] = (pos, mir, col, idx)

    def resolve(expr):
        if expr.startswith("0x"):
            return int(expr.replace("_", ""), 16)
        other, field = expr.split(".")
        return resolve(raw[other][0 if field == "positions" else 1])

    cards = [None] * 16
    for name, (pos, mir, col, idx) in raw.items():
        cards[idx] = dict(name=name.capitalize(), positions=resolve(pos), mirror=resolve(mir),
                          color=RED if col == "Red" else BLUE, index=idx)
    assert all(c is not None for c in cards)
    assert [c["name"] for c in cards] == NAMES
    return cards


CARDS = parse_cards()


def mask_to_offsets(mask):
    """bit (31-n) set -> square n -> offset from the centre square 12."""
    offs = []
    for n in range(25):
        if (mask >> (31 - n)) & 1:
            offs.append((n // 5 - 2, n % 5 - 2))
    return offs


OFFS = [[mask_to_offsets(c["positions"]) for c in CARDS], [mask_to_offsets(c["mirror"]) for c in CARDS]]


def attack_mask(color, card, frm):
    r, c = divmod(frm, 5)
    m = 0
    for dr, dc in OFFS[color][card]:
        rr, cc = r + dr, c + dc
        if 0 <= rr < 5 and 0 <= cc < 5:
            m |= 1 << (31 - (rr * 5 + cc))
    return m


# ----------------------------------------------------------------------------- rules on square sets
class Pos:
    """pieces: dict square -> (color, kind); deck: list of 5 card ids; mirrors state.rs semantics for LEGAL play."""
    __slots__ = ("pc", "deck")

    def __init__(self, deck, pc=None):
        self.deck = list(deck)
        if pc is None:
            pc = {}
            for s in (0, 1, 3, 4):
                pc[s] = (BLUE, PAWN)
            pc[2] = (BLUE, KING)
            for s in (20, 21, 23, 24):
                pc[s] = (RED, PAWN)
            pc[22] = (RED, KING)
        self.pc = pc

    def copy(self):
        return Pos(self.deck, dict(self.pc))

    def boards(self):
        b = {(RED, PAWN): 0, (RED, KING): 0, (BLUE, PAWN): 0, (BLUE, KING): 0}
        for s, ck in self.pc.items():
            b[ck] |= 1 << (31 - s)
        return [b[(RED, PAWN)], b[(BLUE, PAWN)]], [b[(RED, KING)], b[(BLUE, KING)]]

    def moves(self, color):
        """(slot, from, to, kind) in the reference order: slot asc, from asc, to asc (state.rs:301-378)."""
        out = []
        base = 0 if color == RED else 2
        for slot in (base, base + 1):
            card = self.deck[slot]
            for frm in range(25):
                ck = self.pc.get(frm)
                if ck is None or ck[0] != color:
                    continue
                r, c = divmod(frm, 5)
                tos = []
                for dr, dc in OFFS[color][card]:
                    rr, cc = r + dr, c + dc
                    if 0 <= rr < 5 and 0 <= cc < 5:
                        t = rr * 5 + cc
                        o = self.pc.get(t)
                        if o is None or o[0] != color:
                            tos.append(t)
                for t in sorted(tos):
                    out.append((slot, frm, t, ck[1]))
        return out

    def make_move(self, color, slot, frm, to, kind):
        """returns 'red'/'blue'/None (state.rs:145-202 for legal moves)."""
        res = None
        tgt = self.pc.get(to)
        if tgt is not None and tgt[1] == KING:
            res = color
        del self.pc[frm]
        self.pc[to] = (color, kind)
        if kind == KING and ((color == RED and to == 2) or (color == BLUE and to == 22)):
            res = color
        self.deck[slot], self.deck[4] = self.deck[4], self.deck[slot]
        return res

    def winner(self):
        """state.rs:120-134"""
        kings = {c: s for s, (c, k) in self.pc.items() if k == KING}
        if RED not in kings or kings.get(BLUE) == 22:
            return BLUE
        if BLUE not in kings or kings.get(RED) == 2:
            return RED
        return None


def first_mover(deck):
    return CARDS[deck[4]]["color"]  # game_state.rs:34-41


def action_code(slot, frm, to, kind):
    return to | (frm << 5) | (slot << 10) | (kind << 12)


# ----------------------------------------------------------------------------- counter RNG (project-defined)
M64 = (1 << 64) - 1


def mix64(z):
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return z


def rand_u32(seed, game, step, draw):
    h = mix64((seed + 0x9E3779B97F4A7C15 * (game + 1)) & M64)
    h = mix64(h ^ ((step << 32) | draw))
    return h >> 32


def rand_index(r, n):
    return (r * n) >> 32


def deal(seed, game, epoch):
    ids = list(range(16))
    for i in range(5):
        j = i + rand_index(rand_u32(seed, game, epoch, 8 + i), 16 - i)
        ids[i], ids[j] = ids[j], ids[i]
    return ids[:5]


# ----------------------------------------------------------------------------- perft
def perft(pos, color, depth):
    nodes = [0] * depth
    wins = [0] * depth

    def rec(p, col, d):
        for (slot, frm, to, kind) in p.moves(col):
            ch = p.copy()
            w = ch.make_move(col, slot, frm, to, kind)
            nodes[d] += 1
            if w is not None:
                wins[d] += 1
                continue
            if d + 1 < depth:
                rec(ch, col ^ 1, d + 1)

    rec(pos, color, 0)
    return nodes, wins


# ----------------------------------------------------------------------------- encoder (common.rs:26-80)
def encode(pos, color):
    out = [0.0] * 525
    pawns, kings = pos.boards()
    for p, b in enumerate([pawns[RED], kings[RED], pawns[BLUE], kings[BLUE]]):
        for i in range(25):
            out[p * 25 + i] = float((b >> (31 - i)) & 1)
    base = 0 if color == RED else 2
    for slot in (base, base + 1):
        for i in range(25):
            out[(4 + pos.deck[slot]) * 25 + i] = 1.0
    if color == BLUE:
        for i in range(25):
            out[20 * 25 + i] = 1.0
    return out


# ----------------------------------------------------------------------------- PUCT arena (alphazero_mcts/mcts_arena.rs)
def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def total_key(x):
    b = struct.unpack("q", struct.pack("d", x))[0]
    if b < 0:
        b ^= 0x7FFFFFFFFFFFFFFF
    return b


class Node:
    __slots__ = ("parent", "children", "mov", "visits", "reward", "winrate", "terminal", "expanded", "color", "prob")

    def __init__(self, parent, mov, color, prob):
        self.parent, self.mov, self.color, self.prob = parent, mov, color, prob
        self.children = []
        self.visits, self.reward, self.winrate = 0, 0.0, 0.0
        self.terminal = self.expanded = False


def puct_search(pos, color, c_puct, sims, policy_fn):
    arena = [Node(None, None, color, 1.0)]
    depth_sum = 0
    for _ in range(sims):
        p = pos.copy()
        col = color
        ni = 0
        while arena[ni].expanded and not arena[ni].terminal:
            par = arena[ni]
            best = None
            bestk = None
            sq = math.sqrt(float(par.visits))
            for ci in par.children:
                ch = arena[ci]
                u = ch.winrate + c_puct * ch.prob * (sq / float(ch.visits + 1))
                k = total_key(u)
                if bestk is None or k >= bestk:
                    best, bestk = ci, k
            ni = best
            depth_sum += 1
            slot, frm, to, kind = arena[ni].mov
            w = p.make_move(arena[arena[ni].parent].color, slot, frm, to, kind)
            col ^= 1
            if w is not None:
                arena[ni].terminal = True
        node = arena[ni]
        winner = p.winner()
        value = 0.0
        if not node.expanded and not node.terminal:
            policy, value = policy_fn(p, col)
            moves = p.moves(col)
            pri = [[0.0] * 25, [0.0] * 25]
            for (slot, frm, to, kind) in moves:
                pri[slot & 1][to] = float(policy[(slot & 1) * 25 + to])
            for c in range(2):
                s = 0.0
                for v in pri[c]:
                    s += v
                if s > 0.0:
                    pri[c] = [v / s for v in pri[c]]
            for (slot, frm, to, kind) in moves:
                arena.append(Node(ni, (slot, frm, to, kind), node.color ^ 1, pri[slot & 1][to]))
                node.children.append(len(arena) - 1)
            node.expanded = True
        par_idx = node.parent if node.parent is not None else 0
        rc = arena[par_idx].color
        if winner is not None:
            r = 1.0 if winner == rc else -1.0
        else:
            r = value
        n = ni
        while True:
            nd = arena[n]
            nd.visits += 1
            nd.reward += r
            nd.winrate = nd.reward / float(nd.visits)
            if nd.parent is None:
                break
            n = nd.parent
            r = -r
    root = arena[0]
    best = None
    for ci in root.children:
        a = arena[ci].visits / float(root.visits)
        if best is None or total_key(a) >= total_key(arena[best].visits / float(root.visits)):
            best = ci
    return arena, best, depth_sum / sims


def uniform_policy(p, col):
    return [f32(1.0 / 50.0)] * 50, 0.0


# ----------------------------------------------------------------------------- random playout (policy a10)
def playout(seed, game, deck=None, max_plies=100000):
    d = deck if deck is not None else deal(seed, game, 0)
    pos = Pos(d)
    col = first_mover(d)
    ply = 0
    trace = 0
    winner = None
    while winner is None and ply < max_plies:
        mv = pos.moves(col)
        if not mv:
            slot = (0 if col == RED else 2) + rand_index(rand_u32(seed, game, ply, 1), 2)
            pos.deck[slot], pos.deck[4] = pos.deck[4], pos.deck[slot]
            a = (slot << 10) | (1 << 13)
        else:
            slot, frm, to, kind = mv[rand_index(rand_u32(seed, game, ply, 0), len(mv))]
            winner = pos.make_move(col, slot, frm, to, kind)
            a = action_code(slot, frm, to, kind)
        trace = mix64(trace ^ a)
        col ^= 1
        ply += 1
    pawns, kings = pos.boards()
    return dict(game=game, deck0=d, plies=ply, winner=(-1 if winner is None else winner), trace=str(trace),
                pawns=pawns, kings=kings, cards=pos.deck, side=col)


# ----------------------------------------------------------------------------- main
def main():
    quick = "--quick" in sys.argv
    g = {"_generated_by": "tests/golden/gen_golden.py (independent pure-Python restatement; Rust reference not runnable here)"}

    # card table + attack maps
    g["cards"] = [dict(name=c["name"], positions=c["positions"], mirror=c["mirror"], color=c["color"]) for c in CARDS]
    att = [[[attack_mask(col, card, frm) for frm in range(25)] for card in range(16)] for col in range(2)]
    flat = [att[c][k][f] for c in range(2) for k in range(16) for f in range(25)]
    g["attack_maps"] = flat
    g["attack_maps_crc32"] = "%08x" % zlib.crc32(struct.pack("<800I", *flat))
    g["attack_maps_sum"] = sum(flat)
    g["attack_maps_popcount"] = sum(bin(x).count("1") for x in flat)

    # --- reference known-answer tests, transcribed from the reference's own test sources ---
    ref = {}
    # common/mod.rs:95-134
    ref["from_2d_to_bitboard"] = [[[0, 0], 0x80000000], [[2, 1], 0x00100000], [[4, 4], 0x00000080]]
    # common/mod.rs:82-92: get_bit(bits, i) for i in 0..32, i = 0 is the MSB
    ref["get_bit"] = dict(cite="common/mod.rs:82-92", bits=0b00001111000011110000111100001111,
                          expected=[0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1])
    # state.rs:420-492 (moves given as [[r0,c0],[r1,c1],piece])
    ref["opening_moves"] = [
        dict(cite="state.rs:420-455", deck=[4, 3, 1, 0, 2], color=RED, card=4,
             moves=[[[4, 0], [3, 0], PAWN], [[4, 1], [3, 1], PAWN], [[4, 2], [3, 2], KING], [[4, 3], [3, 3], PAWN], [[4, 4], [3, 4], PAWN]]),
        dict(cite="state.rs:420-455", deck=[4, 3, 1, 0, 2], color=RED, card=3,
             moves=[[[4, 0], [3, 1], PAWN], [[4, 1], [3, 2], PAWN], [[4, 2], [3, 3], KING], [[4, 3], [3, 4], PAWN]]),
        dict(cite="state.rs:457-492", deck=[1, 0, 4, 3, 2], color=BLUE, card=4,
             moves=[[[0, 0], [1, 0], PAWN], [[0, 1], [1, 1], PAWN], [[0, 2], [1, 2], KING], [[0, 3], [1, 3], PAWN], [[0, 4], [1, 4], PAWN]]),
        dict(cite="state.rs:457-492", deck=[1, 0, 4, 3, 2], color=BLUE, card=3,
             moves=[[[0, 1], [1, 0], PAWN], [[0, 2], [1, 1], KING], [[0, 3], [1, 2], PAWN], [[0, 4], [1, 3], PAWN]]),
    ]
    # state.rs:495-816: deck CRAB,RABBIT,DRAGON,TIGER,FROG = [4,3,1,0,2]; overrides before the move; expectations after
    D = [4, 3, 1, 0, 2]
    SP = dict(pawns=[0x00000D80, 0xD8000000], kings=[0x00000200, 0x20000000])
    ref["make_move"] = [
        dict(cite="state.rs:495-520", deck=D, set={}, color=RED, mov=[20, 15, PAWN], card_idx=0, result="InProgress",
             bits=[["pawns", RED, 15, 1], ["pawns", RED, 20, 0]], neutral=4),
        dict(cite="state.rs:522-547", deck=D, set={}, color=BLUE, mov=[1, 11, PAWN], card_idx=3, result="InProgress",
             bits=[["pawns", BLUE, 11, 1], ["pawns", BLUE, 1, 0]], neutral=0),
        dict(cite="state.rs:549-593", deck=D, set={"pawns1": 0x58010000}, color=RED, mov=[20, 15, PAWN], card_idx=0, result="Capture",
             bits=[["pawns", RED, 15, 1], ["pawns", RED, 20, 0]], neutral=4, equals=[["pawns", BLUE, 0x58000000]]),
        dict(cite="state.rs:595-639", deck=D, set={"pawns1": 0x58200000}, color=BLUE, mov=[10, 20, PAWN], card_idx=3, result="Capture",
             bits=[["pawns", BLUE, 20, 1], ["pawns", BLUE, 10, 0]], neutral=0, equals=[["pawns", RED, 0x00000580]]),
        dict(cite="state.rs:641-685", deck=D, set={"pawns0": 0x01000680}, color=RED, mov=[7, 2, PAWN], card_idx=0, result="RedWin",
             bits=[["pawns", RED, 2, 1], ["pawns", RED, 7, 0]], neutral=4, equals=[["kings", BLUE, 0]]),
        dict(cite="state.rs:687-731", deck=D, set={"pawns1": 0x58080000}, color=BLUE, mov=[12, 22, PAWN], card_idx=3, result="BlueWin",
             bits=[["pawns", BLUE, 22, 1], ["pawns", BLUE, 12, 0]], neutral=0, equals=[["kings", RED, 0]]),
        dict(cite="state.rs:733-774", deck=D, set={"kings1": 0x00080000, "kings0": 0x00002000}, color=BLUE, mov=[12, 22, KING], card_idx=3,
             result="BlueWin", bits=[["kings", BLUE, 22, 1], ["kings", BLUE, 12, 0]], neutral=0, nonzero=[["kings", RED]]),
        dict(cite="state.rs:776-816", deck=D, set={"kings1": 0x00080000, "kings0": 0x01000000}, color=RED, mov=[7, 2, KING], card_idx=0,
             result="RedWin", bits=[["kings", RED, 2, 1], ["kings", RED, 7, 0]], neutral=4, nonzero=[["kings", BLUE]]),
    ]
    ref["start_position"] = SP
    # state.rs:819-889
    ref["no_moves"] = [
        dict(cite="state.rs:819-850", deck=[1, 0, 3, 11, 2], kings=[512, 67108864], pawns=[61568, 3221225472], color=BLUE, cards=[3]),
        dict(cite="state.rs:852-889", deck=[1, 3, 0, 11, 2], kings=[16384, 131072], pawns=[2148009984, 138416256], color=BLUE, cards=[0, 11]),
    ]
    # ai/mcts/mcts_arena.rs:403-457: deck Dragon,Frog,Tiger,Rabbit,Horse; child order after expand
    ref["expand_order"] = dict(
        cite="onitama-game/src/ai/mcts/mcts_arena.rs:403-457", deck=[1, 2, 0, 3, 11],
        root=["Dragon a1-c2", "Dragon b1-d2", "Dragon c1-a2", "Dragon c1-e2", "Dragon d1-b2", "Dragon e1-c2",
              "Frog b1-a2", "Frog c1-b2", "Frog d1-c2", "Frog e1-d2"],
        # second expand is done with the ROOT state (not the child's) but the child's colour (Blue)
        child_of_first_with_root_state=["Tiger a5-a3", "Tiger b5-b3", "Tiger c5-c3", "Tiger d5-d3", "Tiger e5-e3",
                                         "Rabbit b5-a4", "Rabbit c5-b4", "Rabbit d5-c4", "Rabbit e5-d4"])
    g["reference_tests"] = ref

    # --- independent restatement outputs ---
    # perft
    decks = [[1, 2, 0, 3, 11], [4, 3, 1, 0, 2], [0, 1, 2, 3, 4]]
    depth = 4 if quick else 5
    g["perft"] = []
    for d in decks:
        n, w = perft(Pos(d), first_mover(d), depth)
        g["perft"].append(dict(deck=d, first=first_mover(d), nodes=n, wins=w))
        print("perft", d, n, w, flush=True)
    # SURVEY.md Appendix A (values produced by the surveyor's own throwaway restatement; kept as a third witness)
    g["survey_appendix_a"] = dict(
        perft={"1,2,0,3,11": dict(leaves=[10, 90, 949, 11019, 125967, 1638008], cum_wins=[0, 0, 5, 118, 2431, 32844]),
               "4,3,1,0,2": dict(leaves=[9, 99, 986, 11883, 139068, 1747555], cum_wins=[0, 0, 4, 130, 2515, 35610]),
               "0,1,2,3,4": dict(leaves=[8, 88, 1008, 11266, 136468, 1735302], cum_wins=[0, 0, 4, 120, 2462, 33888])},
        all_deals=dict(sum_perft1=1375920, sum_perft2=14375088, n_deals=131040),
        attack_maps_crc32="bf2e435b", attack_maps_sum=220528000640, attack_maps_popcount=1806,
        puct=[dict(deck=[1, 2, 0, 3, 11], c="sqrt2", sims=400, visits=[36, 36, 36, 36, 36, 36, 45, 46, 46, 46], nodes=4510),
              dict(deck=[1, 2, 0, 3, 11], c="sqrt2", sims=800, visits=[71, 80, 71, 71, 71, 71, 88, 89, 98, 89], nodes=9096),
              dict(deck=[1, 2, 0, 3, 11], c="2.0", sims=800, visits=[71, 78, 71, 71, 71, 71, 90, 90, 96, 90]),
              dict(deck=[1, 2, 0, 3, 11], c="5.0", sims=800, visits=[72, 74, 72, 72, 72, 72, 90, 90, 95, 90]),
              dict(deck=[4, 3, 1, 0, 2], c="sqrt2", sims=400, visits=[40, 40, 40, 40, 40, 49, 50, 50, 50], nodes=4743),
              dict(deck=[4, 3, 1, 0, 2], c="sqrt2", sims=800, visits=[77, 87, 77, 87, 77, 96, 96, 106, 96]),
              dict(deck=[4, 3, 1, 0, 2], c="2.0", sims=400, visits=[39, 40, 40, 40, 40, 50, 50, 50, 50])])

    # all canonical deals: sum perft(1), perft(2)
    if not quick:
        s1 = s2 = cnt = 0
        for a in range(16):
            for b in range(a + 1, 16):
                for c in range(16):
                    if c in (a, b):
                        continue
                    for d in range(c + 1, 16):
                        if d in (a, b):
                            continue
                        for e in range(16):
                            if e in (a, b, c, d):
                                continue
                            deck = [a, b, c, d, e]
                            n, _ = perft(Pos(deck), first_mover(deck), 2)
                            s1 += n[0]
                            s2 += n[1]
                            cnt += 1
            print("deals", a, cnt, flush=True)
        g["all_deals"] = dict(n_deals=cnt, sum_perft1=s1, sum_perft2=s2)

    # PUCT searches, uniform evaluator
    g["puct"] = []
    cases = [([1, 2, 0, 3, 11], math.sqrt(2.0), 400), ([4, 3, 1, 0, 2], math.sqrt(2.0), 400), ([1, 2, 0, 3, 11], 2.0, 400),
             ([0, 1, 2, 3, 4], 5.0, 200)]
    if not quick:
        cases += [([1, 2, 0, 3, 11], math.sqrt(2.0), 800), ([4, 3, 1, 0, 2], 2.0, 400)]
    for deck, c, sims in cases:
        arena, best, md = puct_search(Pos(deck), first_mover(deck), c, sims, uniform_policy)
        root = arena[0]
        g["puct"].append(dict(deck=deck, c_puct=c, sims=sims, visits=[arena[i].visits for i in root.children],
                              best=action_code(*arena[best].mov), n_nodes=len(arena), root_q=root.winrate, mean_depth=md,
                              child_q=[arena[i].winrate for i in root.children],
                              child_prior=[arena[i].prob for i in root.children]))
        print("puct", deck, c, sims, g["puct"][-1]["visits"], len(arena), flush=True)

    # encoder vectors on positions reached by random play
    g["encode"] = []
    for game in range(4):
        d = deal(7, game, 0)
        pos = Pos(d)
        col = first_mover(d)
        for ply in range(3 + game):
            mv = pos.moves(col)
            slot, frm, to, kind = mv[rand_index(rand_u32(7, game, ply, 0), len(mv))]
            pos.make_move(col, slot, frm, to, kind)
            col ^= 1
        pawns, kings = pos.boards()
        planes = encode(pos, col)
        g["encode"].append(dict(pawns=pawns, kings=kings, cards=pos.deck, side=col,
                                ones=[i for i, v in enumerate(planes) if v == 1.0]))

    # RNG known answers + deals + playouts under the shared counter RNG
    g["rng"] = [dict(seed=s, game=ga, step=st, draw=dr, value=rand_u32(s, ga, st, dr))
                for (s, ga, st, dr) in [(0, 0, 0, 0), (1, 2, 3, 4), (0xDEADBEEF, 12345, 77, 1), (2 ** 63 + 5, 2 ** 40 + 3, 4000000000, 12)]]
    g["deals"] = [dict(seed=s, game=ga, epoch=e, deck=deal(s, ga, e)) for (s, ga, e) in [(0, 0, 0), (42, 7, 0), (42, 7, 9), (1234567, 99999, 3)]]
    g["playouts"] = [playout(2024, game) for game in range(16 if quick else 64)]
    g["playouts_fixed_deck"] = [playout(5, game, deck=[1, 2, 0, 3, 11]) for game in range(8)]

    with open(OUT, "w") as f:
        json.dump(g, f, indent=None, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
