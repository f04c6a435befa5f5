"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE: the CPU restatement of the reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

STATE_DTYPE = np.dtype([("pawns", "<u4", (2,)), ("kings", "<u4", (2,)), ("cards", "u1", (5,)), ("side", "u1"),
                        ("result", "u1"), ("flags", "u1")])
assert STATE_DTYPE.itemsize == 24

RESULT_NAMES = {0: "Capture", 1: "RedWin", 2: "BlueWin", 3: "InProgress"}


def build():
    src = os.path.join(ORACLE_DIR, "onb_oracle.cpp")
    if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


class TreeDump(C.Structure):
    _fields_ = [("visits", C.c_void_p), ("reward", C.c_void_p), ("winrate", C.c_void_p), ("prior", C.c_void_p),
                ("action", C.c_void_p), ("parent", C.c_void_p), ("first_child", C.c_void_p), ("n_child", C.c_void_p),
                ("flags", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_rand_u32.restype = C.c_uint32
        L.orc_rand_u32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_deal.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_new_games.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_gen_moves.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_gen_moves_card.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_make_move.argtypes = [C.c_void_p, C.c_uint16]
        L.orc_current_state.argtypes = [C.c_void_p]
        L.orc_legal_masks.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_encode.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_env_step.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_env_step_random.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p]
        L.orc_playout_games.restype = C.c_int64
        L.orc_playout_games.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
        L.orc_perft.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_mcts_search.restype = C.c_int64
        L.orc_mcts_search.argtypes = [C.c_void_p, C.c_double, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64]
        L.orc_hash_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_mcts_set_noise.argtypes = [C.c_int, C.c_double, C.c_double, C.c_uint64, C.c_uint64]
        L.orc_uct_search_batch.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int] + [C.c_void_p] * 7
        L.orc_net_forward.restype = C.c_int
        L.orc_net_forward.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_mcts_search_batch.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_uint32, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bench_env.restype = C.c_double
        L.orc_bench_env.argtypes = [C.c_int64, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        L.orc_bench_env_steps.restype = C.c_double
        L.orc_bench_env_steps.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        L.orc_bench_mcts.restype = C.c_double
        L.orc_bench_mcts.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_uint32, C.c_int, C.c_void_p]
        L.orc_bench_perft.restype = C.c_double
        L.orc_bench_perft.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------- helpers
def attack_maps():
    out = np.zeros(800, dtype=np.uint32)
    lib().orc_attack_maps(_p(out))
    return out


def new_games(n, seed=0, game0=0, epoch=0, deck=None):
    g = np.zeros(n, dtype=STATE_DTYPE)
    d = np.asarray(deck, dtype=np.uint8) if deck is not None else None
    lib().orc_new_games(_p(g), n, game0, seed, epoch, _p(d))
    return g


def make_state(deck, pawns=None, kings=None, side=None):
    g = new_games(1, deck=deck)
    if pawns is not None:
        g["pawns"][0] = pawns
    if kings is not None:
        g["kings"][0] = kings
    if side is not None:
        g["side"][0] = side
    return g


def gen_moves(g, side=None):
    out = np.zeros(40, dtype=np.uint16)
    s = int(g["side"][0]) if side is None else side
    n = lib().orc_gen_moves(_p(g), s, _p(out))
    return out[:n].copy()


def gen_moves_card(g, side, card_index):
    out = np.zeros(40, dtype=np.uint16)
    n = lib().orc_gen_moves_card(_p(g), side, card_index, _p(out))
    return out[:n].copy()


def make_move(g, action):
    return lib().orc_make_move(_p(g), int(action))


def legal_masks(g):
    out = np.zeros((len(g), 2), dtype=np.uint32)
    lib().orc_legal_masks(_p(g), len(g), _p(out))
    return out


def encode(g):
    out = np.zeros((len(g), 21, 5, 5), dtype=np.float32)
    lib().orc_encode(_p(g), len(g), _p(out))
    return out


def env_step(g, actions):
    a = np.ascontiguousarray(actions, dtype=np.uint16)
    lib().orc_env_step(_p(g), len(g), _p(a))


def env_step_random(g, seed, step, policy=0, auto_reset=False, game0=0, deck=None):
    acts = np.zeros(len(g), dtype=np.uint16)
    d = np.asarray(deck, dtype=np.uint8) if deck is not None else None
    lib().orc_env_step_random(_p(g), len(g), game0, seed, step, policy, int(auto_reset), _p(d), _p(acts))
    return acts


def playout_games(n, seed, game0=0, policy=0, max_plies=1 << 30, deck=None):
    g = np.zeros(n, dtype=STATE_DTYPE)
    plies = np.zeros(n, dtype=np.uint32)
    trace = np.zeros(n, dtype=np.uint64)
    d = np.asarray(deck, dtype=np.uint8) if deck is not None else None
    total = lib().orc_playout_games(_p(g), n, game0, seed, policy, max_plies, _p(d), _p(plies), _p(trace))
    return g, plies, trace, total


def perft(g, depth):
    nodes = np.zeros(depth, dtype=np.uint64)
    wins = np.zeros(depth, dtype=np.uint64)
    zero = np.zeros(depth, dtype=np.uint64)
    lib().orc_perft(_p(g), depth, _p(nodes), _p(wins), _p(zero))
    return nodes, wins, zero


EVAL_CB = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)


def mcts_search(g, c_puct, sims, evaluator=0, dump=False, callback=None):
    """callback(planes[525] float32 ndarray) -> (policy[50], value): a host evaluator (the network as a black box)."""
    best = np.zeros(1, dtype=np.uint16)
    pi = np.zeros(50, dtype=np.float32)
    rv = np.zeros(1, dtype=np.uint32)
    rq = np.zeros(1, dtype=np.float64)
    ps = np.zeros(1, dtype=np.int32)
    md = np.zeros(1, dtype=np.float64)
    mc = np.zeros(1, dtype=np.float64)
    cap = 1 + 40 * sims + 2
    res = {}
    td = None
    if dump:
        arrs = dict(visits=np.zeros(cap, np.uint32), reward=np.zeros(cap, np.float64), winrate=np.zeros(cap, np.float64),
                    prior=np.zeros(cap, np.float64), action=np.zeros(cap, np.uint16), parent=np.zeros(cap, np.int32),
                    first_child=np.zeros(cap, np.uint32), n_child=np.zeros(cap, np.uint32), flags=np.zeros(cap, np.uint8))
        td = TreeDump(*[_p(arrs[k]) for k, _ in TreeDump._fields_])
    cb = None
    if callback is not None:
        def _cb(planes_p, pol_p, val_p, _user):
            planes = np.ctypeslib.as_array(planes_p, shape=(525,)).copy()
            pol, val = callback(planes)
            np.ctypeslib.as_array(pol_p, shape=(50,))[:] = np.asarray(pol, dtype=np.float32).reshape(50)
            val_p[0] = float(val)
        cb = EVAL_CB(_cb)
        evaluator = 2
    n = lib().orc_mcts_search(_p(g), c_puct, sims, evaluator, C.cast(cb, C.c_void_p) if cb else None, None, _p(best), _p(pi), _p(rv), _p(rq), _p(ps), _p(md),
                              _p(mc), C.byref(td) if td is not None else None, cap)
    res.update(n_nodes=int(n), best=int(best[0]), pi=pi.reshape(2, 25), root_visits=int(rv[0]), root_q=float(rq[0]),
               pass_seen=int(ps[0]), mean_depth=float(md[0]), mean_children=float(mc[0]))
    if dump:
        res["tree"] = {k: v[:n].copy() for k, v in arrs.items()}
    return res


def mcts_search_batch(roots, c_puct, sims, evaluator=0, threads=1):
    n = len(roots)
    best = np.zeros(n, dtype=np.uint16)
    cv = np.zeros((n, 40), dtype=np.uint32)
    nn = np.zeros(n, dtype=np.int64)
    rq = np.zeros(n, dtype=np.float64)
    pi = np.zeros((n, 2, 25), dtype=np.float32)
    ps = np.zeros(n, dtype=np.int32)
    lib().orc_mcts_search_batch(_p(roots), n, c_puct, sims, evaluator, threads, _p(best), _p(cv), _p(nn), _p(rq), _p(pi), _p(ps))
    return dict(best=best, child_visits=cv, n_nodes=nn, root_q=rq, pi=pi, pass_seen=ps)


def mcts_set_noise(enabled, epsilon=0.25, alpha=0.03, seed=0, game0=0):
    """train-mode root noise for the searches that follow (tree i of a batch uses global tree id game0 + i)"""
    lib().orc_mcts_set_noise(int(bool(enabled)), epsilon, alpha, seed, game0)


def uct_search_batch(roots, exploration_c, min_node_visits, sims, seed=0, game0=0, threads=1):
    """Plain UCT with random rollouts (onitama-game/src/ai/mcts/mcts_arena.rs), one tree per root."""
    n = len(roots)
    best = np.zeros(n, dtype=np.uint16)
    cv = np.zeros((n, 40), dtype=np.uint32)
    cr = np.zeros((n, 40), dtype=np.int32)
    nn = np.zeros(n, dtype=np.int64)
    wr = np.zeros(n, dtype=np.float32)
    ps = np.zeros(n, dtype=np.int32)
    plies = np.zeros(1, dtype=np.uint64)
    lib().orc_uct_search_batch(_p(roots), n, exploration_c, min_node_visits, sims, seed, game0, threads, _p(best), _p(cv), _p(cr), _p(nn), _p(wr),
                               _p(ps), _p(plies))
    return dict(best=best, child_visits=cv, child_rewards=cr, n_nodes=nn, best_winrate=wr, pass_seen=ps, rollout_plies=int(plies[0]))


def net_forward(params, planes):
    """CPU restatement of ConvResNet::forward (net.rs:215-232): params = {VarStore name: f32 array}; planes [n,21,5,5]."""
    names, arrays = [], []
    for k, v in params.items():
        if k.endswith("num_batches_tracked"):
            continue
        if hasattr(v, "detach"):
            v = v.detach().float().cpu().numpy()
        names.append(k.encode())
        arrays.append(np.ascontiguousarray(v, dtype=np.float32))
    m = len(names)
    planes = np.ascontiguousarray(planes, dtype=np.float32)
    n = planes.shape[0]
    pol = np.zeros((n, 50), dtype=np.float32)
    val = np.zeros(n, dtype=np.float32)
    c_names = (C.c_char_p * m)(*names)
    c_data = (C.c_void_p * m)(*[a.ctypes.data for a in arrays])
    c_numel = (C.c_int64 * m)(*[a.size for a in arrays])
    rc = lib().orc_net_forward(m, c_names, c_data, c_numel, _p(planes), n, _p(pol), _p(val))
    if rc != 0:
        raise ValueError("orc_net_forward: missing or mis-sized tensor")
    return pol, val


def hash_eval(planes):
    planes = np.ascontiguousarray(planes, dtype=np.float32).reshape(-1, 525)
    pol = np.zeros((len(planes), 50), dtype=np.float32)
    val = np.zeros(len(planes), dtype=np.float32)
    for i in range(len(planes)):
        lib().orc_hash_eval(_p(planes[i]), _p(pol[i]), C.c_void_p(val[i:].ctypes.data))
    return pol, val


def decode_action(a):
    a = int(a)
    return dict(to=a & 31, frm=(a >> 5) & 31, card_idx=(a >> 10) & 3, piece=(a >> 12) & 1, is_pass=(a >> 13) & 1)


def sq_name(n):
    return "abcde"[n % 5] + str(5 - n // 5)
