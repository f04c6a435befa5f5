"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/onb.h declares, its host-side helpers
agree with the oracle and the golden vectors, and it FAILS LOUDLY without a CUDA device (no CPU fallback)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from onitama_alphazero_b200 import _lib
    return _lib.load()


def test_header_symbols_all_exported(lib):
    from onitama_alphazero_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "onb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(onb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "libonb.so does not export %s" % name
    # the Python binding table covers exactly the header
    assert declared == set(_lib.SYMBOLS)
    assert lib.onb_version() == 200


def test_attack_table_matches_reference_hex(lib):
    out = np.zeros(800, dtype=np.uint32)
    assert lib.onb_attack_maps(out.ctypes.data) == 0
    assert out.tolist() == G["attack_maps"]          # parsed from the reference's card.rs by gen_golden.py
    assert out.tolist() == O.attack_maps().tolist()  # oracle's shift/mask restatement (card.rs:520-604)


def test_rng_and_deal_match_oracle_and_golden(lib):
    for r in G["rng"]:
        assert lib.onb_rand_u32(r["seed"], r["game"], r["step"], r["draw"]) == r["value"]
    rs = np.random.RandomState(0)
    for _ in range(200):
        seed, game = int(rs.randint(0, 2 ** 62)), int(rs.randint(0, 2 ** 40))
        step, draw = int(rs.randint(0, 2 ** 31)), int(rs.randint(0, 16))
        assert lib.onb_rand_u32(seed, game, step, draw) == O.lib().orc_rand_u32(seed, game, step, draw)
        a = np.zeros(5, np.uint8)
        b = np.zeros(5, np.uint8)
        assert lib.onb_deal(seed, game, step, a.ctypes.data) == 0
        O.lib().orc_deal(seed, game, step, b.ctypes.data)
        assert a.tolist() == b.tolist() and len(set(a.tolist())) == 5
    for d in G["deals"]:
        a = np.zeros(5, np.uint8)
        lib.onb_deal(d["seed"], d["game"], d["epoch"], a.ctypes.data)
        assert a.tolist() == d["deck"]


def test_start_states_match_oracle(lib):
    from onitama_alphazero_b200 import start_states
    decks = np.array([[1, 2, 0, 3, 11], [4, 3, 1, 0, 2], [0, 1, 2, 3, 4], [15, 14, 13, 12, 10]], dtype=np.uint8)
    s = start_states(decks)
    for i, d in enumerate(decks):
        assert s[i:i + 1].tobytes() == O.new_games(1, deck=d).tobytes()
    bad = np.array([[1, 2, 0, 3, 16]], dtype=np.uint8)
    out = np.zeros(1, dtype=s.dtype)
    assert lib.onb_start_states(bad.ctypes.data, 1, out.ctypes.data) == -1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from onitama_alphazero_b200 import Context, OnbError
    with pytest.raises(OnbError) as e:
        Context(16)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    assert lib.onb_sync(None) == -1 and lib.onb_env_step(None, None, 0, 0, 0) == -1


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "onitama_alphazero_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in src and "liboracle" not in src and "oracle/" not in src.replace("oracle/onb_oracle.cpp", ""), f


def test_cpp_host_mirror_builds_and_fails_loudly_without_gpu(lib):
    """include/onitama_b200.hpp compiles against the C ABI; without a device its first GPU call throws (no CPU fallback)."""
    import subprocess
    import torch
    import __graft_entry__ as ge
    exe = ge.build_cpp_mirror_test()
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "no CPU fallback" in r.stdout


def test_header_is_c99_clean_and_links_from_c(lib):
    """include/onb.h compiles as strict C99 (-Wall -Wextra -Werror -pedantic); a pure C program links libonb.so, uses the
    host-side helpers and sees onb_create fail loudly without a device (or round-trips a reset with one)."""
    import subprocess
    src = os.path.join(ROOT, "tests", "c", "abi_host_only.c")
    exe = os.path.join(ROOT, "tests", "c", "abi_host_only")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L" + os.path.join(ROOT, "onitama_alphazero_b200"), "-lonb", "-Wl,-rpath,$ORIGIN/../../onitama_alphazero_b200"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C ABI OK" in r.stdout, r.stdout + r.stderr


def test_rust_bindings_cover_the_header():
    """include/onb_sys.rs (the unbuilt Rust side of the boundary) binds every symbol onb.h declares, with the
    same arity."""
    hdr = open(os.path.join(ROOT, "include", "onb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    rs = open(os.path.join(ROOT, "include", "onb_sys.rs")).read()
    decls = re.findall(r"ONB_API\s+[^;]+?\s+(onb_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)
    assert len(decls) >= 36
    for name, args in decls:
        m = re.search(r"pub fn %s\(([^)]*)\)" % name, rs)
        assert m, "onb_sys.rs does not bind %s" % name
        n_c = 0 if args.strip() in ("", "void") else len(args.split(","))
        n_rs = 0 if not m.group(1).strip() else len(m.group(1).split(","))
        assert n_c == n_rs, name
    assert "pub struct onb_state" in rs and "pub struct onb_config" in rs


def test_missing_extension_fails_loudly(monkeypatch):
    """Without the built CUDA library the package refuses to work (no silent fallback of any kind)."""
    from onitama_alphazero_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "onitama_alphazero_b200", "does_not_exist.so"))
    with pytest.raises(ImportError) as e:
        _lib.load()
    assert "no CPU fallback" in str(e.value)


def test_replay_minibatch_indices_are_distinct_and_uniform(lib):
    """onb_replay_indices (the keyed permutation behind onb_replay_sample = choose_multiple, train.rs:280-283): min(batch, size) DISTINCT
    slots below size, the same for the same (seed, size), every slot equally likely over seeds."""
    import numpy as np

    def indices(size, batch, seed):
        out = np.zeros(min(size, batch), dtype=np.int64)
        assert lib.onb_replay_indices(size, batch, seed, out.ctypes.data_as(C.c_void_p)) == 0
        return out

    for size, batch in ((1, 1), (2, 5), (7, 7), (1000, 64), (180_000, 512), (5, 0)):
        idx = indices(size, batch, 42)
        assert len(idx) == min(size, batch) and len(set(idx.tolist())) == len(idx)
        assert len(idx) == 0 or (0 <= idx.min() and idx.max() < size)
        assert np.array_equal(idx, indices(size, batch, 42))
    assert sorted(indices(10, 20, 1).tolist()) == list(range(10))          # batch >= size: every sample exactly once
    assert not np.array_equal(indices(1000, 64, 1), indices(1000, 64, 2))
    cnt = np.zeros(500)
    for seed in range(1500):
        cnt[indices(500, 50, seed)] += 1
    assert abs(cnt.mean() - 150) < 1e-9 and cnt.std() < 3 * np.sqrt(150 * 0.9) / np.sqrt(2) * 1.2 and cnt.min() > 100 and cnt.max() < 205
    assert lib.onb_replay_indices(-1, 3, 0, None) == -1
