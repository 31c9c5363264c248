"""CPU-side tests of the product's host logic (no GPU): the C-ABI library loads and exports every symbol
include/tss.h declares, and its host-side mirror of the reference interface agrees with the oracle."""
import ctypes as C
import os
import re
from collections import Counter

import numpy as np
import pytest

import oracle.oracle as O
import timberborn_support_solver_b200 as T
from conftest import ROOT, golden, rows_to_grid, synth_terrain

ONE = T.PlatformDef(1, 1)


def o_defs(defs):
    return [d.dims() for d in defs]


def canon(clauses):
    return Counter(tuple(sorted(c)) for c in clauses)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "tss.h")).read()
    declared = set(re.findall(r"\b(tss_[a-z0-9_]+)\s*\(", header))
    assert declared == set(T.SIGNATURES), declared ^ set(T.SIGNATURES)
    lib = C.CDLL(T.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tss_version() == 103


def test_engine_needs_a_gpu_no_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(T.TssError) as e:
        T.Engine()
    assert e.value.code == -3  # TSS_E_CUDA


def test_sls_spec_constants_agree_with_model():
    """The CPU model re-declares the SLS spec's hash / tie-break functions (oracle never includes product headers):
    both sides evaluate them at the same probe points."""
    model = (C.c_uint32 * 9)()
    O.lib().tsso_sls_constants(model)
    product = (C.c_uint32 * 9)()
    T.load().tss_sls_spec_probe(product)
    assert list(model) == list(product)
    assert model[0] == 0x9E3779B9 and model[2] == 26 and model[8] == 1 << 20


# ---- world -----------------------------------------------------------------------------------------
def test_world_parse_matches_oracle(fixtures):
    for name in ("ex1", "ex2", "ex3"):
        text = O.world_to_toml(fixtures[name])
        g = T.WorldGrid.from_toml(text)
        assert np.array_equal(g.data, fixtures[name]) and not g.ragged
        assert g.to_toml() == text
    g = T.WorldGrid.from_toml('[world]\ngrid = ["XX", "X", ""]\n')
    assert g.ragged and g.data.tolist() == [[1, 1], [1, 0], [0, 0]]


@pytest.mark.parametrize("text,msg", [
    ('[world]\ngrid = ["X.X"]\n', "expected `X` or ` `"),   # world.rs:58
    ("[world]\ngrid = []\n", "invalid length 0"),             # world.rs:63-65
    ("[world]\n", "missing field"),
    ('[world]\ngrid = ["XX"\n', "unterminated"),
])
def test_world_errors(text, msg):
    with pytest.raises(T.TssError, match=re.escape(msg)) as e:
        T.WorldGrid.from_toml(text)
    assert e.value.code == -5
    with pytest.raises(O.WorldParseError):
        O.parse_world(text)


def test_synthetic_terrain_generator():
    """the host generator and the numpy one in conftest produce identical grids."""
    for (w, h, seed, t) in [(32, 32, 1, 0), (32, 32, 1, 99999), (256, 256, 1, 0), (16, 16, 7, 3)]:
        assert np.array_equal(T.WorldGrid.synthetic(w, h, seed, t).data, synth_terrain(w, h, seed, t))
    d = T.WorldGrid.synthetic(256, 256, 1, 0).data.mean()
    assert 0.69 < d < 0.71


# ---- platform ----------------------------------------------------------------------------------------
def test_platform_overlap_tables():
    g = golden("platform_overlap")
    mk = lambda r: T.Platform(r[0], r[1], T.PlatformDef(r[2], r[3]), bool(r[4]))
    for a, b in g["overlap_yes"]:
        assert mk(a).overlaps(mk(b)) and mk(b).overlaps(mk(a))
    for a, b in g["overlap_no"]:
        assert not mk(a).overlaps(mk(b)) and not mk(b).overlaps(mk(a))
    assert T.Platform(0, 0, T.PlatformDef(1, 4), True).dims() == (4, 1)


# ---- encoder -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["ex1", "ex2", "ex3"])
@pytest.mark.parametrize("defs", [T.PLATFORMS_DEFAULT[:1], T.PLATFORMS_DEFAULT], ids=["1x1", "default8"])
def test_encoder_matches_oracle(fixtures, name, defs):
    g = fixtures[name]
    enc = T.Encoding.encode(defs, T.WorldGrid(g))
    ref = O.Encoding(o_defs(defs), g)
    ocnf = ref.cnf()
    assert (enc.n_vars, enc.n_clauses, enc.n_lits) == (ocnf.n_vars, ocnf.n_clauses, ocnf.n_lits)
    assert enc.vars().dims == ref.dims
    assert np.array_equal(enc.vars().plat_var, ref.plat_var) and np.array_equal(enc.vars().terr_var, ref.terr_var)
    assert canon(enc.cnf().clauses()) == canon(ocnf.clauses())


def test_encoder_random_terrains_and_platform_sets():
    rng = np.random.default_rng(5)
    sets = [[(1, 1)], [(1, 1), (2, 2)], [(1, 1), (1, 3), (2, 2)], [(1, 1), (2, 3), (3, 3)], [(1, 1), (1, 2), (1, 3), (3, 3), (5, 5)]]
    for i in range(10):
        w, h = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        g = (rng.random((h, w)) < 0.7).astype(np.uint8)
        defs = sets[i % len(sets)]
        enc = T.Encoding.encode([T.PlatformDef(*d) for d in defs], T.WorldGrid(g))
        ref = O.Encoding(defs, g).cnf()
        assert canon(enc.cnf().clauses()) == canon(ref.clauses()), (w, h, defs)


def test_encoder_rejects_missing_1x1():
    with pytest.raises(T.TssError):
        T.Encoding.encode([T.PlatformDef(3, 3)], T.WorldGrid(np.ones((4, 4))))


def test_with_limits_matches_oracle(fixtures):
    g = fixtures["ex3"]
    enc = T.Encoding.encode(T.PLATFORMS_DEFAULT, T.WorldGrid(g))
    ref = O.Encoding(O.PLATFORMS_DEFAULT, g)
    weights = {(1, 1): 5, (1, 2): 1, (1, 3): 1, (1, 4): 1, (1, 5): 1, (1, 6): 1, (3, 3): 2, (5, 5): 4}   # crates/gui/src/app.rs:53-62
    for card, wl in [({(1, 1): 4}, None), ({(1, 1): 0}, None), ({(1, 1): 10, (1, 4): 2, (5, 5): 1}, None), ({}, 20), ({(1, 1): 3}, 12)]:
        lim = T.PlatformLimits({T.PlatformDef(*k): v for k, v in card.items()}, {T.PlatformDef(*k): v for k, v in weights.items()} if wl else {}, wl)
        mine = enc.with_limits(lim)
        theirs = ref.with_limits(card, weights if wl else None, wl)
        assert mine.n_vars == theirs.n_vars and canon(mine.clauses()) == canon(theirs.clauses()), (card, wl)


# ---- layout decode -------------------------------------------------------------------------------------
def test_layout_from_assignment_matches_oracle(fixtures):
    g = fixtures["ex2"]
    enc = T.Encoding.encode(T.PLATFORMS_DEFAULT, T.WorldGrid(g))
    ref = O.Encoding(O.PLATFORMS_DEFAULT, g)
    r, a, _ = ref.with_limits({(1, 1): 6}).solve()
    assert r == 10
    mine = T.PlatformLayout.from_assignment(a[: enc.n_vars + 1], enc)
    theirs = ref.layout_from_assignment(a)
    got = sorted((p.x, p.y, p.definition.width, p.definition.height, int(p.rotated)) for p in mine.platforms().values())
    assert got == sorted(theirs) and mine.platform_count() <= 6
    rng = np.random.default_rng(1)  # arbitrary (even unsound) assignments decode identically: largest def per anchor wins
    for _ in range(5):
        a = (rng.random(enc.n_vars + 1) < 0.02).astype(np.uint8)
        a[rng.integers(1, enc.n_vars, 20)] = 2
        got = sorted((p.x, p.y, p.definition.width, p.definition.height, int(p.rotated)) for p in T.PlatformLayout.from_assignment(a, enc).platforms().values())
        assert got == sorted(ref.layout_from_assignment(a))


def test_trivial_optimization_and_total_weight(fixtures):
    w = T.World(T.WorldGrid(fixtures["ex1"]))
    lay = T.PlatformLayout([T.Platform(0, 0, T.PlatformDef(5, 5)), T.Platform(3, 5, ONE), T.Platform(2, 3, ONE)])
    lay.run_trivial_optimization(w)
    assert list(lay.platforms().values()) == [T.Platform(0, 0, T.PlatformDef(5, 5))]
    weights = {T.PlatformDef(*k): v for k, v in {(1, 1): 5, (1, 2): 1, (1, 3): 1, (1, 4): 1, (1, 5): 1, (1, 6): 1, (3, 3): 2, (5, 5): 4}.items()}
    for plats in ([(0, 0, 5, 5, 0)], [(0, 0, 1, 6, 1)], [(0, 0, 3, 3, 0), (4, 4, 1, 1, 0)], [(1, 1, 1, 4, 0), (0, 0, 1, 2, 1)]):
        lay = T.PlatformLayout(T.Platform(x, y, T.PlatformDef(dw, dh), bool(r)) for x, y, dw, dh, r in plats)
        assert lay.total_weight(weights) == O.total_weight(plats, {k.dims(): v for k, v in weights.items()})
    assert T.PlatformLayout([T.Platform(0, 0, ONE), T.Platform(1, 0, ONE)]).platform_stats() == {ONE: 2}


def build_c_abi_smoke(tmp_path):
    """gcc (plain C, -std=c99 -pedantic) compiles tests/c_abi_smoke.c against include/tss.h and links libtss.so."""
    import subprocess
    exe = os.path.join(str(tmp_path), "c_abi_smoke")
    pkg = os.path.dirname(T.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L", pkg, "-l:libtss.so", f"-Wl,-rpath,{pkg}"])
    return exe


def test_header_is_plain_c_and_links(tmp_path):
    exe = build_c_abi_smoke(tmp_path)
    assert os.path.exists(exe)


def test_render_world_matches_the_repl_printout_rules():
    """crates/repl/src/main.rs:390-480: glyph + two spaces per tile; 1x1 = '☐', larger platforms as an outline (interior
    blank, straight runs as lines), ceiling '▒', unsupported ceiling yellow, overlapping platforms red, tiles outside the
    grid dropped.  Expected strings derived by hand from the reference's match arms and its NSWE glyph table."""
    grid = np.ones((5, 6), np.uint8)
    grid[4, :] = 0
    world = T.World(T.WorldGrid(grid))
    lay = T.PlatformLayout([T.Platform(0, 0, T.PlatformDef(3, 3)), T.Platform(4, 0, T.PlatformDef(1, 3)), T.Platform(3, 3, T.PlatformDef(1, 3), rotated=True),
                            T.Platform(5, 1, T.PlatformDef(1, 1))])
    text = T.render_world(world, lay)
    want = ["┌  ─  ┐  ▒  ╷  ▒  ",
            "│     │  ▒  │  ☐  ",
            "└  ─  ┘  ▒  ╵  ▒  ",
            "▒  ▒  ▒  ╶  ─  ╴  ",
            "                  "]
    assert text == "".join(r + "\n" for r in want)
    # no layout: the terrain alone (main.rs:244)
    assert T.render_world(world).splitlines()[0] == "▒  " * 6 and T.render_world(world).splitlines()[4] == "   " * 6
    # colours and clipping: an out-of-bounds 1x2 platform keeps its in-grid tile, drawn red because it overlaps another one
    a, b = T.Platform(5, 3, T.PlatformDef(1, 2), rotated=True), T.Platform(5, 3, T.PlatformDef(1, 1))
    val = T.ValidationResult({(0, 0)}, {a}, {a})
    text = T.render_world(world, T.PlatformLayout([a]), val, color=True)
    rows = text.splitlines()
    assert rows[0].startswith("\x1b[33m▒\x1b[39m  ▒  ") and rows[3].endswith("\x1b[31m╶\x1b[39m  ")
    assert b.overlaps(a)


def test_merge_supports_keeps_layouts_complete_and_disjoint():
    """tss_layout_merge_supports (host side of what the engine does on grids larger than 32x32): a complete layout of 1x1
    supports stays complete under the reference's validate(), footprints stay disjoint and in bounds, only platforms of the
    set appear, and the platform count drops."""
    rng = np.random.default_rng(3)
    for (w, h, dens) in ((48, 40, 0.8), (70, 33, 0.6), (20, 20, 1.0)):
        grid = (rng.random((h, w)) < dens).astype(np.uint8)
        # a complete start: supports on a 4 x 4 lattice plus one under every tile that lattice leaves unsupported
        sites = np.zeros_like(grid)
        sites[1::4, 1::4] = grid[1::4, 1::4]
        unc = O.validate(grid, [(int(x), int(y), 1, 1, 0) for y, x in zip(*np.nonzero(sites))]).unsupported
        sites |= (unc != 0).astype(np.uint8)
        plats = [T.Platform(int(x), int(y), T.PlatformDef(1, 1)) for y, x in zip(*np.nonzero(sites))]
        assert O.validate(grid, [(p.x, p.y, 1, 1, 0) for p in plats]).is_valid
        lay = T.PlatformLayout(plats)
        lay.merge_supports(T.World(T.WorldGrid(grid)), T.PLATFORMS_DEFAULT)
        out = list(lay.platforms().values())
        v = O.validate(grid, [(p.x, p.y, p.definition.width, p.definition.height, int(p.rotated)) for p in out])
        assert v.is_valid and len(out) < len(plats)
        assert all(p.definition in T.PLATFORMS_DEFAULT for p in out) and any(p.definition != T.PlatformDef(1, 1) for p in out)
    # with the 1x1-only set nothing changes; layouts that already hold larger platforms are refused
    lay = T.PlatformLayout(plats)
    lay.merge_supports(T.World(T.WorldGrid(grid)), T.PLATFORMS_DEFAULT[:1])
    assert sorted(lay.platforms()) == sorted((p.x, p.y) for p in plats)
    with pytest.raises(T.TssError):
        T.PlatformLayout([T.Platform(0, 0, T.PlatformDef(3, 3))]).merge_supports(T.World(T.WorldGrid(grid)), T.PLATFORMS_DEFAULT)


# ---- instance bridge (tss_instance_find): from the clauses a solver is handed back to terrain / platform set / limits ----
def test_instance_registry_finds_the_instance_from_its_clauses(fixtures):
    """The drivers hand their solver a bare Cnf (crates/repl/src/solver_runner.rs:8-20).  with_limits records every instance it
    lowers; the solver side finds it again from the clauses alone — exactly (same CNF) or by its base clauses — and an
    unrelated CNF is not found."""
    from timberborn_support_solver_b200 import _lib
    lib = T.load()
    g = T.WorldGrid(fixtures["ex1"])
    enc = T.Encoding.encode(T.PLATFORMS_DEFAULT, g)
    other = T.Encoding.encode(T.PLATFORMS_DEFAULT[:1], T.WorldGrid(fixtures["ex3"]))
    weights = {T.PlatformDef(1, 1): 5, T.PlatformDef(3, 3): 2}
    cnf_w = enc.with_limits(T.PlatformLimits({}, weights, 11))
    cnf_o = other.with_limits(T.PlatformLimits.new_unweighted({ONE: 4}))
    cnf = enc.with_limits(T.PlatformLimits.new_unweighted({ONE: 2, T.PlatformDef(5, 5): 1}))

    def find(c):
        lits, offs = np.ascontiguousarray(c.lits, np.int32), np.ascontiguousarray(c.offsets, np.uint32)
        h, info, wts = C.c_void_p(), _lib.InstanceInfo(), np.zeros((16, 3), np.int32)
        rc = lib.tss_instance_find(lits.ctypes.data_as(C.POINTER(C.c_int32)), offs.ctypes.data_as(C.POINTER(C.c_uint32)), c.n_clauses, c.n_vars,
                                   C.byref(h), C.byref(info), wts.ctypes.data_as(C.POINTER(C.c_int32)), 16)
        terrain = None
        if rc == 10:
            out = np.zeros(info.w * info.h, np.uint8)
            w, hh = C.c_int32(), C.c_int32()
            assert lib.tss_encoding_terrain(h, out.ctypes.data_as(C.POINTER(C.c_uint8)), out.size, C.byref(w), C.byref(hh)) == 0
            terrain = out.reshape(hh.value, w.value)
            defs, n = (_lib.Dims * 16)(), C.c_int32()
            assert lib.tss_encoding_defs(h, defs, 16, C.byref(n)) == 0 and n.value == info.n_defs
            lib.tss_encoding_destroy(h)
        return rc, info, wts, terrain

    rc, info, wts, terrain = find(cnf)
    assert rc == 10 and info.exact == 1 and (info.w, info.h, info.n_defs) == (g.width, g.height, 8)
    assert info.card_limit_1x1 == 2 and info.n_other_card_limits == 1 and not info.has_weight_limit
    assert np.array_equal(terrain, g.data)
    rc, info, wts, terrain = find(cnf_w)                       # an older record is still there, with its own limits
    assert rc == 10 and info.exact == 1 and info.card_limit_1x1 == -1 and info.has_weight_limit and info.weight_limit == 11
    assert sorted(map(tuple, wts[: info.n_weights].tolist())) == [(1, 1, 5), (3, 3, 2)]
    rc, info, _, terrain = find(cnf_o)
    assert rc == 10 and info.n_defs == 1 and info.card_limit_1x1 == 4 and np.array_equal(terrain, fixtures["ex3"])
    # the same base clauses with the limits lowered by "someone else" (here: one more clause): found by its base, flagged inexact
    ext = T.Cnf(cnf.n_vars + 1, np.append(cnf.lits, np.int32(cnf.n_vars + 1)), np.append(cnf.offsets, np.uint32(len(cnf.lits) + 1)))
    rc, info, _, _ = find(ext)
    assert rc == 10 and info.exact == 0 and info.card_limit_1x1 == 2
    # unrelated clauses
    junk = T.Cnf(3, np.array([1, -2, 3], np.int32), np.array([0, 2, 3], np.uint32))
    assert find(junk)[0] == 0
    # the encoding handle may be destroyed by its owner: the registry keeps the instance alive
    del enc
    import gc
    gc.collect()
    assert find(cnf)[0] == 10


# ---- an independent third reader for the project format (oracle and product parse TOML with sibling code) --------
def test_world_toml_agrees_with_an_independent_toml_parser():
    """src/world.rs:21-40,49-79: `[world] grid = [rows of 'X' / ' ']`.  Python's tomllib knows nothing of either implementation:
    what the product serialises must read back through it to the same grid, and what it reads the product and the oracle
    must read identically (rows of unequal length are left-aligned and padded, world.rs:82-86)."""
    import tomllib
    rng = np.random.default_rng(5)
    for h, w in [(1, 1), (3, 7), (16, 16), (21, 16), (32, 32), (5, 40)]:
        g = (rng.random((h, w)) < 0.6).astype(np.uint8)
        g[0, w - 1] = 1                                          # keep the width recoverable: the longest row defines it
        text = T.WorldGrid(g).to_toml()
        rows = tomllib.loads(text)["world"]["grid"]
        assert len(rows) == h and all(set(r) <= {"X", " "} for r in rows)
        back = np.array([[1 if (i < len(r) and r[i] == "X") else 0 for i in range(w)] for r in rows], np.uint8)
        assert np.array_equal(back, g)
        assert np.array_equal(T.WorldGrid.from_toml(text).data, g)
        assert np.array_equal(O.parse_world(text)[0], g)
        assert O.world_to_toml(g) == text
    # hand-written document with comments, ragged rows and another table: all three readers agree
    doc = ('# a project\n[meta]\nname = "x # not a comment"\ntags = [\n  "a", ["b"],\n]\n\n[world]\nseed = 7  # serde ignores undeclared fields\n'
           'grid = [\n  "XX X",  # first row\n  "X",\n  "  XXXXX",\n]\nextra = { k = "v" }\n')
    rows = tomllib.loads(doc)["world"]["grid"]
    want = np.array([[1 if (i < len(r) and r[i] == "X") else 0 for i in range(7)] for r in rows], np.uint8)
    assert np.array_equal(T.WorldGrid.from_toml(doc).data, want) and np.array_equal(O.parse_world(doc)[0], want)
