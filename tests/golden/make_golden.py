"""Regenerates tests/golden/*.json from the reference tree (run in the build container only:
/root/reference does not exist on the GPU box).  These are the ONLY golden vectors the reference holds
for the hot path:

  readme_layouts.json   README.md:44-119 — four valid 1x1 layouts (18/17/16/15 supports) on the 21x16 README terrain
  platform_overlap.json src/platform.rs:139-233 — the test_case/test_matrix tables of Platform::overlaps
  fixtures.json         test/ex1.toml, ex2.toml, ex3.toml — grids as row strings (inputs only)
  dag_default8.json     src/encoder.rs:45-51 — the doc-comment diagram of the transitively reduced platform DAG
"""
import itertools
import json
import os
import re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def readme_layouts():
    lines = open(os.path.join(REF, "README.md"), encoding="utf-8").read().splitlines()
    layouts, cur, marked = [], None, None
    for ln in lines:
        m = re.match(r"Solution: \((\d+) marked\)", ln)
        if m:
            cur, marked = [], int(m.group(1))
            continue
        if cur is not None:
            cells = [c for c in ln if c in "▒░█"]
            if cells:
                cur.append(cells)
            else:
                layouts.append((marked, cur))
                cur = None
    out = []
    for marked, rows in layouts:
        w = max(len(r) for r in rows)
        terrain = ["".join("X" if c in "▒█" else " " for c in r).ljust(w) for r in rows]
        supports = [[x, y] for y, r in enumerate(rows) for x, c in enumerate(r) if c == "█"]
        assert len(supports) == marked, (len(supports), marked)
        out.append({"marked": marked, "terrain": terrain, "supports": supports})
    assert [o["marked"] for o in out] == [18, 17, 16, 15]
    assert all(o["terrain"] == out[0]["terrain"] for o in out)
    return {"source": "README.md:44-119", "terrain": out[0]["terrain"],
            "layouts": [{"marked": o["marked"], "supports": o["supports"]} for o in out],
            "unresolved_bound": 14}


def platform_overlap():
    src = open(os.path.join(REF, "src/platform.rs"), encoding="utf-8").read()
    tests = src[src.index("mod tests"):]
    yes_src, no_src = tests.split("fn platform_overlap_yes")[0], tests.split("fn platform_overlap_yes")[1].split("fn platform_overlap_no")[0]

    def plat(s):
        m = re.match(r"platform!\((\d)x(\d) @ (\d+), (\d+)\)", s.strip())
        w, h, x, y = map(int, m.groups())
        return [x, y, w, h, 0]

    def cases(block):
        out = []
        for m in re.finditer(r"#\[test_case\((platform![^)]*\)), (platform![^)]*\))\)\]", block):
            out.append([plat(m.group(1)), plat(m.group(2))])
        for m in re.finditer(r"#\[test_matrix\(\s*\[(.*?)\],\s*\[(.*?)\]\s*\)\]", block, re.S):
            a = re.findall(r"platform!\([^)]*\)", m.group(1))
            b = re.findall(r"platform!\([^)]*\)", m.group(2))
            out += [[plat(x), plat(y)] for x, y in itertools.product(a, b)]
        return out

    yes, no = cases(yes_src), cases(no_src)
    assert len(yes) == 22 and len(no) == 18, (len(yes), len(no))
    return {"source": "src/platform.rs:139-233", "record": "[x, y, def_w, def_h, rotated]", "overlap_yes": yes, "overlap_no": no}


def fixtures():
    import tomllib
    out = {}
    for name in ("ex1", "ex2", "ex3"):
        with open(os.path.join(REF, "test", name + ".toml"), "rb") as f:
            out[name] = {"source": f"test/{name}.toml", "grid": tomllib.load(f)["world"]["grid"]}
    return out


def dag_default8():
    # transcribed from the diagram at src/encoder.rs:45-51 (arrows point larger -> smaller = implication)
    chain_a = ["1x1", "1x2", "1x3", "1x4", "1x5", "1x6"]
    chain_b = ["1x1", "2x1", "3x1", "4x1", "5x1", "6x1"]
    edges = [[a, b] for a, b in zip(chain_a, chain_a[1:])] + [[a, b] for a, b in zip(chain_b, chain_b[1:])]
    edges += [["1x3", "3x3"], ["3x1", "3x3"], ["1x5", "5x5"], ["5x1", "5x5"], ["3x3", "5x5"]]
    return {"source": "src/encoder.rs:45-51", "edges_smaller_to_larger": edges}


if __name__ == "__main__":
    for name, fn in [("readme_layouts", readme_layouts), ("platform_overlap", platform_overlap), ("fixtures", fixtures),
                     ("dag_default8", dag_default8)]:
        with open(os.path.join(OUT, name + ".json"), "w", encoding="utf-8") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
