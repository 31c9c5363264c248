// Host-side mirror of the reference library API for the hot path (see host_model.hpp).  Flat arrays and index
// tables instead of the reference's HashMap/petgraph machinery; results (clause multiset, decoded layouts) are
// checked against the oracle restatement in tests/.
#include "host_model.hpp"

#include <algorithm>
#include <cctype>
#include <cstring>
#include <map>
#include <sstream>

namespace tss {

int dims_partial_cmp(Dims a, Dims b) {  // src/math/dimensions.rs:74-114
    bool ae = a.w <= 0 || a.h <= 0, be = b.w <= 0 || b.h <= 0;
    if (ae || be) return ae && be ? 0 : (ae ? -1 : 1);
    int cw = (a.w > b.w) - (a.w < b.w), ch = (a.h > b.h) - (a.h < b.h);
    if (cw == 0 && ch == 0) return 0;
    if (cw * ch < 0) return 2;
    return (cw < 0 || ch < 0) ? -1 : 1;
}

BitGrid BitGrid::from_bytes(const uint8_t* g, int w, int h) {
    BitGrid b(w, h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if (g[(size_t)y * w + x]) b.set(x, y);
    return b;
}
int BitGrid::count() const {
    int n = 0;
    for (uint32_t r : rows) n += __builtin_popcount(r);
    return n;
}

// ------------------------------------------------------------------------------------------------ world
namespace {
struct Cursor {
    const char* s;
    size_t i = 0, n;
    explicit Cursor(const char* t) : s(t), n(std::strlen(t)) {}
    bool eof() const { return i >= n; }
    char peek() const { return s[i]; }
    void skip(bool newlines) {
        while (i < n) {
            char c = s[i];
            if (c == ' ' || c == '\t' || c == '\r' || (newlines && c == '\n')) i++;
            else if (c == '#') while (i < n && s[i] != '\n') i++;
            else break;
        }
    }
};
}  // namespace

std::string parse_world_toml(const char* text, std::vector<uint8_t>& grid, int& w, int& h, bool& ragged) {
    // The project file is `[world]` + `grid = [ "..", ... ]` (crates/repl/src/main.rs:272-278 via the toml crate);
    // this reader accepts that subset: table headers, one array-of-strings key, comments, trailing commas; other tables and
    // keys are skipped the way serde skips undeclared fields.
    Cursor c(text);
    std::vector<std::string> rows;
    bool in_world = false, have_grid = false;
    for (c.skip(true); !c.eof(); c.skip(true)) {
        if (c.peek() == '[') {
            const char* close = std::strchr(c.s + c.i, ']');
            if (!close) return "unterminated table header";
            std::string name(c.s + c.i + 1, close);
            name.erase(std::remove_if(name.begin(), name.end(), [](unsigned char ch) { return std::isspace(ch); }), name.end());
            in_world = name == "world";
            c.i = (size_t)(close - c.s) + 1;
            continue;
        }
        size_t k0 = c.i;
        while (!c.eof() && (std::isalnum((unsigned char)c.peek()) || c.peek() == '_' || c.peek() == '-')) c.i++;
        std::string key(c.s + k0, c.s + c.i);
        if (key.empty()) return "expected a key";
        c.skip(false);
        if (c.eof() || c.peek() != '=') return "expected `=` after key `" + key + "`";
        c.i++;
        c.skip(false);
        if (!in_world || key != "grid") {
            // serde ignores fields a struct does not declare (neither Project, src/lib.rs:14-17, nor World, src/world.rs:13-16, denies
            // unknown fields): skip the value — a string, a (nested) array or inline table, or a bare scalar up to the end of the line
            int depth = 0;
            bool done = false;
            while (!c.eof() && !done) {
                const char ch = c.peek();
                if (ch == '"' || ch == '\'') {
                    c.i++;
                    while (!c.eof() && c.peek() != ch) { if (ch == '"' && c.peek() == '\\') c.i++; c.i++; }
                    if (c.eof()) return "unterminated string";
                    c.i++;
                } else if (ch == '[' || ch == '{') { depth++; c.i++; }
                else if (ch == ']' || ch == '}') { if (--depth < 0) return "unbalanced bracket in the value of `" + key + "`"; c.i++; }
                else if (ch == '#') { while (!c.eof() && c.peek() != '\n') c.i++; }
                else if (ch == '\n' && depth == 0) done = true;
                else c.i++;
            }
            if (depth != 0) return "unterminated value of `" + key + "`";
            continue;
        }
        if (c.eof() || c.peek() != '[') return "invalid type: expected an array of \"X\" and \" \" characters forming a grid";
        c.i++;
        for (;;) {
            c.skip(true);
            if (c.eof()) return "unterminated array";
            if (c.peek() == ']') { c.i++; break; }
            char q = c.peek();
            if (q != '"' && q != '\'') return "invalid type: expected a string row";
            c.i++;
            std::string row;
            for (;;) {
                if (c.eof() || c.peek() == '\n') return "unterminated string";
                char ch = c.s[c.i++];
                if (ch == q) break;
                if (q == '"' && ch == '\\') return "escape sequences are not supported in grid rows";
                if (ch != ' ' && ch != 'X')  // world.rs:55-59
                    return std::string("invalid value: character `") + ch + "`, expected `X` or ` `";
                row.push_back(ch);
            }
            rows.push_back(std::move(row));
            c.skip(true);
            if (!c.eof() && c.peek() == ',') c.i++;
        }
        have_grid = true;
    }
    if (!have_grid) return "missing field `grid`";
    if (rows.empty()) return "invalid length 0, expected 1 or more";  // world.rs:63-65
    size_t width = 0;
    for (auto& r : rows) width = std::max(width, r.size());
    w = (int)width;
    h = (int)rows.size();
    grid.assign(width * rows.size(), 0);  // left-aligned, padded false (world.rs:82-86)
    ragged = false;
    for (size_t y = 0; y < rows.size(); y++) {
        ragged = ragged || rows[y].size() != width;
        for (size_t x = 0; x < rows[y].size(); x++) grid[y * width + x] = rows[y][x] == 'X';
    }
    return "";
}

std::string world_to_toml(const uint8_t* grid, int w, int h) {  // world.rs:21-40
    std::string s = "[world]\ngrid = [\n";
    for (int y = 0; y < h; y++) {
        s += "    \"";
        for (int x = 0; x < w; x++) s.push_back(grid[(size_t)y * w + x] ? 'X' : ' ');
        s += "\",\n";
    }
    return s + "]\n";
}

static inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void synthetic_world(int w, int h, uint64_t seed, uint64_t t, uint32_t density_q24, uint8_t* grid) {
    uint64_t base = seed * 0x9E3779B97F4A7C15ull + (t << 20);
    for (int i = 0; i < w * h; i++) grid[i] = (splitmix64(base + (uint64_t)i) >> 40) < density_q24;
}

// ------------------------------------------------------------------------------------------------ platform
bool platform_overlaps(const tss_platform& a, const tss_platform& b) {  // platform.rs:86-97 (inclusive corners)
    Dims da = platform_dims(a), db = platform_dims(b);
    if (da.w <= 0 || da.h <= 0 || db.w <= 0 || db.h <= 0) return false;  // corner_point_incl() is None
    return b.x + db.w - 1 >= a.x && b.y + db.h - 1 >= a.y && b.x <= a.x + da.w - 1 && b.y <= a.y + da.h - 1;
}

int trivial_optimization(const uint8_t* grid, int w, int h, tss_platform* plats, int n) {  // platform_layout.rs:151-172
    int kept = 0;
    for (int i = 0; i < n; i++) {
        Dims d = platform_dims(plats[i]);
        bool any = false;
        for (int dy = 0; dy < d.h && !any; dy++)
            for (int dx = 0; dx < d.w && !any; dx++) {
                int x = plats[i].x + dx, y = plats[i].y + dy;
                any = x >= 0 && y >= 0 && x < w && y < h && grid[(size_t)y * w + x];
            }
        if (any) plats[kept++] = plats[i];
    }
    return kept;
}

long total_weight(const tss_platform* plats, int n, const std::vector<PlatformLimits::Entry>& weights) {  // :174-183
    long sum = 0;
    for (int i = 0; i < n; i++)
        for (auto& e : weights)
            if (dims_le(e.def, Dims{plats[i].def_w, plats[i].def_h})) sum += e.value;
    return sum;
}

// ------------------------------------------------------------------------------------------------ encoder
namespace {
// The platform-selection DAG of encoder.rs:281-426 reduced to the three tables encode() consumes.
struct DagTables {
    std::vector<std::pair<int, int>> impl;  // (smaller key, larger key): transitive reduction of the dims order
    struct Sibling { int a, b; std::vector<int> lcub; };
    std::vector<Sibling> siblings;          // encoder.rs:460-489
    struct PointEdge { int dx, dy, key; };
    std::vector<PointEdge> points;          // point offset -> minimal platform(s) containing it (encoder.rs:368-373)
};

DagTables build_dag(const std::vector<Dims>& keys) {
    DagTables t;
    int K = (int)keys.size();
    auto lt = [&](int a, int b) { return dims_partial_cmp(keys[a], keys[b]) == -1; };
    std::vector<std::vector<int>> succ(K);  // reduced successors
    for (int a = 0; a < K; a++)
        for (int b = 0; b < K; b++) {
            if (!lt(a, b)) continue;
            bool via = false;
            for (int c = 0; c < K && !via; c++) via = lt(a, c) && lt(c, b);
            if (!via) { t.impl.push_back({a, b}); succ[a].push_back(b); }
        }
    for (int s = 0; s < K; s++)
        for (size_t i = 0; i < succ[s].size(); i++)
            for (size_t j = i + 1; j < succ[s].size(); j++) {
                int a = succ[s][i], b = succ[s][j];
                std::vector<int> common;
                for (int c = 0; c < K; c++)
                    if (lt(a, c) && lt(b, c)) common.push_back(c);
                DagTables::Sibling sib{a, b, {}};
                for (int c : common) {  // keep the least common upper bounds (encoder.rs:419-425)
                    bool above_other = false;
                    for (int d : common) above_other = above_other || lt(d, c);
                    if (!above_other) sib.lcub.push_back(c);
                }
                t.siblings.push_back(std::move(sib));
            }
    int mw = 1, mh = 1;
    for (Dims d : keys) { mw = std::max(mw, d.w); mh = std::max(mh, d.h); }
    for (int py = 0; py < mh; py++)
        for (int px = 0; px < mw; px++)
            for (int k = 0; k < K; k++) {
                if (!(px < keys[k].w && py < keys[k].h)) continue;
                bool minimal = true;
                for (int c = 0; c < K && minimal; c++) minimal = !(px < keys[c].w && py < keys[c].h && lt(c, k));
                if (minimal) t.points.push_back({px, py, k});
            }
    return t;
}
}  // namespace

int Encoding::key_index(Dims d) const {
    for (int k = 0; k < K(); k++)
        if (keys[k] == d) return k;
    return -1;
}

std::string Encoding::encode(const uint8_t* grid, int w, int h, const std::vector<Dims>& defs, Encoding& e) {
    if (w <= 0 || h <= 0 || !grid) return "empty grid";
    e = Encoding();
    e.w = w;
    e.h = h;
    e.defs = defs;
    for (size_t i = 0; i < defs.size(); i++) {  // encoder.rs:121-130: both orientations of every def
        if (defs[i].w <= 0 || defs[i].h <= 0) return "empty platform dimensions";
        for (Dims k : {defs[i], Dims{defs[i].h, defs[i].w}})
            if (e.key_index(k) < 0) { e.keys.push_back(k); e.key_def.push_back((int)i); }
    }
    const int one = e.key_index(Dims{1, 1});
    if (one < 0) return "the platform set must contain 1x1 (the overlap clauses are expressed through it, src/encoder.rs:559-571)";
    const int K = e.K(), tiles = w * h;
    Cnf& f = e.base;
    // encoder.rs:184-195: per tile, K platform vars then 4 terrain-layer vars iff ceiling
    e.plat_var.assign((size_t)tiles * K, 0);
    e.terr_var.assign((size_t)tiles * 4, 0);
    for (int t = 0; t < tiles; t++) {
        for (int k = 0; k < K; k++) e.plat_var[(size_t)t * K + k] = f.new_var();
        if (grid[t])
            for (int l = 0; l < 4; l++) e.terr_var[(size_t)t * 4 + l] = f.new_var();
    }
    const DagTables dag = build_dag(e.keys);
    auto in = [&](int x, int y) { return x >= 0 && y >= 0 && x < w && y < h; };
    auto P = [&](int x, int y, int k) { return e.plat_var[(size_t)(y * w + x) * K + k]; };
    auto T = [&](int x, int y, int l) { return e.terr_var[(size_t)(y * w + x) * 4 + l]; };
    static const int NX[4] = {1, 0, -1, 0}, NY[4] = {0, 1, 0, -1};  // point.rs:46-53

    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            for (auto& [s, l] : dag.impl) f.add2(-P(x, y, l), P(x, y, s));  // larger -> smaller (encoder.rs:449-458)
            for (auto& sib : dag.siblings) {                                 // (a & b) -> lcub (encoder.rs:460-489)
                f.lits.push_back(-P(x, y, sib.a));
                f.lits.push_back(-P(x, y, sib.b));
                for (int c : sib.lcub) f.lits.push_back(P(x, y, c));
                f.close();
            }
            if (grid[y * w + x]) {
                f.lits.push_back(-T(x, y, 3));  // T3 -> platforms reaching this tile (encoder.rs:500-516)
                for (auto& pe : dag.points)
                    if (in(x - pe.dx, y - pe.dy)) f.lits.push_back(P(x - pe.dx, y - pe.dy, pe.key));
                f.close();
                for (int l = 0; l < 3; l++) {  // T_l -> T_{l+1} of a ceiling neighbour or itself (encoder.rs:520-537)
                    f.lits.push_back(-T(x, y, l));
                    for (int d = 0; d < 4; d++)
                        if (in(x + NX[d], y + NY[d]) && grid[(y + NY[d]) * w + x + NX[d]]) f.lits.push_back(T(x + NX[d], y + NY[d], l + 1));
                    f.lits.push_back(T(x, y, l + 1));
                    f.close();
                }
                f.add1(T(x, y, 0));  // encoder.rs:543
            }
            for (auto& pe : dag.points)  // another anchor inside this platform (encoder.rs:559-571)
                if ((pe.dx || pe.dy) && in(x + pe.dx, y + pe.dy)) f.add2(-P(x, y, pe.key), -P(x + pe.dx, y + pe.dy, one));
            for (auto& top : dag.points) {  // top edge x left edge crossings (encoder.rs:576-596)
                if (top.dy != 0 || top.dx == 0) continue;
                for (auto& left : dag.points) {
                    if (left.dx != 0 || left.dy == 0) continue;
                    int qx = x + top.dx, qy = y - left.dy;
                    if (in(qx, qy)) f.add2(-P(x, y, top.key), -P(qx, qy, left.key));
                }
            }
            for (auto& pe : dag.points)  // encoder.rs:601-609
                if (!in(x + pe.dx, y + pe.dy)) f.add1(-P(x, y, pe.key));
        }
    return "";
}

// rustsat `into_cnf` stand-ins (source not in the reference tree; clause sets are this repo's own, parity
// unpinned): totalizer for cardinality, generalized totalizer for pseudo-boolean upper bounds.
namespace {
std::vector<int> totalizer_node(Cnf& f, const std::vector<int>& lits, size_t lo, size_t hi, size_t cap) {
    if (hi - lo == 1) return {lits[lo]};
    size_t mid = lo + (hi - lo) / 2;
    std::vector<int> a = totalizer_node(f, lits, lo, mid, cap), b = totalizer_node(f, lits, mid, hi, cap);
    std::vector<int> o(std::min(cap, a.size() + b.size()));
    for (int& v : o) v = f.new_var();
    for (size_t i = 0; i <= a.size(); i++)
        for (size_t j = 0; j <= b.size(); j++) {
            size_t s = i + j;
            if (s == 0 || s > o.size()) continue;
            if (i) f.lits.push_back(-a[i - 1]);
            if (j) f.lits.push_back(-b[j - 1]);
            f.lits.push_back(o[s - 1]);
            f.close();
        }
    return o;
}
void card_upper_bound(Cnf& f, const std::vector<int>& lits, size_t k) {
    if (k >= lits.size()) return;
    if (k == 0) { for (int l : lits) f.add1(-l); return; }
    std::vector<int> o = totalizer_node(f, lits, 0, lits.size(), k + 1);
    f.add1(-o[k]);
}
using WeightedOut = std::map<long, int>;
WeightedOut gte_node(Cnf& f, const std::vector<std::pair<int, long>>& wl, size_t lo, size_t hi, long cap) {
    if (hi - lo == 1) return {{std::min(wl[lo].second, cap), wl[lo].first}};
    size_t mid = lo + (hi - lo) / 2;
    WeightedOut a = gte_node(f, wl, lo, mid, cap), b = gte_node(f, wl, mid, hi, cap), o;
    auto out = [&](long s) { s = std::min(s, cap); auto it = o.find(s); if (it == o.end()) it = o.emplace(s, f.new_var()).first; return it->second; };
    for (auto& [wa, va] : a) f.add2(-va, out(wa));
    for (auto& [wb, vb] : b) f.add2(-vb, out(wb));
    for (auto& [wa, va] : a)
        for (auto& [wb, vb] : b) f.add3(-va, -vb, out(wa + wb));
    return o;
}
void pb_upper_bound(Cnf& f, const std::vector<std::pair<int, long>>& wl, long limit) {
    std::vector<std::pair<int, long>> pos;
    for (auto [l, wgt] : wl) {
        if (wgt == 0) continue;
        if (wgt < 0) { limit -= wgt; pos.push_back({-l, -wgt}); } else pos.push_back({l, wgt});
    }
    if (limit < 0) { f.close(); return; }  // empty clause
    long total = 0;
    for (auto& p : pos) total += p.second;
    if (pos.empty() || total <= limit) return;
    for (auto& [s, v] : gte_node(f, pos, 0, pos.size(), limit + 1))
        if (s > limit) f.add1(-v);
}
}  // namespace

Cnf Encoding::with_limits(const PlatformLimits& limits) const {  // encoder.rs:619-667
    Cnf f = base;
    const int tiles = w * h, K_ = K();
    std::vector<Dims> types;  // card_limits.keys().chain(weights.keys()).unique()
    auto note = [&](Dims d) { if (std::find(types.begin(), types.end(), d) == types.end()) types.push_back(d); };
    for (auto& c : limits.card_limits) note(c.def);
    for (auto& c : limits.weights) note(c.def);
    std::vector<std::pair<std::vector<int>, size_t>> cards;
    std::vector<std::pair<int, long>> weighted;
    for (Dims type : types) {
        std::vector<int> lits;
        int k0 = key_index(type), k1 = key_index(Dims{type.h, type.w});
        if (type.w != type.h) {  // rectangular: one fresh var per tile implied by both orientations (encoder.rs:629-641)
            for (int t = 0; t < tiles; t++) {
                int lv = f.new_var();
                lits.push_back(lv);
                if (k0 >= 0) f.add2(-plat_var[(size_t)t * K_ + k0], lv);
                if (k1 >= 0) f.add2(-plat_var[(size_t)t * K_ + k1], lv);
            }
        } else if (k0 >= 0) {  // that dims' var at EVERY tile incl. non-ceiling (encoder.rs:643-646)
            for (int t = 0; t < tiles; t++) lits.push_back(plat_var[(size_t)t * K_ + k0]);
        }
        for (auto& c : limits.card_limits)
            if (c.def == type) cards.push_back({lits, (size_t)std::max(0l, c.value)});
        if (limits.has_weight_limit)
            for (auto& c : limits.weights)
                if (c.def == type)
                    for (int l : lits) weighted.push_back({l, c.value});
    }
    for (auto& [lits, k] : cards) card_upper_bound(f, lits, k);
    if (limits.has_weight_limit) pb_upper_bound(f, weighted, limits.weight_limit);
    return f;
}

std::vector<tss_platform> Encoding::layout_from_assignment(const uint8_t* a, int n) const {  // platform_layout.rs:26-52
    std::vector<tss_platform> out;
    const int K_ = K(), tiles = w * h;
    for (int t = 0; t < tiles; t++) {  // variables ascend with (tile, key): same visiting order as Assignment::iter()
        int best = -1;
        for (int k = 0; k < K_; k++) {
            int v = plat_var[(size_t)t * K_ + k];
            if (v >= n || a[v] != 1) continue;
            // keep the first, replace only by a strictly larger def (compares canonical def dims)
            if (best < 0 || dims_partial_cmp(defs[key_def[best]], defs[key_def[k]]) == -1) best = k;
        }
        if (best < 0) continue;
        Dims def = defs[key_def[best]];
        out.push_back(tss_platform{t % w, t / w, def.w, def.h, !(def == keys[best])});
    }
    return out;
}

}  // namespace tss
