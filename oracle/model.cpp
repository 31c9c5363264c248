// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  Restatement of the reference's domain model,
// encoder and layout validator.  Written to follow the reference's structure (sets/maps, per-tile loops),
// not for speed: it is the checker, the CUDA path in timberborn_support_solver_b200/ is the product.
#include "oracle.hpp"

#include <algorithm>
#include <array>
#include <cctype>
#include <cstring>
#include <sstream>

namespace tsso {

// ---------------------------------------------------------------- math
void neighbors(Point p, Point out[4]) {  // src/math/point.rs:46-53
    out[0] = {p.x + 1, p.y};
    out[1] = {p.x, p.y + 1};
    out[2] = {p.x - 1, p.y};
    out[3] = {p.x, p.y - 1};
}

std::vector<Point> iter_within_manhattan(Point c, unsigned dist) {  // src/math/point.rs:110-132
    std::vector<Point> out;
    long d = (long)dist;
    Point rel{0, -d};
    while (rel.y <= d) {
        out.push_back(rel + c);
        rel.x += 1;
        if (std::labs(rel.x) + std::labs(rel.y) > d) {
            rel.y += 1;
            rel.x = -(d - std::labs(rel.y));
        }
    }
    return out;
}

bool Dims::operator<(const Dims& o) const {
    // total order for containers; empty dims collapse to one key like the reference's Hash/Eq (dimensions.rs:68-72,116-123)
    unsigned long aw = empty() ? 0 : width, ah = empty() ? 0 : height;
    unsigned long bw = o.empty() ? 0 : o.width, bh = o.empty() ? 0 : o.height;
    return aw != bw ? aw < bw : ah < bh;
}

POrd partial_cmp(Dims a, Dims b) {  // src/math/dimensions.rs:74-114
    if (a.empty() && b.empty()) return POrd::Equal;
    if (a.empty()) return POrd::Less;
    if (b.empty()) return POrd::Greater;
    int cw = a.width < b.width ? -1 : (a.width > b.width ? 1 : 0);
    int ch = a.height < b.height ? -1 : (a.height > b.height ? 1 : 0);
    if (cw == 0 && ch == 0) return POrd::Equal;
    if ((cw < 0 && ch > 0) || (cw > 0 && ch < 0)) return POrd::None;
    if (cw < 0 || ch < 0) return POrd::Less;
    return POrd::Greater;
}

std::vector<Point> iter_within(Dims d) {  // src/math/dimensions.rs:138-156
    std::vector<Point> out;
    Point cur{0, 0};
    while (cur.y < (long)d.height) {
        // NB: like the reference, a zero-width non-zero-height rectangle would yield x=0 points; unused.
        out.push_back(cur);
        cur.x += 1;
        if (cur.x >= (long)d.width) { cur.x = 0; cur.y += 1; }
    }
    return out;
}

// ---------------------------------------------------------------- world (TOML subset)
// The reference uses the `toml` crate + serde (crates/repl/src/main.rs:272-278).  The project files are
// `[world]` + `grid = [ "..", ... ]`; this parser accepts exactly that subset: one table header, one key,
// an array of basic ("...") or literal ('...') strings, comments, trailing comma.
std::string parse_world_toml(const std::string& text, WorldGrid& out, bool* ragged) {
    size_t i = 0, n = text.size();
    auto skip_ws = [&](bool newlines) {
        while (i < n) {
            char c = text[i];
            if (c == ' ' || c == '\t' || c == '\r' || (newlines && c == '\n')) { i++; continue; }
            if (c == '#') { while (i < n && text[i] != '\n') i++; continue; }
            break;
        }
    };
    bool in_world = false, have_grid = false;
    std::vector<std::vector<bool>> rows;
    while (true) {
        skip_ws(true);
        if (i >= n) break;
        if (text[i] == '[') {
            size_t j = text.find(']', i);
            if (j == std::string::npos) return "unterminated table header";
            std::string name = text.substr(i + 1, j - i - 1);
            name.erase(std::remove_if(name.begin(), name.end(), [](char c) { return c == ' ' || c == '\t'; }), name.end());
            in_world = (name == "world");
            i = j + 1;
            continue;
        }
        size_t ks = i;
        while (i < n && (std::isalnum((unsigned char)text[i]) || text[i] == '_' || text[i] == '-')) i++;
        std::string key = text.substr(ks, i - ks);
        if (key.empty()) return "expected a key";
        skip_ws(false);
        if (i >= n || text[i] != '=') return "expected `=` after key `" + key + "`";
        i++;
        skip_ws(false);
        if (!(in_world && key == "grid")) {
            // not a field of Project / World: serde skips it (no deny_unknown_fields, src/lib.rs:14-17, src/world.rs:13-16).
            // Walk over the value: quoted strings, bracketed values over several lines, or the rest of the line.
            std::vector<char> open;
            for (; i < n; i++) {
                const char c = text[i];
                if (c == '#') { while (i < n && text[i] != '\n') i++; i--; continue; }
                if (c == '"' || c == '\'') {
                    size_t j = i + 1;
                    while (j < n && text[j] != c) j += (c == '"' && text[j] == '\\') ? 2 : 1;
                    if (j >= n) return "unterminated string";
                    i = j;
                    continue;
                }
                if (c == '[' || c == '{') open.push_back(c);
                else if (c == ']' || c == '}') { if (open.empty()) return "unbalanced bracket in the value of `" + key + "`"; open.pop_back(); }
                else if (c == '\n' && open.empty()) break;
            }
            if (!open.empty()) return "unterminated value of `" + key + "`";
            continue;
        }
        if (i >= n || text[i] != '[') return "invalid type: expected an array of \"X\" and \" \" characters forming a grid";
        i++;
        while (true) {
            skip_ws(true);
            if (i >= n) return "unterminated array";
            if (text[i] == ']') { i++; break; }
            char q = text[i];
            if (q != '"' && q != '\'') return "invalid type: expected a string row";
            i++;
            std::vector<bool> row;
            while (true) {
                if (i >= n || text[i] == '\n') return "unterminated string";
                char c = text[i++];
                if (c == q) break;
                if (q == '"' && c == '\\') return "escape sequences are not supported in grid rows";
                if (c == ' ') row.push_back(false);          // world.rs:56
                else if (c == 'X') row.push_back(true);      // world.rs:57
                else return std::string("invalid value: character `") + c + "`, expected `X` or ` `";  // world.rs:58
            }
            rows.push_back(std::move(row));
            skip_ws(true);
            if (i < n && text[i] == ',') i++;
        }
        have_grid = true;
    }
    if (!have_grid) return "missing field `grid`";
    if (rows.empty()) return "invalid length 0, expected 1 or more";  // world.rs:63-65
    size_t w = 0;
    for (auto& r : rows) w = std::max(w, r.size());
    out.dims = Dims{w, rows.size()};
    out.data.assign(w * rows.size(), false);
    bool rg = false;
    for (size_t y = 0; y < rows.size(); y++) {
        if (rows[y].size() != w) rg = true;
        for (size_t x = 0; x < rows[y].size(); x++) out.data[y * w + x] = rows[y][x];
    }
    if (ragged) *ragged = rg;
    return "";
}

std::string world_to_toml(const WorldGrid& g) {  // world.rs:21-40
    std::ostringstream os;
    os << "[world]\ngrid = [\n";
    for (size_t y = 0; y < g.dims.height; y++) {
        os << "    \"";
        for (size_t x = 0; x < g.dims.width; x++) os << (g.data[y * g.dims.width + x] ? 'X' : ' ');
        os << "\",\n";
    }
    os << "]\n";
    return os.str();
}

// ---------------------------------------------------------------- platform
const PlatformDef PLATFORMS_DEFAULT[8] = {  // platform.rs:23-32
    {{1, 1}}, {{1, 2}}, {{1, 3}}, {{1, 4}}, {{1, 5}}, {{1, 6}}, {{3, 3}}, {{5, 5}},
};

bool Platform::overlaps(const Platform& o) const {  // platform.rs:75-97
    Dims a = dims(), b = o.dims();
    if (a.empty() || b.empty()) return false;  // corner_point_incl() == None
    Point sn = point, sf = {point.x + (long)a.width - 1, point.y + (long)a.height - 1};
    Point on = o.point, of = {o.point.x + (long)b.width - 1, o.point.y + (long)b.height - 1};
    return of.x >= sn.x && of.y >= sn.y && on.x <= sf.x && on.y <= sf.y;
}

bool Platform::operator<(const Platform& o) const {
    if (!(point == o.point)) return point < o.point;
    if (!(def == o.def)) return def < o.def;
    return rotated < o.rotated;
}

// ---------------------------------------------------------------- EncodingVars
static EncodingVars make_vars(const std::vector<PlatformDef>& defs, const WorldGrid& terrain, SatInstance& inst) {
    EncodingVars v;
    // encoder.rs:121-130 dims_platform_map: both orientations of every def.
    for (const auto& d : defs) {
        for (Dims k : {d.dims, d.dims.flipped()}) {
            if (!v.dim_map.count(k)) { v.dim_map.emplace(k, d); v.dim_keys.push_back(k); }
        }
    }
    // encoder.rs:191-195: per tile (row-major) one var per dims key, then 4 terrain vars iff ceiling.
    v.grid.dims = terrain.dims;
    v.grid.data.resize(terrain.data.size());
    for (size_t i = 0; i < terrain.data.size(); i++) {
        EncodingTileVars& tv = v.grid.data[i];
        for (Dims k : v.dim_keys) tv.dims_vars[k] = inst.new_var();
        if (terrain.data[i]) {
            std::array<int, TERRAIN_SUPPORT_DISTANCE> t{};
            for (int l = 0; l < TERRAIN_SUPPORT_DISTANCE; l++) t[l] = inst.new_var();
            tv.terrain = t;
        }
    }
    // encoder.rs:196-204 var_map
    for (size_t i = 0; i < v.grid.data.size(); i++) {
        Point p = v.grid.index_to_point(i);
        for (auto& [d, var] : v.grid.data[i].dims_vars) v.var_map[var] = EncodedItem{true, p, d, 0};
        if (v.grid.data[i].terrain)
            for (int l = 0; l < TERRAIN_SUPPORT_DISTANCE; l++) v.var_map[(*v.grid.data[i].terrain)[l]] = EncodedItem{false, p, {}, l};
    }
    return v;
}

std::optional<Platform> EncodingVars::var_to_platform(int var) const {  // encoder.rs:232-249
    auto it = var_map.find(var);
    if (it == var_map.end() || !it->second.is_platform) return std::nullopt;
    const PlatformDef& def = dim_map.at(it->second.dims);
    bool rotated = !(def.dims.width == it->second.dims.width && def.dims.height == it->second.dims.height);
    return Platform{it->second.point, def, rotated};
}

// ---------------------------------------------------------------- EncodingDag
static POrd node_cmp(const EncodingNode& a, const EncodingNode& b) {  // encoder.rs:288-303
    if (a.is_platform && b.is_platform) return partial_cmp(a.dims, b.dims);
    if (a.is_platform && !b.is_platform) return a.dims.contains(b.point) ? POrd::Greater : POrd::None;
    if (!a.is_platform && b.is_platform) return b.dims.contains(a.point) ? POrd::Less : POrd::None;
    return a.point == b.point ? POrd::Equal : POrd::None;
}

EncodingDag::EncodingDag(const std::vector<Dims>& platform_dims) {
    Dims mx{1, 1};  // encoder.rs:318-320
    for (Dims d : platform_dims) mx = Dims{std::max(mx.width, d.width), std::max(mx.height, d.height)};
    std::vector<EncodingNode> all;
    for (Dims d : platform_dims) all.push_back({true, d, {}});
    for (Point p : iter_within(mx)) all.push_back({false, {}, p});
    // encoder.rs:136-156 dag_by_partial_ord: edge other -> this iff other < this
    size_t n = all.size();
    std::vector<std::vector<bool>> adj(n, std::vector<bool>(n, false));
    for (size_t t = 0; t < n; t++)
        for (size_t o = 0; o < n; o++)
            if (node_cmp(all[o], all[t]) == POrd::Less) adj[o][t] = true;
    // encoder.rs:331 retain nodes with at least one undirected neighbour
    std::vector<int> keep;
    for (size_t i = 0; i < n; i++) {
        bool any = false;
        for (size_t j = 0; j < n; j++) any = any || adj[i][j] || adj[j][i];
        if (any) keep.push_back((int)i);
    }
    size_t m = keep.size();
    nodes.resize(m);
    std::vector<std::vector<bool>> g(m, std::vector<bool>(m, false));
    for (size_t i = 0; i < m; i++) {
        nodes[i] = all[keep[i]];
        for (size_t j = 0; j < m; j++) g[i][j] = adj[keep[i]][keep[j]];
    }
    // encoder.rs:337-344: transitive closure (Warshall) and reduction (edge i->j survives iff there is no
    // k with i ->+ k ->+ j).  petgraph's tred does the same on the toposorted adjacency list.
    closure = g;
    for (size_t k = 0; k < m; k++)
        for (size_t i = 0; i < m; i++)
            if (closure[i][k])
                for (size_t j = 0; j < m; j++)
                    if (closure[k][j]) closure[i][j] = true;
    reduced.assign(m, std::vector<bool>(m, false));
    for (size_t i = 0; i < m; i++)
        for (size_t j = 0; j < m; j++) {
            if (!g[i][j]) continue;
            bool via = false;
            for (size_t k = 0; k < m && !via; k++) via = closure[i][k] && closure[k][j];
            reduced[i][j] = !via;
        }
}

std::vector<std::pair<Dims, Dims>> EncodingDag::platform_edges_reduced() const {
    std::vector<std::pair<Dims, Dims>> out;
    for (size_t i = 0; i < nodes.size(); i++)
        for (size_t j = 0; j < nodes.size(); j++)
            if (reduced[i][j] && nodes[i].is_platform && nodes[j].is_platform) out.push_back({nodes[i].dims, nodes[j].dims});
    return out;
}

std::vector<std::pair<Point, Dims>> EncodingDag::point_platform_edges_reduced() const {
    std::vector<std::pair<Point, Dims>> out;
    for (size_t i = 0; i < nodes.size(); i++)
        for (size_t j = 0; j < nodes.size(); j++)
            if (reduced[i][j] && !nodes[i].is_platform && nodes[j].is_platform) out.push_back({nodes[i].point, nodes[j].dims});
    return out;
}

std::vector<EncodingDag::Sibling> EncodingDag::sibling_clauses() const {
    std::vector<Sibling> out;
    size_t m = nodes.size();
    for (size_t s = 0; s < m; s++) {
        if (!nodes[s].is_platform) continue;
        std::vector<size_t> targets;  // encoder.rs:375-399 reduced out-edges of a platform node
        for (size_t j = 0; j < m; j++)
            if (reduced[s][j]) targets.push_back(j);
        for (size_t ia = 0; ia < targets.size(); ia++)
            for (size_t ib = ia + 1; ib < targets.size(); ib++) {  // itertools tuple_combinations
                size_t a = targets[ia], b = targets[ib];
                std::vector<size_t> common;  // encoder.rs:401-417
                for (size_t j = 0; j < m; j++)
                    if (closure[a][j] && closure[b][j] && nodes[j].is_platform) common.push_back(j);
                Sibling sib{nodes[a].dims, nodes[b].dims, {}};
                for (size_t c : common) {  // encoder.rs:419-425 maximal_from: drop n if some m ->+ n
                    bool dominated = false;
                    for (size_t d : common) dominated = dominated || closure[d][c];
                    if (!dominated) sib.lcub.push_back(nodes[c].dims);
                }
                out.push_back(std::move(sib));
            }
    }
    return out;
}

// ---------------------------------------------------------------- Encoding::encode  (encoder.rs:435-613)
Encoding Encoding::encode(const std::vector<PlatformDef>& defs, const WorldGrid& terrain) {
    Encoding e;
    e.vars = make_vars(defs, terrain, e.instance);
    SatInstance& inst = e.instance;
    const EncodingVars& vars = e.vars;
    EncodingDag dag(vars.dim_keys);
    auto plat_edges = dag.platform_edges_reduced();
    auto pt_edges = dag.point_platform_edges_reduced();
    auto siblings = dag.sibling_clauses();
    const Dims one{1, 1};
    const int T_LAST = TERRAIN_SUPPORT_DISTANCE - 1;

    for (Point cur : iter_within(terrain.dims)) {
        const EncodingTileVars& cv = *vars.grid.get(cur);

        // ===== Platform selection DAG =====  encoder.rs:449-458: larger -> smaller
        for (auto& [smaller, larger] : plat_edges) inst.add_lit_impl_lit(cv.dims_vars.at(larger), cv.dims_vars.at(smaller), F_DAG_IMPL);

        // encoder.rs:460-489: (a & b) -> (least common upper bounds)
        for (auto& s : siblings) {
            std::vector<int> rhs;
            for (Dims d : s.lcub) rhs.push_back(cv.dims_vars.at(d));
            inst.add_cube_impl_clause({cv.dims_vars.at(s.a), cv.dims_vars.at(s.b)}, rhs, F_DAG_SIBLING);
        }

        // ===== Platform-terrain clauses =====  encoder.rs:500-516
        if (cv.terrain) {
            std::vector<int> plats;
            for (auto& [offset, dims] : pt_edges)
                if (const EncodingTileVars* o = vars.grid.get(cur - offset)) plats.push_back(o->dims_vars.at(dims));
            inst.add_lit_impl_clause((*cv.terrain)[T_LAST], plats, F_T3_PLATFORM);
        }

        // ===== Terrain support =====  encoder.rs:520-544
        if (cv.terrain) {
            std::vector<std::array<int, TERRAIN_SUPPORT_DISTANCE>> nb;
            Point ns[4];
            neighbors(cur, ns);
            for (Point q : ns)
                if (const EncodingTileVars* o = vars.grid.get(q))
                    if (o->terrain) nb.push_back(*o->terrain);
            nb.push_back(*cv.terrain);  // chain(iter::once(point_terrain))
            for (int i = 0; i + 1 < TERRAIN_SUPPORT_DISTANCE; i++) {
                int j = i + 1;
                std::vector<int> rhs;
                for (auto& t : nb) rhs.push_back(t[j]);
                inst.add_lit_impl_clause((*cv.terrain)[i], rhs, F_LAYER);
            }
            inst.add_unit((*cv.terrain)[0], F_UNIT_T0);
        }

        // ===== Platform overlap =====  encoder.rs:559-571: top-left corner of another platform inside this one
        for (auto& [offset, dims] : pt_edges) {
            if (offset == Point{0, 0}) continue;
            const EncodingTileVars* o = vars.grid.get(cur + offset);
            if (!o) continue;
            inst.add_lit_impl_lit(cv.dims_vars.at(dims), -o->dims_vars.at(one), F_OVERLAP_ANCHOR);  // assumes 1x1 exists
        }
        // encoder.rs:576-596: top-edge point (x;0) against platforms reaching it through their left edge (0;y)
        for (auto& [o1, d1] : pt_edges) {
            if (o1 == Point{0, 0} || o1.y != 0) continue;
            for (auto& [o2, d2] : pt_edges) {
                if (o2 == Point{0, 0} || o2.x != 0) continue;
                const EncodingTileVars* o = vars.grid.get(cur + o1 - o2);
                if (!o) continue;
                inst.add_lit_impl_lit(cv.dims_vars.at(d1), -o->dims_vars.at(d2), F_OVERLAP_CROSS);
            }
        }

        // ===== Out-of-bounds platforms =====  encoder.rs:601-609
        for (auto& [offset, dims] : pt_edges)
            if (!terrain.dims.contains(cur + offset)) inst.add_unit(-cv.dims_vars.at(dims), F_OOB);
    }
    return e;
}

// ---------------------------------------------------------------- PlatformLayout
PlatformLayout PlatformLayout::from_assignment(const Assignment& a, const EncodingVars& vars) {  // platform_layout.rs:26-52
    PlatformLayout out;
    // rustsat Assignment::iter() walks assigned vars in index order; only positive platform lits matter.
    for (int var = 1; var < (int)a.size(); var++) {
        if (a[var] != 1) continue;
        auto plat = vars.var_to_platform(var);
        if (!plat) continue;
        auto it = out.platforms.find(plat->point);
        if (it == out.platforms.end()) out.platforms.emplace(plat->point, *plat);
        else if (partial_cmp(it->second.def.dims, plat->def.dims) == POrd::Less) it->second = *plat;  // update only if larger
    }
    return out;
}

std::map<PlatformDef, size_t> PlatformLayout::platform_stats() const {  // :66-79
    std::map<PlatformDef, size_t> m;
    for (auto& [p, plat] : platforms) m[plat.def]++;
    return m;
}

ValidationResult PlatformLayout::validate(const WorldGrid& world) const {  // :85-149
    struct Tile { std::optional<bool> terrain_supported; const Platform* occupied_by = nullptr; };
    ValidationResult res;
    Grid<Tile> tracking;
    tracking.dims = world.dims;
    tracking.data.resize(world.data.size());
    for (size_t i = 0; i < world.data.size(); i++)
        tracking.data[i].terrain_supported = world.data[i] ? std::optional<bool>(false) : std::nullopt;

    for (auto& [anchor, plat] : platforms) {
        for (Point offset : iter_within(plat.dims())) {
            Point point = offset + plat.point;
            if (Tile* tile = tracking.get_mut(point)) {
                if (tile->occupied_by) {
                    res.overlapping_platforms.insert(plat);
                    res.overlapping_platforms.insert(*tile->occupied_by);
                } else {
                    tile->occupied_by = &plat;
                }
                if (tile->terrain_supported) tile->terrain_supported = true;  // only terrain can be supported
            } else {
                res.out_of_bounds_platforms.insert(plat);
            }
        }
    }
    // Extend terrain support: 3 rounds of 4-neighbour spreading restricted to terrain  (:127-141)
    for (int round = 0; round < TERRAIN_SUPPORT_DISTANCE - 1; round++) {
        std::set<Point> supported_set;
        for (size_t i = 0; i < tracking.data.size(); i++)
            if (tracking.data[i].terrain_supported == std::optional<bool>(true)) {
                Point ns[4];
                neighbors(tracking.index_to_point(i), ns);
                for (Point q : ns) supported_set.insert(q);
            }
        for (Point p : supported_set)
            if (Tile* tile = tracking.get_mut(p))
                if (tile->terrain_supported) tile->terrain_supported = true;
    }
    for (size_t i = 0; i < tracking.data.size(); i++)
        if (tracking.data[i].terrain_supported == std::optional<bool>(false)) res.unsupported_terrain.insert(tracking.index_to_point(i));
    return res;
}

void PlatformLayout::run_trivial_optimization(const WorldGrid& world) {  // :151-172
    for (auto it = platforms.begin(); it != platforms.end();) {
        bool any = false;
        for (Point offset : iter_within(it->second.dims())) {
            const uint8_t* b = world.get(it->first + offset);
            if (b && *b) { any = true; break; }
        }
        it = any ? std::next(it) : platforms.erase(it);
    }
}

static bool dims_le(Dims a, Dims b) {  // PartialOrd `<=`
    POrd o = partial_cmp(a, b);
    return o == POrd::Less || o == POrd::Equal;
}

long PlatformLayout::total_weight(const std::vector<std::pair<PlatformDef, long>>& weights) const {  // :174-183
    long sum = 0;
    for (auto& [p, plat] : platforms)
        for (auto& [def, w] : weights)
            if (dims_le(def.dims, plat.def.dims)) sum += w;
    return sum;
}

long assignment_total_weight(const Assignment& a, const EncodingVars& vars,
                             const std::vector<std::pair<PlatformDef, long>>& weights) {  // encoder.rs:670-692
    long sum = 0;
    for (auto& [def, w] : weights)
        for (auto& tile : vars.grid.data) {
            bool any = false;
            for (Dims d : {def.dims, def.dims.flipped()}) {
                auto it = tile.dims_vars.find(d);
                if (it != tile.dims_vars.end() && it->second < (int)a.size() && a[it->second] == 1) any = true;
            }
            if (any) sum += w;
        }
    return sum;
}

}  // namespace tsso
