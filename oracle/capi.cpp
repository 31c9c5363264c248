// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  Flat C entry points for ctypes (oracle/oracle.py).
#include "oracle.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

using namespace tsso;

namespace {
struct EncHandle { Encoding enc; WorldGrid world; std::vector<PlatformDef> defs; };
struct CnfHandle { SatInstance inst; };

WorldGrid make_world(const uint8_t* grid, int w, int h) {
    WorldGrid g;
    g.dims = Dims{(unsigned long)w, (unsigned long)h};
    g.data.resize((size_t)w * h);
    for (size_t i = 0; i < g.data.size(); i++) g.data[i] = grid[i] != 0;
    return g;
}
std::vector<PlatformDef> make_defs(const int* wh, int n) {
    std::vector<PlatformDef> d;
    for (int i = 0; i < n; i++) d.push_back(PlatformDef{Dims{(unsigned long)wh[2 * i], (unsigned long)wh[2 * i + 1]}});
    return d;
}
// platform record: x, y, def_w, def_h, rotated
PlatformLayout make_layout(const int* plats, int n) {
    PlatformLayout l;
    for (int i = 0; i < n; i++) {
        const int* p = plats + 5 * i;
        Platform pl{Point{p[0], p[1]}, PlatformDef{Dims{(unsigned long)p[2], (unsigned long)p[3]}}, p[4] != 0};
        l.platforms[pl.point] = pl;
    }
    return l;
}
int write_layout(const PlatformLayout& l, int* out, int cap) {
    int n = 0;
    for (auto& [pt, pl] : l.platforms) {
        if (n < cap) {
            int* p = out + 5 * n;
            p[0] = (int)pl.point.x; p[1] = (int)pl.point.y; p[2] = (int)pl.def.dims.width; p[3] = (int)pl.def.dims.height; p[4] = pl.rotated;
        }
        n++;
    }
    return n;
}
PlatformLimits make_limits(const int* card, int n_card, const int* weights, int n_weights, int has_weight_limit, long weight_limit) {
    PlatformLimits lim;
    for (int i = 0; i < n_card; i++)
        lim.card_limits.push_back({PlatformDef{Dims{(unsigned long)card[3 * i], (unsigned long)card[3 * i + 1]}}, (unsigned long)card[3 * i + 2]});
    for (int i = 0; i < n_weights; i++)
        lim.weights.push_back({PlatformDef{Dims{(unsigned long)weights[3 * i], (unsigned long)weights[3 * i + 1]}}, (long)weights[3 * i + 2]});
    if (has_weight_limit) lim.weight_limit = weight_limit;
    return lim;
}
}  // namespace

extern "C" {

// ---- math / platform unit surface
int tsso_dims_partial_cmp(int aw, int ah, int bw, int bh) {  // -1 less, 0 equal, 1 greater, 2 none
    switch (partial_cmp(Dims{(unsigned long)aw, (unsigned long)ah}, Dims{(unsigned long)bw, (unsigned long)bh})) {
        case POrd::Less: return -1; case POrd::Equal: return 0; case POrd::Greater: return 1; default: return 2;
    }
}
int tsso_iter_within(int w, int h, int* out_xy, int cap) {
    auto pts = iter_within(Dims{(unsigned long)w, (unsigned long)h});
    for (size_t i = 0; i < pts.size() && (int)i < cap; i++) { out_xy[2 * i] = (int)pts[i].x; out_xy[2 * i + 1] = (int)pts[i].y; }
    return (int)pts.size();
}
int tsso_iter_manhattan(int cx, int cy, int dist, int* out_xy, int cap) {
    auto pts = iter_within_manhattan(Point{cx, cy}, (unsigned)dist);
    for (size_t i = 0; i < pts.size() && (int)i < cap; i++) { out_xy[2 * i] = (int)pts[i].x; out_xy[2 * i + 1] = (int)pts[i].y; }
    return (int)pts.size();
}
void tsso_neighbors(int x, int y, int* out_xy) {
    Point ns[4]; neighbors(Point{x, y}, ns);
    for (int i = 0; i < 4; i++) { out_xy[2 * i] = (int)ns[i].x; out_xy[2 * i + 1] = (int)ns[i].y; }
}
int tsso_platform_overlaps(const int* a, const int* b) {
    Platform pa{Point{a[0], a[1]}, PlatformDef{Dims{(unsigned long)a[2], (unsigned long)a[3]}}, a[4] != 0};
    Platform pb{Point{b[0], b[1]}, PlatformDef{Dims{(unsigned long)b[2], (unsigned long)b[3]}}, b[4] != 0};
    return pa.overlaps(pb);
}

// ---- world
// returns 0 ok, 1 error (message in err); *ragged set when rows had unequal lengths
int tsso_parse_world(const char* text, uint8_t* out, int cap, int* w, int* h, int* ragged, char* err, int err_cap) {
    WorldGrid g; bool rg = false;
    std::string e = parse_world_toml(text, g, &rg);
    if (!e.empty()) { if (err && err_cap > 0) { std::strncpy(err, e.c_str(), err_cap - 1); err[err_cap - 1] = 0; } return 1; }
    *w = (int)g.dims.width; *h = (int)g.dims.height; if (ragged) *ragged = rg;
    if ((int)g.data.size() > cap) return 2;
    for (size_t i = 0; i < g.data.size(); i++) out[i] = g.data[i];
    return 0;
}
int tsso_world_to_toml(const uint8_t* grid, int w, int h, char* out, int cap) {
    std::string s = world_to_toml(make_world(grid, w, h));
    if ((int)s.size() + 1 > cap) return -(int)s.size() - 1;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

// ---- DAG (for the doc-comment diagram check, encoder.rs:45-51)
int tsso_dag_platform_edges(const int* defs_wh, int n_defs, int* out, int cap) {  // records: sw, sh, lw, lh
    EncodingVars v; SatInstance tmp;
    std::vector<Dims> keys;
    for (auto& d : make_defs(defs_wh, n_defs))
        for (Dims k : {d.dims, d.dims.flipped()}) if (std::find(keys.begin(), keys.end(), k) == keys.end()) keys.push_back(k);
    EncodingDag dag(keys);
    auto e = dag.platform_edges_reduced();
    for (size_t i = 0; i < e.size() && (int)i < cap; i++) {
        out[4 * i] = (int)e[i].first.width; out[4 * i + 1] = (int)e[i].first.height;
        out[4 * i + 2] = (int)e[i].second.width; out[4 * i + 3] = (int)e[i].second.height;
    }
    return (int)e.size();
}
int tsso_dag_point_edges(const int* defs_wh, int n_defs, int* out, int cap) {  // records: px, py, w, h
    std::vector<Dims> keys;
    for (auto& d : make_defs(defs_wh, n_defs))
        for (Dims k : {d.dims, d.dims.flipped()}) if (std::find(keys.begin(), keys.end(), k) == keys.end()) keys.push_back(k);
    EncodingDag dag(keys);
    auto e = dag.point_platform_edges_reduced();
    for (size_t i = 0; i < e.size() && (int)i < cap; i++) {
        out[4 * i] = (int)e[i].first.x; out[4 * i + 1] = (int)e[i].first.y;
        out[4 * i + 2] = (int)e[i].second.width; out[4 * i + 3] = (int)e[i].second.height;
    }
    return (int)e.size();
}

// ---- encoder
void* tsso_encode(const uint8_t* grid, int w, int h, const int* defs_wh, int n_defs) {
    auto* e = new EncHandle();
    e->world = make_world(grid, w, h);
    e->defs = make_defs(defs_wh, n_defs);
    e->enc = Encoding::encode(e->defs, e->world);
    return e;
}
void tsso_encoding_free(void* h) { delete (EncHandle*)h; }
void* tsso_encoding_cnf(void* h) { auto* c = new CnfHandle(); c->inst = ((EncHandle*)h)->enc.instance; return c; }
int tsso_encoding_num_dims(void* h) { return (int)((EncHandle*)h)->enc.vars.dim_keys.size(); }
void tsso_encoding_dims(void* h, int* out_wh) {
    auto& k = ((EncHandle*)h)->enc.vars.dim_keys;
    for (size_t i = 0; i < k.size(); i++) { out_wh[2 * i] = (int)k[i].width; out_wh[2 * i + 1] = (int)k[i].height; }
}
// plat_var[tile * K + k] (1-based var), terr_var[tile * 4 + layer] (0 = absent)
void tsso_encoding_var_maps(void* h, int* plat_var, int* terr_var) {
    auto& v = ((EncHandle*)h)->enc.vars;
    size_t K = v.dim_keys.size();
    for (size_t t = 0; t < v.grid.data.size(); t++) {
        for (size_t k = 0; k < K; k++) plat_var[t * K + k] = v.grid.data[t].dims_vars.at(v.dim_keys[k]);
        for (int l = 0; l < TERRAIN_SUPPORT_DISTANCE; l++) terr_var[t * 4 + l] = v.grid.data[t].terrain ? (*v.grid.data[t].terrain)[l] : 0;
    }
}
void* tsso_with_limits(void* h, const int* card, int n_card, const int* weights, int n_weights, int has_wl, long wl) {
    auto* c = new CnfHandle();
    c->inst = ((EncHandle*)h)->enc.with_limits(make_limits(card, n_card, weights, n_weights, has_wl, wl));
    return c;
}

// ---- CNF handle
void tsso_cnf_free(void* c) { delete (CnfHandle*)c; }
int tsso_cnf_num_vars(void* c) { return ((CnfHandle*)c)->inst.n_vars; }
int tsso_cnf_num_clauses(void* c) { return (int)((CnfHandle*)c)->inst.clauses.size(); }
long tsso_cnf_num_lits(void* c) { long n = 0; for (auto& cl : ((CnfHandle*)c)->inst.clauses) n += (long)cl.size(); return n; }
void tsso_cnf_get(void* c, int* lits, unsigned* offsets, uint8_t* family) {
    auto& inst = ((CnfHandle*)c)->inst;
    unsigned o = 0;
    for (size_t i = 0; i < inst.clauses.size(); i++) {
        offsets[i] = o;
        for (int l : inst.clauses[i]) lits[o++] = l;
        if (family) family[i] = inst.family[i];
    }
    offsets[inst.clauses.size()] = o;
}
// stats: conflicts, decisions, propagations, restarts, learnts ; returns 10/20/0
int tsso_cnf_solve(void* c, uint8_t* assignment /* n_vars+1 */, long conflict_budget, const volatile int* interrupt,
                   unsigned long long* stats5, double* seconds) {
    auto& inst = ((CnfHandle*)c)->inst;
    Assignment a; SolveStats st;
    int r = solve_cnf(inst.n_vars, inst.clauses, a, &st, conflict_budget, interrupt);
    if (r == 10 && assignment) std::memcpy(assignment, a.data(), a.size());
    if (stats5) { stats5[0] = st.conflicts; stats5[1] = st.decisions; stats5[2] = st.propagations; stats5[3] = st.restarts; stats5[4] = st.learnts; }
    if (seconds) *seconds = st.seconds;
    return r;
}
// solve a caller-provided CSR CNF (lets tests hand the PRODUCT's CNF to the oracle's CDCL as the Glucose stand-in)
int tsso_solve_csr(const int* lits, const unsigned* offsets, int n_clauses, int n_vars, uint8_t* assignment, long conflict_budget) {
    std::vector<Clause> cls((size_t)n_clauses);
    for (int i = 0; i < n_clauses; i++) cls[i].assign(lits + offsets[i], lits + offsets[i + 1]);
    Assignment a;
    int r = solve_cnf(n_vars, cls, a, nullptr, conflict_budget, nullptr);
    if (r == 10 && assignment) std::memcpy(assignment, a.data(), a.size());
    return r;
}
// plain CPU clause check of a full assignment (1/0/2): returns number of falsified clauses
int tsso_cnf_count_falsified(void* c, const uint8_t* assignment, int* first) {
    auto& inst = ((CnfHandle*)c)->inst;
    int n = 0; if (first) *first = -1;
    for (size_t i = 0; i < inst.clauses.size(); i++) {
        bool sat = false;
        for (int l : inst.clauses[i]) { uint8_t v = assignment[std::abs(l)]; if ((l > 0 && v == 1) || (l < 0 && v == 0)) { sat = true; break; } }
        if (!sat) { if (n == 0 && first) *first = (int)i; n++; }
    }
    return n;
}

// Unit propagation to the fixpoint, the specification kernel (c)'s tss_cnf_propagate is compared with (tss.h).  What a SAT
// solver does with the encoder's clauses once the platform variables are decided (src/encoder.rs:500-544: T-layer
// implications; crates/repl/src/solver_runner.rs:12-16 hands exactly these clauses to the solver).  Synchronous rounds:
// every clause is looked at against the state at the START of the round; a clause with no true literal and exactly one
// unassigned literal assigns it.  Two clauses may force opposite values of one variable in the same round: the variable
// keeps both marks, counts as assigned, and reads as True ("True wins"), so the clause that wanted False is reported
// as the conflict.  rounds = rounds executed including the last one that changed nothing.  conflict = lowest index of a
// clause whose literals are all false at the fixpoint, or -1.  assignment: u8[n_vars+1] 0 F / 1 T / 2 unassigned, in/out.
static int propagate_clauses(const std::vector<Clause>& clauses, int nv, uint8_t* assignment, int* conflict, int* rounds) {
    std::vector<uint8_t> pos((size_t)nv + 1, 0), neg((size_t)nv + 1, 0);
    for (int v = 1; v <= nv; v++) { pos[v] = assignment[v] == 1; neg[v] = assignment[v] == 0; }
    int r = 0;
    for (bool changed = true; changed;) {
        changed = false;
        r++;
        std::vector<uint8_t> npos = pos, nneg = neg;
        for (auto& cl : clauses) {
            bool sat = false;
            int n_open = 0, open_lit = 0;
            for (int l : cl) {
                const int v = std::abs(l);
                if (l > 0 ? pos[v] : neg[v]) sat = true;
                if (!pos[v] && !neg[v]) { n_open++; open_lit = l; }
            }
            if (sat || n_open != 1) continue;
            const int v = std::abs(open_lit);
            if (open_lit > 0) npos[v] = 1; else nneg[v] = 1;
            changed = true;
        }
        pos.swap(npos); neg.swap(nneg);
    }
    for (int v = 1; v <= nv; v++) assignment[v] = pos[v] ? 1 : (neg[v] ? 0 : 2);
    int first = -1;
    for (size_t i = 0; i < clauses.size() && first < 0; i++) {
        bool all_false = true;
        for (int l : clauses[i]) { const uint8_t a = assignment[std::abs(l)]; if (a == 2 || (l > 0 ? a == 1 : a == 0)) { all_false = false; break; } }
        if (all_false) first = (int)i;
    }
    if (conflict) *conflict = first;
    if (rounds) *rounds = r;
    return 0;
}
int tsso_cnf_propagate(void* c, uint8_t* assignment, int* conflict, int* rounds) {
    auto& inst = ((CnfHandle*)c)->inst;
    return propagate_clauses(inst.clauses, inst.n_vars, assignment, conflict, rounds);
}
// the same on a caller-provided CSR CNF (random clause sets in tests/test_gpu.py: kernel (c) on inputs no encoder produces)
int tsso_propagate_csr(const int* lits, const unsigned* offsets, int n_clauses, int n_vars, uint8_t* assignment, int* conflict, int* rounds) {
    std::vector<Clause> cls((size_t)n_clauses);
    for (int i = 0; i < n_clauses; i++) cls[i].assign(lits + offsets[i], lits + offsets[i + 1]);
    return propagate_clauses(cls, n_vars, assignment, conflict, rounds);
}

// ---- layout
int tsso_layout_from_assignment(void* h, const uint8_t* assignment, int n, int* out_plats, int cap) {
    Assignment a(assignment, assignment + n);
    return write_layout(PlatformLayout::from_assignment(a, ((EncHandle*)h)->enc.vars), out_plats, cap);
}
// out_unsupported: u8[w*h] mask; out_flags: u8[n] bit0 = overlapping, bit1 = out of bounds.  returns #unsupported
int tsso_validate(const uint8_t* grid, int w, int h, const int* plats, int n, uint8_t* out_unsupported, uint8_t* out_flags) {
    WorldGrid world = make_world(grid, w, h);
    ValidationResult r = make_layout(plats, n).validate(world);
    if (out_unsupported) { std::memset(out_unsupported, 0, (size_t)w * h); for (auto& p : r.unsupported_terrain) out_unsupported[p.x + p.y * w] = 1; }
    if (out_flags)
        for (int i = 0; i < n; i++) {
            const int* p = plats + 5 * i;
            Platform pl{Point{p[0], p[1]}, PlatformDef{Dims{(unsigned long)p[2], (unsigned long)p[3]}}, p[4] != 0};
            out_flags[i] = (r.overlapping_platforms.count(pl) ? 1 : 0) | (r.out_of_bounds_platforms.count(pl) ? 2 : 0);
        }
    return (int)r.unsupported_terrain.size();
}
int tsso_trivial_optimization(const uint8_t* grid, int w, int h, const int* plats, int n, int* out_plats, int cap) {
    WorldGrid world = make_world(grid, w, h);
    PlatformLayout l = make_layout(plats, n);
    l.run_trivial_optimization(world);
    return write_layout(l, out_plats, cap);
}
long tsso_total_weight(const int* plats, int n, const int* weights, int n_weights) {
    return make_layout(plats, n).total_weight(make_limits(nullptr, 0, weights, n_weights, 0, 0).weights);
}
long tsso_assignment_total_weight(void* h, const uint8_t* assignment, int n, const int* weights, int n_weights) {
    Assignment a(assignment, assignment + n);
    return assignment_total_weight(a, ((EncHandle*)h)->enc.vars, make_limits(nullptr, 0, weights, n_weights, 0, 0).weights);
}

// Batched `validate` over 1x1-only layouts given as u8 site masks [n][w*h]; the CPU "layouts evaluated / s"
// baseline beside kernel (a).  out_uncovered / out_count are int32[n].  Returns seconds of wall time.
double tsso_validate_sites_batch(const uint8_t* grid, int w, int h, const uint8_t* sites, long n, int threads,
                                 int* out_uncovered, int* out_count) {
    WorldGrid world = make_world(grid, w, h);
    auto t0 = std::chrono::steady_clock::now();
    std::atomic<long> next{0};
    auto work = [&]() {
        const long chunk = 64;
        while (true) {
            long b = next.fetch_add(chunk);
            if (b >= n) break;
            for (long i = b; i < std::min(n, b + chunk); i++) {
                PlatformLayout l;
                const uint8_t* s = sites + (size_t)i * w * h;
                for (int t = 0; t < w * h; t++)
                    if (s[t]) { Point p{t % w, t / w}; l.platforms[p] = Platform{p, PlatformDef{Dims{1, 1}}, false}; }
                ValidationResult r = l.validate(world);
                out_uncovered[i] = (int)r.unsupported_terrain.size();
                out_count[i] = (int)l.platform_count();
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Same computation as tsso_validate_sites_batch, restated the way a plain C port of platform_layout.rs:85-149 would be
// written (flat byte grids instead of HashSet<Point>/HashMap): the stronger, fairer CPU baseline for kernel (a).
// A test asserts it agrees with the structure-faithful version above.
double tsso_validate_sites_batch_flat(const uint8_t* grid, int w, int h, const uint8_t* sites, long n, int threads,
                                      int* out_uncovered, int* out_count) {
    auto t0 = std::chrono::steady_clock::now();
    std::atomic<long> next{0};
    auto work = [&]() {
        const long chunk = 256;
        std::vector<uint8_t> sup((size_t)w * h), nxt((size_t)w * h);
        while (true) {
            long b = next.fetch_add(chunk);
            if (b >= n) break;
            for (long i = b; i < std::min(n, b + chunk); i++) {
                const uint8_t* s = sites + (size_t)i * w * h;
                int count = 0;
                for (int t = 0; t < w * h; t++) { count += s[t] != 0; sup[t] = s[t] && grid[t]; }  // only terrain can be supported
                for (int round = 0; round < TERRAIN_SUPPORT_DISTANCE - 1; round++) {
                    nxt = sup;
                    for (int y = 0; y < h; y++)
                        for (int x = 0; x < w; x++) {
                            if (!sup[y * w + x]) continue;
                            if (x + 1 < w && grid[y * w + x + 1]) nxt[y * w + x + 1] = 1;
                            if (y + 1 < h && grid[(y + 1) * w + x]) nxt[(y + 1) * w + x] = 1;
                            if (x > 0 && grid[y * w + x - 1]) nxt[y * w + x - 1] = 1;
                            if (y > 0 && grid[(y - 1) * w + x]) nxt[(y - 1) * w + x] = 1;
                        }
                    sup.swap(nxt);
                }
                int unc = 0;
                for (int t = 0; t < w * h; t++) unc += grid[t] && !sup[t];
                out_uncovered[i] = unc;
                out_count[i] = count;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- solver_loop (crates/repl/src/main.rs:280-366).  steps: records of 4 longs (bound, result, count, valid) +
// stats (conflicts) ; returns number of steps.  best layout written to out_plats.
int tsso_solver_loop(const uint8_t* grid, int w, int h, const int* defs_wh, int n_defs, long initial_limit_1x1,
                     long conflict_budget, const volatile int* interrupt, long* steps, int steps_cap,
                     double* step_seconds, int* out_plats, int plats_cap, int* n_plats, int* proved_optimal) {
    WorldGrid world = make_world(grid, w, h);
    PlatformLimits lim;
    if (initial_limit_1x1 >= 0) lim.card_limits.push_back({PlatformDef{Dims{1, 1}}, (unsigned long)initial_limit_1x1});
    LoopResult r = solver_loop(world, make_defs(defs_wh, n_defs), lim, conflict_budget, interrupt);
    for (size_t i = 0; i < r.steps.size() && (int)i < steps_cap; i++) {
        steps[5 * i] = r.steps[i].bound; steps[5 * i + 1] = r.steps[i].result; steps[5 * i + 2] = (long)r.steps[i].count;
        steps[5 * i + 3] = r.steps[i].valid; steps[5 * i + 4] = (long)r.steps[i].stats.conflicts;
        if (step_seconds) step_seconds[i] = r.steps[i].stats.seconds;
    }
    if (n_plats) *n_plats = write_layout(r.best, out_plats, plats_cap);
    if (proved_optimal) *proved_optimal = r.proved_optimal;
    return (int)r.steps.size();
}

}  // extern "C"
