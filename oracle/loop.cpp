// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Encoding::with_limits (src/encoder.rs:619-667), the CNF lowering of its cardinality / pseudo-boolean
// constraints (rustsat 0.7.2 `SatInstance::into_cnf`, crates/repl/src/main.rs:293 — third-party, source not in
// the tree: PARITY UNPINNED, restated from the published encodings) and the REPL's bound-tightening
// solver_loop (crates/repl/src/main.rs:280-366).
#include "oracle.hpp"

#include <algorithm>
#include <map>

namespace tsso {
namespace {

// Totalizer, upper-bound direction only (Bailleux & Boufkhad, CP'03), outputs truncated at k+1.
// out[i] == "at least i+1 of the leaves below are true".  Clauses: (A_i & B_j) -> O_{i+j}.
std::vector<int> totalizer(SatInstance& inst, const std::vector<int>& lits, size_t lo, size_t hi, size_t cap) {
    if (hi - lo == 1) return {lits[lo]};
    size_t mid = lo + (hi - lo) / 2;
    std::vector<int> a = totalizer(inst, lits, lo, mid, cap), b = totalizer(inst, lits, mid, hi, cap);
    size_t m = std::min(cap, a.size() + b.size());
    std::vector<int> o(m);
    for (auto& v : o) v = inst.new_var();
    for (size_t i = 0; i <= a.size(); i++)
        for (size_t j = 0; j <= b.size(); j++) {
            size_t s = i + j;
            if (s == 0 || s > m) continue;
            Clause c;
            if (i) c.push_back(-a[i - 1]);
            if (j) c.push_back(-b[j - 1]);
            c.push_back(o[s - 1]);
            inst.add(std::move(c), F_CARD);
        }
    return o;
}

void add_card_ub(SatInstance& inst, const std::vector<int>& lits, unsigned long k) {
    if (k >= lits.size()) return;  // trivially satisfied
    if (k == 0) { for (int l : lits) inst.add_unit(-l, F_CARD); return; }
    std::vector<int> o = totalizer(inst, lits, 0, lits.size(), k + 1);
    inst.add_unit(-o[k], F_CARD);  // not (at least k+1)
}

// Generalized totalizer (Joshi, Martins & Manquinho, CP'15): node = map weight-sum -> output var, sums above
// limit collapse into limit+1.
std::map<long, int> gte(SatInstance& inst, const std::vector<std::pair<int, long>>& wl, size_t lo, size_t hi, long cap) {
    if (hi - lo == 1) return {{std::min(wl[lo].second, cap), wl[lo].first}};
    size_t mid = lo + (hi - lo) / 2;
    auto a = gte(inst, wl, lo, mid, cap), b = gte(inst, wl, mid, hi, cap);
    std::map<long, int> o;
    auto out_var = [&](long s) { s = std::min(s, cap); auto it = o.find(s); if (it == o.end()) it = o.emplace(s, inst.new_var()).first; return it->second; };
    for (auto& [wa, va] : a) inst.add({-va, out_var(wa)}, F_PB);
    for (auto& [wb, vb] : b) inst.add({-vb, out_var(wb)}, F_PB);
    for (auto& [wa, va] : a)
        for (auto& [wb, vb] : b) inst.add({-va, -vb, out_var(wa + wb)}, F_PB);
    return o;
}

void add_pb_ub(SatInstance& inst, std::vector<std::pair<int, long>> wl, long limit) {
    // normalise negative weights: w*x = w + |w|*(~x)
    std::vector<std::pair<int, long>> pos;
    for (auto& [l, w] : wl) {
        if (w == 0) continue;
        if (w < 0) { limit += -w; pos.push_back({-l, -w}); } else pos.push_back({l, w});
    }
    if (limit < 0) { inst.add({}, F_PB); return; }
    long total = 0;
    for (auto& p : pos) total += p.second;
    if (total <= limit || pos.empty()) return;
    auto root = gte(inst, pos, 0, pos.size(), limit + 1);
    for (auto& [s, v] : root)
        if (s > limit) inst.add_unit(-v, F_PB);
}

}  // namespace

SatInstance Encoding::with_limits(const PlatformLimits& limits) const {  // encoder.rs:619-667
    SatInstance inst = instance;  // clone (encoder.rs:620)
    std::vector<std::pair<int, long>> weight_pb;
    // card_limits.keys().chain(weights.keys()).unique()
    std::vector<PlatformDef> types;
    auto push_unique = [&](const PlatformDef& d) { if (std::find(types.begin(), types.end(), d) == types.end()) types.push_back(d); };
    for (auto& c : limits.card_limits) push_unique(c.first);
    for (auto& w : limits.weights) push_unique(w.first);

    std::vector<std::pair<std::vector<int>, unsigned long>> cards;
    for (const PlatformDef& type : types) {
        std::vector<int> lits;
        if (type.rectangular()) {  // encoder.rs:629-641: one fresh var per tile implied by both orientations
            for (const EncodingTileVars& tv : vars.grid.data) {
                int limit_var = inst.new_var();
                lits.push_back(limit_var);
                for (Dims d : {type.dims, type.dims.flipped()}) {
                    auto it = tv.dims_vars.find(d);
                    if (it != tv.dims_vars.end()) inst.add_lit_impl_lit(it->second, limit_var, F_LIMIT_LINK);
                }
            }
        } else if (vars.dim_map.count(type.dims)) {  // encoder.rs:643-646: that dims' var at EVERY tile, row-major
            for (const EncodingTileVars& tv : vars.grid.data) lits.push_back(tv.dims_vars.at(type.dims));
        }
        for (auto& c : limits.card_limits)
            if (c.first == type) cards.push_back({lits, c.second});
        if (limits.weight_limit)
            for (auto& w : limits.weights)
                if (w.first == type)
                    for (int l : lits) weight_pb.push_back({l, w.second});
    }
    // into_cnf(): constraints are lowered after all instance variables exist
    for (auto& [lits, k] : cards) add_card_ub(inst, lits, k);
    if (limits.weight_limit) add_pb_ub(inst, weight_pb, *limits.weight_limit);
    return inst;
}

LoopResult solver_loop(const WorldGrid& world, const std::vector<PlatformDef>& defs, PlatformLimits limits,
                       int64_t conflict_budget_per_solve, const volatile int* interrupt) {
    LoopResult out;
    Encoding enc = Encoding::encode(defs, world);  // crates/repl/src/main.rs:254 (once per `solve`)
    const PlatformDef one{{1, 1}};
    while (true) {
        SatInstance inst = enc.with_limits(limits);       // main.rs:292
        Assignment asg;                                   // main.rs:293-295: into_cnf + fresh solver
        LoopStep step{};
        step.bound = -1;
        for (auto& c : limits.card_limits) if (c.first == one) step.bound = (long)c.second;
        step.result = solve_cnf(inst.n_vars, inst.clauses, asg, &step.stats, conflict_budget_per_solve, interrupt);
        if (step.result != 10) {                          // main.rs:331-338 Unsat / Interrupted -> return
            out.proved_optimal = (step.result == 20) && !out.steps.empty();
            out.steps.push_back(step);
            return out;
        }
        PlatformLayout layout = PlatformLayout::from_assignment(asg, enc.vars);  // main.rs:328-329
        step.count = layout.platform_count();
        step.valid = layout.validate(world).is_valid();    // main.rs:353 (warn only)
        out.steps.push_back(step);
        if (layout.platform_count() == 0) { out.best = layout; return out; }     // main.rs:341-344
        // main.rs:346: card_limits[1x1] = count - 1
        bool found = false;
        for (auto& c : limits.card_limits) if (c.first == one) { c.second = layout.platform_count() - 1; found = true; }
        if (!found) limits.card_limits.push_back({one, layout.platform_count() - 1});
        out.best = std::move(layout);
    }
}

}  // namespace tsso
