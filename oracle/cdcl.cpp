// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
//
// CDCL SAT solver standing in for rustsat-glucose 0.7.2 (`rustsat_glucose::simp::Glucose`, C++ Glucose 4,
// a crates.io dependency whose source is NOT under /root/reference: Cargo.lock:2841-2852).  Call sites it
// serves: crates/repl/src/solver_runner.rs:12-16 (add_cnf + solve), crates/repl/src/main.rs:326-329
// (full_solution), crates/gui/src/solver_backend.rs:79-90.
//
// This is a restatement of the PUBLISHED algorithm, not of Glucose's source:
//   * two-watched-literal propagation with blockers, first-UIP learning, basic clause minimisation,
//     VSIDS with a binary heap, phase saving                       (Een & Sorensson, "An Extensible SAT-solver")
//   * LBD ("glue") of learnt clauses, glue<=2 kept forever, clause-DB halving every 2000+300k conflicts
//                                                                 (Audemard & Simon, IJCAI'09)
//   * dynamic restarts: restart when  K * avg(last 50 LBDs) > global average LBD, blocked when the trail is
//     R x longer than its 5000-conflict average                   (Audemard & Simon, CP'12; K=0.8, R=1.4)
// No preprocessing (Glucose's `simp` front-end is not restated); answers (SAT/UNSAT) are solver-independent.
#include "oracle.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>

namespace tsso {
namespace {

using Lit = int;  // 2*var + sign, var 0-based
inline Lit mk(int var, bool neg) { return 2 * var + (neg ? 1 : 0); }
inline Lit neg(Lit l) { return l ^ 1; }
inline int var_of(Lit l) { return l >> 1; }
inline bool sign_of(Lit l) { return l & 1; }

struct Cl {
    std::vector<Lit> lits;
    bool learnt = false;
    bool deleted = false;
    int lbd = 0;
    float act = 0;
};

struct Watcher { int cref; Lit blocker; };

template <int N>
struct BoundedQueue {
    std::vector<uint32_t> buf; size_t head = 0, count = 0; uint64_t sum = 0;
    BoundedQueue() : buf(N, 0) {}
    void push(uint32_t x) {
        if (count == (size_t)N) { sum -= buf[head]; buf[head] = x; head = (head + 1) % N; }
        else { buf[(head + count) % N] = x; count++; }
        sum += x;
    }
    bool full() const { return count == (size_t)N; }
    double avg() const { return count ? (double)sum / (double)count : 0.0; }
    void clear() { head = 0; count = 0; sum = 0; }
};

struct Solver {
    int nv = 0;
    std::vector<Cl> cls;
    std::vector<int> learnts;
    std::vector<std::vector<Watcher>> watches;  // by literal
    std::vector<int8_t> val;                    // by var: 0 undef, 1 true, -1 false
    std::vector<int> level, reason;             // by var
    std::vector<Lit> trail;
    std::vector<int> trail_lim;
    size_t qhead = 0;
    std::vector<double> activity;
    std::vector<int8_t> polarity;  // saved phase: 1 = last assigned false (MiniSat convention: sign)
    std::vector<int> heap, heap_pos;
    double var_inc = 1.0, var_decay = 0.8;
    float cla_inc = 1.0f;
    bool ok = true;
    std::vector<uint8_t> seen;
    std::vector<uint32_t> lbd_stamp; uint32_t lbd_counter = 0;
    SolveStats st;

    explicit Solver(int n) : nv(n), watches(2 * n), val(n, 0), level(n, 0), reason(n, -1), activity(n, 0.0),
                             polarity(n, 1), heap_pos(n, -1), seen(n, 0), lbd_stamp(n + 1, 0) {
        for (int v = 0; v < n; v++) heap_insert(v);
    }

    int8_t value(Lit l) const { int8_t v = val[var_of(l)]; return sign_of(l) ? -v : v; }
    int decision_level() const { return (int)trail_lim.size(); }

    // ---- VSIDS heap (max-heap on activity)
    bool heap_lt(int a, int b) const { return activity[a] > activity[b]; }
    void heap_up(int i) {
        int x = heap[i];
        while (i > 0) { int p = (i - 1) >> 1; if (!heap_lt(x, heap[p])) break; heap[i] = heap[p]; heap_pos[heap[i]] = i; i = p; }
        heap[i] = x; heap_pos[x] = i;
    }
    void heap_down(int i) {
        int x = heap[i]; int n = (int)heap.size();
        while (true) {
            int l = 2 * i + 1, r = l + 1; if (l >= n) break;
            int c = (r < n && heap_lt(heap[r], heap[l])) ? r : l;
            if (!heap_lt(heap[c], x)) break;
            heap[i] = heap[c]; heap_pos[heap[i]] = i; i = c;
        }
        heap[i] = x; heap_pos[x] = i;
    }
    void heap_insert(int v) { if (heap_pos[v] >= 0) return; heap_pos[v] = (int)heap.size(); heap.push_back(v); heap_up(heap_pos[v]); }
    int heap_pop() {
        int x = heap[0]; heap_pos[x] = -1; int last = heap.back(); heap.pop_back();
        if (!heap.empty()) { heap[0] = last; heap_pos[last] = 0; heap_down(0); }
        return x;
    }
    void bump_var(int v) {
        if ((activity[v] += var_inc) > 1e100) { for (auto& a : activity) a *= 1e-100; var_inc *= 1e-100; }
        if (heap_pos[v] >= 0) heap_up(heap_pos[v]);
    }
    void bump_clause(Cl& c) {
        if ((c.act += cla_inc) > 1e20f) { for (int i : learnts) cls[i].act *= 1e-20f; cla_inc *= 1e-20f; }
    }

    void enqueue(Lit l, int from) {
        int v = var_of(l);
        val[v] = sign_of(l) ? -1 : 1; level[v] = decision_level(); reason[v] = from; trail.push_back(l);
    }

    void attach(int cref) {
        Cl& c = cls[cref];
        watches[neg(c.lits[0])].push_back({cref, c.lits[1]});
        watches[neg(c.lits[1])].push_back({cref, c.lits[0]});
    }

    bool add_clause(std::vector<Lit> ps) {
        if (!ok) return false;
        std::sort(ps.begin(), ps.end());
        std::vector<Lit> out; Lit prev = -1;
        for (Lit l : ps) {
            if (value(l) == 1 || l == neg(prev)) return true;  // satisfied / tautology
            if (value(l) != -1 && l != prev) { out.push_back(l); prev = l; }
        }
        if (out.empty()) return ok = false;
        if (out.size() == 1) { enqueue(out[0], -1); return ok = (propagate() == -1); }
        cls.push_back(Cl{std::move(out)});
        attach((int)cls.size() - 1);
        return true;
    }

    int propagate() {
        int confl = -1;
        while (qhead < trail.size()) {
            Lit p = trail[qhead++];  // p is true; visit clauses watching ~p ... stored under watches[p]
            st.propagations++;
            auto& ws = watches[p];
            size_t i = 0, j = 0, n = ws.size();
            while (i < n) {
                Watcher w = ws[i];
                if (value(w.blocker) == 1) { ws[j++] = ws[i++]; continue; }
                Cl& c = cls[w.cref];
                if (c.deleted) { i++; continue; }
                Lit false_lit = neg(p);
                if (c.lits[0] == false_lit) std::swap(c.lits[0], c.lits[1]);
                i++;
                Lit first = c.lits[0];
                Watcher nw{w.cref, first};
                if (first != w.blocker && value(first) == 1) { ws[j++] = nw; continue; }
                bool found = false;
                for (size_t k = 2; k < c.lits.size(); k++)
                    if (value(c.lits[k]) != -1) {
                        c.lits[1] = c.lits[k]; c.lits[k] = false_lit;
                        watches[neg(c.lits[1])].push_back(nw);
                        found = true; break;
                    }
                if (found) continue;
                ws[j++] = nw;
                if (value(first) == -1) {
                    confl = w.cref; qhead = trail.size();
                    while (i < n) ws[j++] = ws[i++];
                } else {
                    enqueue(first, w.cref);
                }
            }
            ws.resize(j);
            if (confl != -1) break;
        }
        return confl;
    }

    int compute_lbd(const std::vector<Lit>& lits) {
        lbd_counter++;
        int n = 0;
        for (Lit l : lits) { int lv = level[var_of(l)]; if (lbd_stamp[lv] != lbd_counter) { lbd_stamp[lv] = lbd_counter; n++; } }
        return n;
    }

    void analyze(int confl, std::vector<Lit>& out, int& bt_level) {
        int path = 0; Lit p = -1; out.clear(); out.push_back(0);
        int idx = (int)trail.size() - 1;
        std::vector<int> to_clear;
        do {
            Cl& c = cls[confl];
            if (c.learnt) bump_clause(c);
            for (size_t k = (p == -1 ? 0 : 1); k < c.lits.size(); k++) {
                Lit q = c.lits[k]; int v = var_of(q);
                if (!seen[v] && level[v] > 0) {
                    bump_var(v); seen[v] = 1; to_clear.push_back(v);
                    if (level[v] >= decision_level()) path++; else out.push_back(q);
                }
            }
            while (!seen[var_of(trail[idx--])]) {}
            p = trail[idx + 1]; confl = reason[var_of(p)]; seen[var_of(p)] = 0; path--;
            // invariant: a reason clause holds its propagated literal at position 0 (propagate / learnt enqueue)
        } while (path > 0);
        out[0] = neg(p);
        // basic minimisation: drop q if every other literal of reason(q) is already in the clause (or level 0)
        size_t j = 1;
        for (size_t i = 1; i < out.size(); i++) {
            int v = var_of(out[i]); int r = reason[v];
            bool keep = (r == -1);
            if (!keep) {
                const Cl& c = cls[r];
                for (Lit q : c.lits) { int u = var_of(q); if (u != v && !seen[u] && level[u] > 0) { keep = true; break; } }
            }
            if (keep) out[j++] = out[i];
        }
        out.resize(j);
        if (out.size() == 1) bt_level = 0;
        else {
            size_t mx = 1;
            for (size_t i = 2; i < out.size(); i++) if (level[var_of(out[i])] > level[var_of(out[mx])]) mx = i;
            std::swap(out[1], out[mx]);
            bt_level = level[var_of(out[1])];
        }
        for (int v : to_clear) seen[v] = 0;
    }

    void cancel_until(int lvl) {
        if (decision_level() <= lvl) return;
        for (int c = (int)trail.size() - 1; c >= trail_lim[lvl]; c--) {
            int v = var_of(trail[c]);
            polarity[v] = sign_of(trail[c]); val[v] = 0; reason[v] = -1; heap_insert(v);
        }
        qhead = trail_lim[lvl];
        trail.resize(trail_lim[lvl]);
        trail_lim.resize(lvl);
    }

    bool locked(int cref) const { const Cl& c = cls[cref]; int v = var_of(c.lits[0]); return val[v] != 0 && reason[v] == cref && value(c.lits[0]) == 1; }

    void reduce_db() {
        std::sort(learnts.begin(), learnts.end(), [&](int a, int b) {
            const Cl &x = cls[a], &y = cls[b];
            if (x.lbd != y.lbd) return x.lbd > y.lbd;   // worst (high glue) first
            return x.act < y.act;
        });
        size_t limit = learnts.size() / 2, j = 0;
        for (size_t i = 0; i < learnts.size(); i++) {
            Cl& c = cls[learnts[i]];
            if (i < limit && c.lbd > 2 && c.lits.size() > 2 && !locked(learnts[i])) { c.deleted = true; std::vector<Lit>().swap(c.lits); }
            else learnts[j++] = learnts[i];
        }
        learnts.resize(j);
        // purge watchers of deleted clauses
        for (auto& ws : watches) {
            size_t k = 0;
            for (auto& w : ws) if (!cls[w.cref].deleted) ws[k++] = w;
            ws.resize(k);
        }
    }

    // returns 10 SAT, 20 UNSAT, 0 unknown
    int solve(int64_t conflict_budget, const volatile int* interrupt) {
        if (!ok) return 20;
        if (propagate() != -1) { ok = false; return 20; }
        BoundedQueue<50> lbd_q; BoundedQueue<5000> trail_q;
        double sum_lbd = 0; const double K = 0.8, R = 1.4;
        uint64_t next_reduce = 2000, reduce_inc = 300, n_reduce = 0;
        std::vector<Lit> learnt;
        while (true) {
            int confl = propagate();
            if (confl != -1) {
                st.conflicts++;
                if (decision_level() == 0) { ok = false; return 20; }
                trail_q.push((uint32_t)trail.size());
                if (st.conflicts > 10000 && lbd_q.full() && (double)trail.size() > R * trail_q.avg()) lbd_q.clear();  // block restart
                int bt;
                analyze(confl, learnt, bt);
                int lbd = compute_lbd(learnt);
                lbd_q.push((uint32_t)lbd); sum_lbd += lbd;
                cancel_until(bt);
                if (learnt.size() == 1) enqueue(learnt[0], -1);
                else {
                    cls.push_back(Cl{learnt, true, false, lbd, 0.0f});
                    int cref = (int)cls.size() - 1;
                    learnts.push_back(cref); attach(cref); bump_clause(cls[cref]);
                    enqueue(learnt[0], cref);
                    st.learnts++;
                }
                var_inc *= 1.0 / var_decay;
                cla_inc *= 1.0f / 0.999f;
                if (st.conflicts % 5000 == 0 && var_decay < 0.95) var_decay += 0.01;
                if (conflict_budget >= 0 && (int64_t)st.conflicts >= conflict_budget) { cancel_until(0); return 0; }
                if ((st.conflicts & 255) == 0 && interrupt && *interrupt) { cancel_until(0); return 0; }
            } else {
                if (lbd_q.full() && lbd_q.avg() * K > sum_lbd / (double)st.conflicts) {
                    lbd_q.clear(); st.restarts++; cancel_until(0);
                    continue;
                }
                if (st.conflicts >= next_reduce) { n_reduce++; next_reduce = st.conflicts + 2000 + reduce_inc * n_reduce; reduce_db(); }
                int next = -1;
                while (!heap.empty()) { int v = heap_pop(); if (val[v] == 0) { next = v; break; } }
                if (next == -1) return 10;
                st.decisions++;
                trail_lim.push_back((int)trail.size());
                enqueue(mk(next, polarity[next]), -1);
            }
        }
    }
};

}  // namespace

int solve_cnf(int n_vars, const std::vector<Clause>& clauses, Assignment& out, SolveStats* stats,
              int64_t conflict_budget, const volatile int* interrupt) {
    auto t0 = std::chrono::steady_clock::now();
    Solver s(n_vars);
    for (const Clause& c : clauses) {
        std::vector<Lit> ps;
        ps.reserve(c.size());
        for (int l : c) ps.push_back(mk(std::abs(l) - 1, l < 0));
        if (!s.add_clause(std::move(ps))) break;
    }
    int res = s.solve(conflict_budget, interrupt);
    if (res == 10) {
        out.assign(n_vars + 1, 2);
        for (int v = 0; v < n_vars; v++) out[v + 1] = s.val[v] == 1 ? 1 : (s.val[v] == -1 ? 0 : 2);
    }
    s.st.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats) *stats = s.st;
    return res;
}

}  // namespace tsso
