// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
//
// Flat-array CPU port of the batched stochastic local search (kernel (b)): the same published step rule and
// counter-based RNG as oracle/sls_model.cpp (which is written for obviousness and rescans the whole grid every step),
// restated the way a plain C port would be — cover counts in a byte array, the uncovered set kept incrementally as row
// bitboards, reach sets as index lists, the site list as an array — so that it is a fair CPU yardstick for the same
// unit of work: one SLS flip (a support added or removed = one candidate layout evaluated incrementally, SURVEY.md
// §8(d)).  bench.py times it on all host threads as the `cpu_baseline` / `--impl reference` leg; tests/test_oracle.py
// checks that it reproduces sls_model.cpp's trajectories bit for bit (so it also IS the step rule, not a look-alike).
//
// Semantics anchor as for sls_model.cpp: a site's reach is what PlatformLayout::validate's three ceiling-masked
// 4-neighbour dilations produce from that site alone (src/encoder/platform_layout.rs:127-141).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <memory>
#include <vector>

namespace {

constexpr uint32_t K1 = 0x9E3779B9u, K2 = 0x85EBCA6Bu, K3 = 0xC2B2AE35u;
constexpr int NO_BOUND = 1 << 20;
constexpr uint32_t TABU_BIT = 0x40000000u;

inline uint32_t fmix32(uint32_t h) { h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }
inline uint32_t chain_base(uint64_t seed, uint32_t chain) { return fmix32((uint32_t)seed ^ fmix32((uint32_t)(seed >> 32) + chain * K1 + 0x5bd1e995u)); }
inline uint32_t step_hash(uint32_t base, uint32_t step) { return fmix32(base ^ (step * K1)); }
inline uint32_t noise_q7(int noise_pct) { return (uint32_t)((noise_pct * 128 + 50) / 100); }
inline uint32_t tie_add(uint32_t hs, uint32_t cell) { return (hs * ((2u * cell + 1u) * K2)) >> 16; }
inline uint32_t tie_remove(uint32_t hs, uint32_t i) { return (hs * ((2u * i + 1u) * K3)) >> 16; }
inline int tenure_of(uint32_t global_chain) { static const int t[4] = {3, 6, 12, 20}; return t[global_chain & 3u]; }
inline int effective_tenure(int tenure, int k) { int c = k / 3; c = c < 2 ? 2 : c; return tenure < c ? tenure : c; }
inline uint16_t stamp_reset(uint32_t step) { return (uint16_t)(step - 0x8000u); }
inline bool is_tabu(uint32_t step, uint16_t stamp, int tenure) { return (uint16_t)((uint16_t)step - stamp) < (uint16_t)tenure; }
inline int pick_rotated(uint32_t bits, uint32_t o) {
    uint32_t rot = o ? ((bits >> o) | (bits << (32 - o))) : bits;
    return (int)((__builtin_ctz(rot) + o) & 31u);
}

struct Terrain {
    uint8_t ceil[1024];
    uint16_t reach[1024][25];   // tiles within geodesic distance <= 3 through ceiling, row-major
    uint8_t n_reach[1024];
    uint64_t window[1024];      // the same set as a 7x7 mask around the site: bit (dy+3)*7 + dx+3
    // WINDOW mode (csrc/lns.cu, see sls_model.cpp): tiles that still need cover, and the core new supports must lie in
    uint8_t need[1024];
    int core_lo = 0, core_hi = 32;
    bool win_mode = false;

    void build(const uint8_t* grid, int w, int h) {
        std::memset(ceil, 0, sizeof ceil);
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) ceil[y * 32 + x] = grid[y * w + x] != 0;
        for (int s = 0; s < 1024; s++) {
            n_reach[s] = 0; window[s] = 0;
            if (!ceil[s]) continue;
            const int sx = s & 31, sy = s >> 5;
            uint8_t cur[7][7] = {}, nxt[7][7];      // platform_layout.rs:127-141 inside the site's own window
            cur[3][3] = 1;
            for (int round = 0; round < 3; round++) {
                std::memcpy(nxt, cur, sizeof cur);
                for (int j = 0; j < 7; j++) for (int i = 0; i < 7; i++) {
                    if (!cur[j][i]) continue;
                    const int nx[4] = {i + 1, i, i - 1, i}, ny[4] = {j, j + 1, j, j - 1};
                    for (int d = 0; d < 4; d++) {
                        if (nx[d] < 0 || nx[d] > 6 || ny[d] < 0 || ny[d] > 6) continue;
                        const int gx = sx + nx[d] - 3, gy = sy + ny[d] - 3;
                        if (gx >= 0 && gx < 32 && gy >= 0 && gy < 32 && ceil[gy * 32 + gx]) nxt[ny[d]][nx[d]] = 1;
                    }
                }
                std::memcpy(cur, nxt, sizeof cur);
            }
            for (int j = 0; j < 7; j++) for (int i = 0; i < 7; i++)
                if (cur[j][i]) { reach[s][n_reach[s]++] = (uint16_t)((sy + j - 3) * 32 + sx + i - 3); window[s] |= 1ull << (j * 7 + i); }
        }
        std::memcpy(need, ceil, sizeof need);
    }
};

struct Chain {
    uint8_t S[1024], bestS[1024], cnt[1024];
    uint16_t stamp[1024];
    uint32_t U[32];              // uncovered ceiling tiles, row bitboards
    uint16_t sites[1024];
    int k = 0, best = NO_BOUND, done = 0;
    uint32_t step = 0;
    uint64_t scored = 0, steps_done = 0, flips = 0;
};

struct Runner {
    const Terrain& T;
    Chain& c;
    uint32_t base;
    int tenure, ten = 1;
    Runner(const Terrain& t, Chain& ch, uint32_t b, int tn) : T(t), c(ch), base(b), tenure(tn) {}

    inline void cover(int s, int d) {
        for (int i = 0; i < T.n_reach[s]; i++) {
            const int t = T.reach[s][i];
            c.cnt[t] = (uint8_t)(c.cnt[t] + d);
            if (c.cnt[t] == 0 && T.need[t]) c.U[t >> 5] |= 1u << (t & 31); else c.U[t >> 5] &= ~(1u << (t & 31));
        }
        c.flips++;
    }
    inline int count_eq(int s, uint8_t value) const {   // (tiles the frozen supports cover are neither a gain nor a loss; need = ceil outside WINDOW mode)
        int n = 0;
        for (int i = 0; i < T.n_reach[s]; i++) { const int t = T.reach[s][i]; n += (c.cnt[t] == value) & T.need[t]; }
        return n;
    }

    void remove_min_loss(bool use_tabu, uint32_t hs) {
        uint32_t best_key = 0xffffffffu;
        int best_i = 0;
        for (int i = 0; i < c.k; i++) {
            const int v = c.sites[i];
            const uint32_t young = (use_tabu && is_tabu(c.step, c.stamp[v], ten)) ? TABU_BIT : 0u;
            const uint32_t key = young | ((uint32_t)count_eq(v, 1) << 16) | tie_remove(hs, (uint32_t)i);
            if (key < best_key) { best_key = key; best_i = i; }
        }
        const int u = c.sites[best_i];
        c.sites[best_i] = c.sites[c.k - 1];
        c.k--;
        cover(u, -1);
        c.S[u] = 0;
        c.stamp[u] = (uint16_t)c.step;
    }

    void run(long long steps, int epoch_bound, int target, int noise_pct) {
        if (c.done) return;
        if (target >= 0 && epoch_bound <= target) return;
        c.k = 0;
        for (int t = 0; t < 1024; t++) if (c.S[t]) c.sites[c.k++] = (uint16_t)t;
        std::memset(c.cnt, 0, sizeof c.cnt);
        for (int i = 0; i < c.k; i++) for (int j = 0; j < T.n_reach[c.sites[i]]; j++) c.cnt[T.reach[c.sites[i]][j]]++;
        for (int y = 0; y < 32; y++) { c.U[y] = 0; for (int x = 0; x < 32; x++) if (T.need[y * 32 + x] && !c.cnt[y * 32 + x]) c.U[y] |= 1u << x; }
        const uint16_t reset = stamp_reset(c.step);
        for (int t = 0; t < 1024; t++) c.stamp[t] = reset;
        const uint32_t nq7 = noise_q7(noise_pct);
        long long it = 0;
        for (; it < steps; it++, c.step++) {
            const int limit = std::min(epoch_bound, c.best);
            const uint32_t hs = step_hash(base, c.step);
            ten = effective_tenure(tenure, c.k);
            if (c.k >= limit) {
                if (c.k == 0) { c.done = 1; break; }
                c.scored += (uint64_t)c.k;
                remove_min_loss(false, hs);
                continue;
            }
            uint32_t rowmask = 0;
            for (int y = 0; y < 32; y++) rowmask |= c.U[y] ? 1u << y : 0u;
            if (!rowmask) {
                c.best = c.k;
                std::memcpy(c.bestS, c.S, 1024);
                if (c.k <= target || c.k == 0) { c.done = 1; it++; c.step++; break; }
                continue;
            }
            if (c.k == limit - 1 && c.k > 0) {
                c.scored += (uint64_t)c.k;
                remove_min_loss(true, hs);
                rowmask = 0;
                for (int y = 0; y < 32; y++) rowmask |= c.U[y] ? 1u << y : 0u;
            }
            const int y = pick_rotated(rowmask, hs & 31u), x = pick_rotated(c.U[y], (hs >> 5) & 31u), t = y * 32 + x;
            const bool noise = ((hs >> 10) & 127u) < nq7;
            const uint64_t wt = T.window[t];   // candidates = tiles of R(t), visited in diamond order; `lane` = index in the 25-tile diamond
            uint32_t mx = 0;
            int v = -1, nc = 0, lane = 0;
            for (int dy = -3; dy <= 3; dy++)
                for (int dx = -3; dx <= 3; dx++) {
                    if (std::abs(dx) + std::abs(dy) > 3) continue;
                    const int ln = lane++;
                    if (!((wt >> ((dy + 3) * 7 + dx + 3)) & 1ull)) continue;
                    if (T.win_mode && (x + dx < T.core_lo || x + dx >= T.core_hi || y + dy < T.core_lo || y + dy >= T.core_hi)) continue;
                    const int cv = (y + dy) * 32 + x + dx;
                    nc++;
                    const uint32_t tie = tie_add(hs, (uint32_t)ln);
                    uint32_t key;
                    if (noise) key = 0x10000u | tie;
                    else key = (is_tabu(c.step, c.stamp[cv], ten) ? 0u : TABU_BIT) | ((uint32_t)(count_eq(cv, 0) + 1) << 16) | tie;
                    if (v < 0 || key > mx) { mx = key; v = cv; }
                }
            if (T.win_mode && nc == 0) { c.done = 1; break; }
            if (!noise) c.scored += (uint64_t)nc;
            cover(v, +1);
            c.S[v] = 1;
            c.sites[c.k++] = (uint16_t)v;
            c.stamp[v] = (uint16_t)c.step;
        }
        c.steps_done += (uint64_t)it;
    }
};

}  // namespace

extern "C" {

// Same contract as tsso_sls_model (oracle/sls_model.cpp) plus: chains are distributed over `threads` host threads inside
// every epoch (they only meet at the epoch boundary, where the bound is shared), out_flips = supports added + removed per
// chain (the unit of SURVEY.md §8(d)), *out_seconds = wall time of the epochs alone (terrain tables and chain setup excluded);
// out_epoch_seconds / out_epoch_flips (optional, [n_epochs]): wall time of every epoch and the cumulative flips after it.
int tsso_sls_flat(const uint8_t* grid, int w, int h, int n_chains, uint32_t chain_offset, uint64_t seed, int noise_pct,
                  const long long* epochs, int n_epochs, int share_bound, const uint8_t* init_S, int threads, uint8_t* out_S, uint8_t* out_bestS,
                  int* out_k, int* out_best, uint32_t* out_step, uint64_t* out_scored, uint64_t* out_steps, uint64_t* out_flips, double* out_seconds,
                  double* out_epoch_seconds, uint64_t* out_epoch_flips) {
    if (w > 32 || h > 32 || n_chains <= 0) return -1;
    std::unique_ptr<Terrain> terrain(new Terrain());   // (60 KB of tables: on the heap, one per call, so concurrent calls do not share them)
    Terrain& T = *terrain;
    T.build(grid, w, h);
    std::vector<Chain> chains((size_t)n_chains);
    for (int i = 0; i < n_chains; i++) {
        Chain& c = chains[(size_t)i];
        std::memset(c.S, 0, sizeof c.S); std::memset(c.bestS, 0, sizeof c.bestS);
        if (init_S) {
            std::memcpy(c.S, init_S + (size_t)i * 1024, 1024);
            c.k = (int)std::count(c.S, c.S + 1024, (uint8_t)1);
        }
    }
    if (threads < 1) threads = 1;
    int shared = NO_BOUND;
    double seconds = 0;
    for (int e = 0; e < n_epochs; e++) {
        const long long steps = epochs[3 * e];
        int bound = (int)epochs[3 * e + 1];
        const int target = (int)epochs[3 * e + 2];
        if (share_bound) bound = std::min(bound, shared);
        const auto t0 = std::chrono::steady_clock::now();
        std::atomic<int> next{0};
        auto work = [&]() {
            for (int i; (i = next.fetch_add(1)) < n_chains;)
                Runner(T, chains[(size_t)i], chain_base(seed, chain_offset + (uint32_t)i), tenure_of(chain_offset + (uint32_t)i)).run(steps, bound, target, noise_pct);
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; t++) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
        const double es = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        seconds += es;
        if (out_epoch_seconds) out_epoch_seconds[e] = es;
        if (out_epoch_flips) { uint64_t f = 0; for (auto& c : chains) f += c.flips; out_epoch_flips[e] = f; }   // cumulative
        for (auto& c : chains) shared = std::min(shared, c.best);
    }
    for (int i = 0; i < n_chains; i++) {
        const Chain& c = chains[(size_t)i];
        if (out_S) std::memcpy(out_S + (size_t)i * 1024, c.S, 1024);
        if (out_bestS) std::memcpy(out_bestS + (size_t)i * 1024, c.bestS, 1024);
        if (out_k) out_k[i] = c.k;
        if (out_best) out_best[i] = c.best;
        if (out_step) out_step[i] = c.step;
        if (out_scored) out_scored[i] = c.scored;
        if (out_steps) out_steps[i] = c.steps_done;
        if (out_flips) out_flips[i] = c.flips;
    }
    if (out_seconds) *out_seconds = seconds;
    return 0;
}

// WINDOW mode through the flat port (same contract as tsso_sls_window_model, oracle/sls_model.cpp) plus out_flips per chain:
// the CPU arm of configs[3] (bench.py) and a second implementation the scalar model is compared with.
int tsso_sls_window_flat(const uint8_t* terrain, const uint8_t* need, int core_lo, int core_hi, int n_chains, uint32_t chain_offset, uint64_t seed,
                         int noise_pct, long long steps, const uint8_t* init_S, uint32_t init_step, uint8_t* out_bestS, int* out_best, int* out_k,
                         uint64_t* out_flips) {
    std::unique_ptr<Terrain> terrain_tables(new Terrain());
    Terrain& T = *terrain_tables;
    T.build(terrain, 32, 32);
    std::memcpy(T.need, need, 1024);
    T.core_lo = core_lo; T.core_hi = core_hi; T.win_mode = true;
    std::unique_ptr<Chain> chain(new Chain());
    for (int i = 0; i < n_chains; i++) {
        Chain& c = *chain;
        c = Chain();
        std::memcpy(c.S, init_S, 1024);
        std::memcpy(c.bestS, init_S, 1024);
        c.k = (int)std::count(c.S, c.S + 1024, (uint8_t)1);
        c.best = c.k;
        c.step = init_step;
        Runner(T, c, chain_base(seed, chain_offset + (uint32_t)i), tenure_of(chain_offset + (uint32_t)i)).run(steps, NO_BOUND, 0, noise_pct);
        std::memcpy(out_bestS + (size_t)i * 1024, c.bestS, 1024);
        out_best[i] = c.best; out_k[i] = c.k;
        if (out_flips) out_flips[i] = c.flips;
    }
    return 0;
}

}  // extern "C"
