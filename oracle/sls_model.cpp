// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
//
// Scalar CPU model of the batched stochastic local search (kernel (b), timberborn_support_solver_b200/csrc/sls.cu).
// The SLS has no counterpart in the reference ("the three CUDA kernels ... have no counterpart"); its
// SEMANTICS are anchored on the reference's range rule: a site's reach is what PlatformLayout::validate's three
// ceiling-masked 4-neighbour dilations produce from that site alone (src/encoder/platform_layout.rs:127-141), and a
// layout is complete iff validate() reports no unsupported terrain.  This model is written tile-by-tile with byte
// counters (no bit tricks) and replays the kernel's published step rule and counter-based RNG so the GPU
// trajectories can be compared bit for bit (tests/test_sls.py).  It deliberately re-declares the constants of
// sls_spec.hpp: the oracle never includes product headers; a test checks they agree.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr uint32_t K1 = 0x9E3779B9u, K2 = 0x85EBCA6Bu, K3 = 0xC2B2AE35u;
constexpr int NO_BOUND = 1 << 20;

uint32_t fmix32(uint32_t h) { h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }
uint32_t chain_base(uint64_t seed, uint32_t chain) { return fmix32((uint32_t)seed ^ fmix32((uint32_t)(seed >> 32) + chain * K1 + 0x5bd1e995u)); }
uint32_t step_hash(uint32_t base, uint32_t step) { return fmix32(base ^ (step * K1)); }
uint32_t noise_q7(int noise_pct) { return (uint32_t)((noise_pct * 128 + 50) / 100); }
uint32_t tie_add(uint32_t hs, uint32_t cell) { return (hs * ((2u * cell + 1u) * K2)) >> 16; }
uint32_t tie_remove(uint32_t hs, uint32_t i) { return (hs * ((2u * i + 1u) * K3)) >> 16; }
constexpr uint32_t TABU_BIT = 0x40000000u;
int tenure_of(uint32_t global_chain) { static const int t[4] = {3, 6, 12, 20}; return t[global_chain & 3u]; }
int effective_tenure(int tenure, int k) { int c = k / 3; c = c < 2 ? 2 : c; return tenure < c ? tenure : c; }
uint16_t stamp_reset(uint32_t step) { return (uint16_t)(step - 0x8000u); }
bool is_tabu(uint32_t step, uint16_t stamp, int tenure) { return (uint16_t)((uint16_t)step - stamp) < (uint16_t)tenure; }

struct Model {
    int w, h;
    std::vector<uint8_t> ceil;               // [32*32], index y*32+x
    std::vector<std::vector<int>> reach;     // per site: tiles within geodesic distance <= 3 through ceiling
    // WINDOW mode (window decomposition of grids larger than 32x32, csrc/lns.cu): the tiles that still need cover (ceiling not
    // covered by the frozen supports outside the window's movable core) and the core [core_lo, core_hi)^2 new supports must lie in.
    // Outside WINDOW mode: need = ceil, core = the whole window.
    std::vector<uint8_t> need;
    int core_lo = 0, core_hi = 32;
    bool window = false;

    void build(const uint8_t* grid, int w_, int h_) {
        w = w_; h = h_;
        ceil.assign(1024, 0);
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) ceil[y * 32 + x] = grid[y * w + x] != 0;
        reach.assign(1024, {});
        for (int s = 0; s < 1024; s++) {
            if (!ceil[s]) continue;
            std::vector<uint8_t> sup(1024, 0);
            sup[s] = 1;
            for (int round = 0; round < 3; round++) {  // platform_layout.rs:127-141
                std::vector<uint8_t> nxt = sup;
                for (int t = 0; t < 1024; t++) {
                    if (!sup[t]) continue;
                    int x = t & 31, y = t >> 5;
                    const int nx[4] = {x + 1, x, x - 1, x}, ny[4] = {y, y + 1, y, y - 1};
                    for (int d = 0; d < 4; d++)
                        if (nx[d] >= 0 && nx[d] < 32 && ny[d] >= 0 && ny[d] < 32 && ceil[ny[d] * 32 + nx[d]]) nxt[ny[d] * 32 + nx[d]] = 1;
                }
                sup = nxt;
            }
            for (int t = 0; t < 1024; t++) if (sup[t]) reach[s].push_back(t);
        }
        need = ceil;
    }
};

struct Chain {
    std::vector<uint8_t> S, bestS;  // [1024]
    std::vector<uint8_t> cnt;       // cover count per tile
    std::vector<int> sites;
    std::vector<uint16_t> stamp;    // step of the last flip of every site (16 bit), reset at the start of an epoch
    int k = 0, best = NO_BOUND, done = 0;
    uint32_t step = 0;
    uint64_t scored = 0, steps_done = 0;
};

int pick_rotated(uint32_t bits, uint32_t o) {
    uint32_t rot = o ? ((bits >> o) | (bits << (32 - o))) : bits;
    return (int)((__builtin_ctz(rot) + o) & 31u);
}

struct Runner {
    const Model& M;
    Chain& c;
    uint32_t base;
    int tenure, ten = 1;  // chain tenure; effective tenure of the current step (fixed at the top of the step)
    Runner(const Model& m, Chain& ch, uint32_t b, int t) : M(m), c(ch), base(b), tenure(t) {}

    bool uncovered(int t) const { return M.need[t] && c.cnt[t] == 0; }
    int loss(int u) const { int n = 0; for (int t : M.reach[u]) n += M.need[t] && c.cnt[t] == 1; return n; }   // (tiles the frozen supports cover are no loss)
    int gain(int v) const { int n = 0; for (int t : M.reach[v]) n += M.need[t] && c.cnt[t] == 0; return n; }

    int remove_min_loss(bool use_tabu, uint32_t hs) {
        uint32_t best_key = 0xffffffffu;
        int best_i = 0;
        for (int i = 0; i < c.k; i++) {  // lowest list index wins ties
            int v = c.sites[i];
            uint32_t young = (use_tabu && is_tabu(c.step, c.stamp[v], ten)) ? TABU_BIT : 0u;  // young supports only as a last resort
            uint32_t key = young | ((uint32_t)loss(v) << 16) | tie_remove(hs, (uint32_t)i);
            if (key < best_key) { best_key = key; best_i = i; }
        }
        int u = c.sites[best_i];
        c.sites[best_i] = c.sites[c.k - 1];
        c.sites.pop_back();
        c.k--;
        for (int t : M.reach[u]) c.cnt[t]--;
        c.S[u] = 0;
        c.stamp[u] = (uint16_t)c.step;
        return u;
    }

    void run(long long steps, int epoch_bound, int target, int noise_pct) {
        if (c.done) return;
        if (target >= 0 && epoch_bound <= target) return;  // a layout within the target is already known: the epoch does nothing
        // epoch start: site list in row-major order
        c.sites.clear();
        for (int t = 0; t < 1024; t++) if (c.S[t]) c.sites.push_back(t);
        c.cnt.assign(1024, 0);
        for (int s : c.sites) for (int t : M.reach[s]) c.cnt[t]++;
        c.stamp.assign(1024, stamp_reset(c.step));
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
            bool any = false;
            for (int t = 0; t < 1024 && !any; t++) any = uncovered(t);
            if (!any) {
                c.best = c.k;
                c.bestS = c.S;
                if (c.k <= target || c.k == 0) { c.done = 1; it++; c.step++; break; }
                continue;
            }
            if (c.k == limit - 1 && c.k > 0) {
                c.scored += (uint64_t)c.k;
                remove_min_loss(true, hs);
            }
            uint32_t rowmask = 0;
            for (int y = 0; y < 32; y++) for (int x = 0; x < 32; x++) if (uncovered(y * 32 + x)) rowmask |= 1u << y;
            int y = pick_rotated(rowmask, hs & 31u);
            uint32_t urow = 0;
            for (int x = 0; x < 32; x++) if (uncovered(y * 32 + x)) urow |= 1u << x;
            int x = pick_rotated(urow, (hs >> 5) & 31u);
            int t = y * 32 + x;
            // candidates = tiles of R(t), visited in diamond order (dy, then dx); `lane` = index in the 25-tile diamond
            std::vector<std::pair<int, int>> cand;  // (site, diamond lane)
            int lane = 0;
            for (int dy = -3; dy <= 3; dy++)
                for (int dx = -3; dx <= 3; dx++) {
                    if (std::abs(dx) + std::abs(dy) > 3) continue;
                    int cx = x + dx, cy = y + dy, ln = lane++;
                    if (cx < 0 || cy < 0 || cx >= 32 || cy >= 32) continue;
                    int v = cy * 32 + cx;
                    if (M.window && (cx < M.core_lo || cx >= M.core_hi || cy < M.core_lo || cy >= M.core_hi)) continue;   // only core sites may receive supports
                    if (std::find(M.reach[t].begin(), M.reach[t].end(), v) != M.reach[t].end()) cand.push_back({v, ln});
                }
            int nc = (int)cand.size();
            if (M.window && nc == 0) { c.done = 1; break; }   // cannot happen from a complete start layout; never spin on it
            int v = cand[0].first;
            const bool noise = ((hs >> 10) & 127u) < noise_q7(noise_pct);
            uint32_t mx = 0;
            bool first = true;
            for (auto& [cv, ln] : cand) {
                uint32_t tie = tie_add(hs, (uint32_t)ln);
                uint32_t fresh = is_tabu(c.step, c.stamp[cv], ten) ? 0u : TABU_BIT;  // recently removed sites only as a last resort
                uint32_t key = noise ? (0x10000u | tie) : (fresh | ((uint32_t)(gain(cv) + 1) << 16) | tie);
                if (first || key > mx) { mx = key; v = cv; first = false; }
            }
            if (!noise) c.scored += (uint64_t)nc;
            for (int tt : M.reach[v]) c.cnt[tt]++;
            c.S[v] = 1;
            c.sites.push_back(v);
            c.k++;
            c.stamp[v] = (uint16_t)c.step;
        }
        c.steps_done += (uint64_t)it;
    }
};

}  // namespace

extern "C" {

// Runs `n_chains` model chains (global ids chain_offset..) for the given epochs on one terrain.
// epochs: records (steps, epoch_bound, target); between epochs the caller-visible state is what the kernel persists.
// If share_bound != 0 the bound of epoch e+1 is min(given bound, best over all chains after epoch e) — the
// portfolio's all-reduce-min.
// init_S (nullable): u8[n_chains][1024] start layouts.
// Outputs per chain: S and bestS as u8[1024] (index y*32+x), k, best, step, scored (u64), steps_done (u64).
int tsso_sls_model(const uint8_t* grid, int w, int h, int n_chains, uint32_t chain_offset, uint64_t seed, int noise_pct,
                   const long long* epochs, int n_epochs, int share_bound, const uint8_t* init_S, uint8_t* out_S, uint8_t* out_bestS, int* out_k,
                   int* out_best, uint32_t* out_step, uint64_t* out_scored, uint64_t* out_steps) {
    if (w > 32 || h > 32) return -1;
    Model M;
    M.build(grid, w, h);
    std::vector<Chain> chains(n_chains);
    for (auto& c : chains) { c.S.assign(1024, 0); c.bestS.assign(1024, 0); c.cnt.assign(1024, 0); }
    if (init_S)  // warm start (tss_search_write_chains): current layouts given, k = number of supports
        for (int i = 0; i < n_chains; i++) {
            chains[i].S.assign(init_S + (size_t)i * 1024, init_S + (size_t)(i + 1) * 1024);
            chains[i].k = (int)std::count(chains[i].S.begin(), chains[i].S.end(), (uint8_t)1);
        }
    int shared = NO_BOUND;
    for (int e = 0; e < n_epochs; e++) {
        long long steps = epochs[3 * e];
        int bound = (int)epochs[3 * e + 1], target = (int)epochs[3 * e + 2];
        if (share_bound) bound = std::min(bound, shared);
        for (int i = 0; i < n_chains; i++) Runner(M, chains[i], chain_base(seed, chain_offset + (uint32_t)i), tenure_of(chain_offset + (uint32_t)i)).run(steps, bound, target, noise_pct);
        for (auto& c : chains) shared = std::min(shared, c.best);
    }
    for (int i = 0; i < n_chains; i++) {
        std::memcpy(out_S + (size_t)i * 1024, chains[i].S.data(), 1024);
        std::memcpy(out_bestS + (size_t)i * 1024, chains[i].bestS.data(), 1024);
        out_k[i] = chains[i].k; out_best[i] = chains[i].best; out_step[i] = chains[i].step;
        out_scored[i] = chains[i].scored; out_steps[i] = chains[i].steps_done;
    }
    return 0;
}

// WINDOW mode of the same step rule (csrc/sls.cu sls_kernel<true>, driven by csrc/lns.cu): `terrain` is the window's true ceiling
// (reach is derived from it), `need` the tiles still to be covered, supports only inside [core_lo, core_hi)^2.  All chains start
// from the same layout init_S with best = its support count and step = init_step (lns.cu extract_windows_kernel) and run ONE
// epoch of `steps` steps without a bound or target.  Outputs per chain: bestS u8[1024], best, k.
int tsso_sls_window_model(const uint8_t* terrain, const uint8_t* need, int core_lo, int core_hi, int n_chains, uint32_t chain_offset, uint64_t seed,
                          int noise_pct, long long steps, const uint8_t* init_S, uint32_t init_step, uint8_t* out_bestS, int* out_best, int* out_k) {
    Model M;
    M.build(terrain, 32, 32);
    M.need.assign(need, need + 1024);
    M.core_lo = core_lo; M.core_hi = core_hi; M.window = true;
    for (int i = 0; i < n_chains; i++) {
        Chain c;
        c.S.assign(init_S, init_S + 1024); c.bestS = c.S; c.cnt.assign(1024, 0);
        c.k = (int)std::count(c.S.begin(), c.S.end(), (uint8_t)1);
        c.best = c.k;
        c.step = init_step;
        Runner(M, c, chain_base(seed, chain_offset + (uint32_t)i), tenure_of(chain_offset + (uint32_t)i)).run(steps, NO_BOUND, 0, noise_pct);
        std::memcpy(out_bestS + (size_t)i * 1024, c.bestS.data(), 1024);
        out_best[i] = c.best; out_k[i] = c.k;
    }
    return 0;
}

// constants of the spec, for the agreement test against sls_spec.hpp
void tsso_sls_constants(uint32_t* out) {
    out[0] = K1; out[1] = K2; out[2] = noise_q7(20); out[3] = tie_remove(0x12345678u, 3) + 1000u * (uint32_t)(tenure_of(0) + 2 * tenure_of(1) + 3 * tenure_of(2) + 4 * tenure_of(3)) + (is_tabu(70000u, stamp_reset(70000u), 20) ? 1u : 0u) + (is_tabu(65540u, (uint16_t)65530u, 12) ? 2u : 0u) + 100000u * (uint32_t)(effective_tenure(20, 14) + effective_tenure(3, 100) + effective_tenure(6, 2)); out[4] = tie_add(0x12345678u, 7u);
    out[5] = step_hash(1u, 2u); out[6] = tie_remove(3u, 40u) ^ K3; out[7] = chain_base(0x0123456789abcdefull, 5u); out[8] = NO_BOUND;
}

}  // extern "C"
