// Host-side mirror of the reference's library API for the hot path (lib crate `timberborn-platform-cruncher`):
// World/WorldGrid, PlatformDef/Platform, Encoding::{encode, with_limits, vars}, PlatformLimits,
// PlatformLayout::{from_assignment, run_trivial_optimization, total_weight}.  Plain C++17, no CUDA, no torch.
// (validate() is NOT here: it is kernel (a), see eval.cu.)  Citations are reference file:line.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/tss.h"

namespace tss {

constexpr int kTerrainSupportDistance = 4;  // src/lib.rs:12

struct Dims {
    int w = 0, h = 0;
    bool operator==(const Dims& o) const { return w == o.w && h == o.h; }
};
// src/math/dimensions.rs:74-114: -1 less, 0 equal, 1 greater, 2 incomparable (non-empty dims only on this path)
int dims_partial_cmp(Dims a, Dims b);
inline bool dims_le(Dims a, Dims b) { int c = dims_partial_cmp(a, b); return c == -1 || c == 0; }

// Bit-packed grid rows: wpr = ceil(w/32) u32 words per row (tss.h conventions).
struct BitGrid {
    int w = 0, h = 0, wpr = 0;
    std::vector<uint32_t> rows;
    BitGrid() = default;
    BitGrid(int w_, int h_) : w(w_), h(h_), wpr((w_ + 31) / 32), rows((size_t)h_ * ((w_ + 31) / 32), 0u) {}
    bool get(int x, int y) const { return x >= 0 && y >= 0 && x < w && y < h && ((rows[(size_t)y * wpr + (x >> 5)] >> (x & 31)) & 1u); }
    void set(int x, int y) { rows[(size_t)y * wpr + (x >> 5)] |= 1u << (x & 31); }
    static BitGrid from_bytes(const uint8_t* g, int w, int h);
    int count() const;
};

// src/world.rs:49-79 (+ :21-40).  Returns "" or an error message.
std::string parse_world_toml(const char* text, std::vector<uint8_t>& grid, int& w, int& h, bool& ragged);
std::string world_to_toml(const uint8_t* grid, int w, int h);
void synthetic_world(int w, int h, uint64_t seed, uint64_t t, uint32_t density_q24, uint8_t* grid);

// CSR CNF (DIMACS-signed literals, 1-based variables)
struct Cnf {
    int n_vars = 0;
    std::vector<int32_t> lits;
    std::vector<uint32_t> offsets{0};
    int n_clauses() const { return (int)offsets.size() - 1; }
    int new_var() { return ++n_vars; }
    void add1(int a) { lits.push_back(a); close(); }
    void add2(int a, int b) { lits.push_back(a); lits.push_back(b); close(); }
    void add3(int a, int b, int c) { lits.push_back(a); lits.push_back(b); lits.push_back(c); close(); }
    void close() { offsets.push_back((uint32_t)lits.size()); }
};

struct PlatformLimits {  // src/encoder/platform_limits.rs:6-13
    struct Entry { Dims def; long value; };
    std::vector<Entry> card_limits, weights;
    bool has_weight_limit = false;
    long weight_limit = 0;
};

struct Encoding {  // src/encoder.rs:428-432
    int w = 0, h = 0;
    std::vector<Dims> defs;       // canonical platform defs
    std::vector<Dims> keys;       // dims keys incl. flipped variants, variable order
    std::vector<int> key_def;     // key -> index into defs
    std::vector<int32_t> plat_var;  // [tile*K + k]
    std::vector<int32_t> terr_var;  // [tile*4 + layer], 0 = no ceiling
    Cnf base;
    int K() const { return (int)keys.size(); }
    int key_index(Dims d) const;  // -1 if absent
    // src/encoder.rs:435-613.  Returns "" or an error (platform set without 1x1, empty grid).
    static std::string encode(const uint8_t* grid, int w, int h, const std::vector<Dims>& defs, Encoding& out);
    // src/encoder.rs:619-667 + into_cnf()
    Cnf with_limits(const PlatformLimits& limits) const;
    // src/encoder/platform_layout.rs:26-52
    std::vector<tss_platform> layout_from_assignment(const uint8_t* assignment, int n) const;
};

// src/platform.rs:86-97
bool platform_overlaps(const tss_platform& a, const tss_platform& b);
inline Dims platform_dims(const tss_platform& p) { return p.rotated ? Dims{p.def_h, p.def_w} : Dims{p.def_w, p.def_h}; }
// src/encoder/platform_layout.rs:151-172 ; :174-183
int trivial_optimization(const uint8_t* grid, int w, int h, tss_platform* plats, int n);
long total_weight(const tss_platform* plats, int n, const std::vector<PlatformLimits::Entry>& weights);

}  // namespace tss
