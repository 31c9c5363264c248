// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's feasibility-and-bound path
// (MetaflameDragon/timberborn_support_solver, citations are file:line relative
// to the reference tree).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call anything in
// this directory.  The product (timberborn_support_solver_b200/) never does.
//
// Parity status: the reference has NO tests/golden vectors for encoder,
// validator or solver loop and cannot be built here (no
// Rust toolchain).  The restatement is pinned against what does exist:
//   * src/platform.rs:148-232 overlap tables (40 cases)          -> tests/golden/platform_overlap.json
//   * src/math/point.rs:146, src/math/dimensions.rs:169           -> unit tests
//   * README.md:47-116 four 1x1 layouts (18/17/16/15)             -> tests/golden/readme_layouts.json
//   * src/encoder.rs:45-51 doc-comment DAG diagram                -> tests/golden/dag_default8.json
// The cardinality / PB CNF (rustsat 0.7.2 `into_cnf`, source not in the tree)
// is "parity unpinned": restated from the published totalizer / generalized
// totalizer encodings, and checked semantically only.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <optional>
#include <array>

namespace tsso {

// ---- src/math/point.rs:14-17 -------------------------------------------------
struct Point {
    long x = 0, y = 0;
    bool operator==(const Point& o) const { return x == o.x && y == o.y; }
    bool operator<(const Point& o) const { return x != o.x ? x < o.x : y < o.y; }  // derive(Ord): x then y
};
inline Point operator+(Point a, Point b) { return {a.x + b.x, a.y + b.y}; }
inline Point operator-(Point a, Point b) { return {a.x - b.x, a.y - b.y}; }
// src/math/point.rs:46-53 — order +x, +y, -x, -y
void neighbors(Point p, Point out[4]);
// src/math/point.rs:110-132 IterManhattan (unused by the encoder; pinned by point.rs:135-152)
std::vector<Point> iter_within_manhattan(Point c, unsigned dist);

// ---- src/math/dimensions.rs:16-19 ---------------------------------------------
struct Dims {
    unsigned long width = 0, height = 0;
    bool empty() const { return width == 0 || height == 0; }                   // dimensions.rs:53-55
    Dims flipped() const { return {height, width}; }                           // dimensions.rs:34-36
    bool contains(Point p) const {                                             // dimensions.rs:46-51
        return p.x >= 0 && p.x < (long)width && p.y >= 0 && p.y < (long)height;
    }
    bool operator==(const Dims& o) const {                                     // dimensions.rs:68-72
        return (empty() && o.empty()) || (width == o.width && height == o.height);
    }
    bool operator<(const Dims& o) const;  // TOTAL order for std::map keys only (not the partial order)
};
enum class POrd { Less, Equal, Greater, None };
POrd partial_cmp(Dims a, Dims b);                                              // dimensions.rs:74-114
std::vector<Point> iter_within(Dims d);                                        // dimensions.rs:138-156 row-major, x fastest

// ---- src/math/grid.rs:8-11 ----------------------------------------------------
template <class T>
struct Grid {
    std::vector<T> data;
    Dims dims;
    std::optional<size_t> data_index(Point p) const {                          // grid.rs:66-68
        if (!dims.contains(p)) return std::nullopt;
        return (size_t)p.x + (size_t)p.y * dims.width;
    }
    const T* get(Point p) const { auto i = data_index(p); return i ? &data[*i] : nullptr; }
    T* get_mut(Point p) { auto i = data_index(p); return i ? &data[*i] : nullptr; }
    Point index_to_point(size_t i) const { return {(long)(i % dims.width), (long)(i / dims.width)}; }  // grid.rs:70-72
};

// ---- src/world.rs:19, 49-79 ----------------------------------------------------
using WorldGrid = Grid<uint8_t>;  // Grid<bool> in the reference; u8 avoids std::vector<bool> proxies
// Parses the `[world] grid = [ "XX ", ... ]` project file.  Returns "" on success, else an error text.
// NOTE (divergence documented in DESIGN.md): the reference's doc comment promises ragged rows are
// left-aligned and padded with `false` (world.rs:82-86) but `copy_from_slice` (world.rs:73) panics on a
// length mismatch; this restatement implements the documented padding and reports `ragged=true`.
std::string parse_world_toml(const std::string& text, WorldGrid& out, bool* ragged = nullptr);
std::string world_to_toml(const WorldGrid& g);                                 // world.rs:21-40

// ---- src/platform.rs ------------------------------------------------------------
struct PlatformDef {
    Dims dims;                                                                  // platform.rs:11-15
    bool rectangular() const { return dims.width != dims.height; }              // platform.rs:51-53
    bool operator==(const PlatformDef& o) const { return dims.width == o.dims.width && dims.height == o.dims.height; }
    bool operator<(const PlatformDef& o) const { return dims < o.dims; }
};
extern const PlatformDef PLATFORMS_DEFAULT[8];                                  // platform.rs:23-32
struct Platform {
    Point point; PlatformDef def; bool rotated = false;                         // platform.rs:64-70
    Dims dims() const { return rotated ? def.dims.flipped() : def.dims; }       // platform.rs:111-113
    bool overlaps(const Platform& o) const;                                     // platform.rs:86-97
    bool operator<(const Platform& o) const;
    bool operator==(const Platform& o) const { return point == o.point && def == o.def && rotated == o.rotated; }
};

constexpr int TERRAIN_SUPPORT_DISTANCE = 4;                                     // src/lib.rs:12

// ---- src/encoder.rs --------------------------------------------------------------
// Literals are DIMACS-signed ints over 1-based variables (rustsat Var idx + 1).
using Clause = std::vector<int>;
enum Family : uint8_t {
    F_DAG_IMPL = 0, F_DAG_SIBLING = 1, F_T3_PLATFORM = 2, F_LAYER = 3, F_UNIT_T0 = 4,
    F_OVERLAP_ANCHOR = 5, F_OVERLAP_CROSS = 6, F_OOB = 7, F_LIMIT_LINK = 8, F_CARD = 9, F_PB = 10,
};
struct SatInstance {
    int n_vars = 0;
    std::vector<Clause> clauses;
    std::vector<uint8_t> family;
    int new_var() { return ++n_vars; }
    void add(Clause c, Family f) { clauses.push_back(std::move(c)); family.push_back(f); }
    // rustsat SatInstance helpers as used at encoder.rs:454,479,512,535,543,570,594,608,638
    void add_unit(int l, Family f) { add({l}, f); }
    void add_lit_impl_lit(int a, int b, Family f) { add({-a, b}, f); }
    void add_lit_impl_clause(int a, const std::vector<int>& c, Family f) {
        Clause cl{-a}; cl.insert(cl.end(), c.begin(), c.end()); add(std::move(cl), f);
    }
    void add_cube_impl_clause(const std::vector<int>& cube, const std::vector<int>& c, Family f) {
        Clause cl; for (int l : cube) cl.push_back(-l); cl.insert(cl.end(), c.begin(), c.end()); add(std::move(cl), f);
    }
};

struct EncodingTileVars {                                                       // encoder.rs:157-167
    std::map<Dims, int> dims_vars;
    std::optional<std::array<int, TERRAIN_SUPPORT_DISTANCE>> terrain;
};
struct EncodedItem { bool is_platform; Point point; Dims dims; int layer; };     // encoder.rs:169-173

struct EncodingVars {                                                           // encoder.rs:175-279
    std::vector<Dims> dim_keys;               // deterministic replacement for HashMap key order
    std::map<Dims, PlatformDef> dim_map;      // dims (incl. flipped) -> def (sets are singletons for distinct defs)
    Grid<EncodingTileVars> grid;
    std::map<int, EncodedItem> var_map;
    std::optional<Platform> var_to_platform(int var) const;                     // encoder.rs:232-249
};

// encoder.rs:281-303 EncodingNode + partial order; encoder.rs:308-426 EncodingDag
struct EncodingNode { bool is_platform; Dims dims; Point point; };
struct EncodingDag {
    std::vector<EncodingNode> nodes;                 // isolated point nodes dropped (encoder.rs:331)
    std::vector<std::vector<bool>> closure;          // closure[i][j]: i ->+ j  (smaller -> larger)
    std::vector<std::vector<bool>> reduced;          // transitive reduction
    explicit EncodingDag(const std::vector<Dims>& platform_dims);
    std::vector<std::pair<Dims, Dims>> platform_edges_reduced() const;            // (smaller, larger) encoder.rs:355-362
    std::vector<std::pair<Point, Dims>> point_platform_edges_reduced() const;     // encoder.rs:368-373
    // encoder.rs:460-489: for every platform node, every 2-combination (a,b) of its reduced platform
    // successors, with the minimal common strict successors of a and b.
    struct Sibling { Dims a, b; std::vector<Dims> lcub; };
    std::vector<Sibling> sibling_clauses() const;
};

struct PlatformLimits {                                                          // platform_limits.rs:6-13
    std::vector<std::pair<PlatformDef, unsigned long>> card_limits;
    std::vector<std::pair<PlatformDef, long>> weights;
    std::optional<long> weight_limit;
};

struct Encoding {                                                                // encoder.rs:428-432
    EncodingVars vars;
    SatInstance instance;
    static Encoding encode(const std::vector<PlatformDef>& defs, const WorldGrid& terrain);   // encoder.rs:435-613
    // encoder.rs:619-667 + rustsat `into_cnf` (totalizer / generalized totalizer; parity unpinned)
    SatInstance with_limits(const PlatformLimits& limits) const;
};

// Assignment: index = var (1-based), value 1 = True, 0 = False, 2 = DontCare (rustsat TernaryVal)
using Assignment = std::vector<uint8_t>;

struct ValidationResult {                                                        // platform_layout.rs:187-191
    std::set<Point> unsupported_terrain;
    std::set<Platform> overlapping_platforms;
    std::set<Platform> out_of_bounds_platforms;
    bool is_valid() const { return unsupported_terrain.empty() && overlapping_platforms.empty() && out_of_bounds_platforms.empty(); }
};

struct PlatformLayout {                                                          // platform_layout.rs:20-23
    std::map<Point, Platform> platforms;
    static PlatformLayout from_assignment(const Assignment& a, const EncodingVars& vars);     // :26-52
    size_t platform_count() const { return platforms.size(); }                               // :58-60
    std::map<PlatformDef, size_t> platform_stats() const;                                     // :66-79
    ValidationResult validate(const WorldGrid& world) const;                                  // :85-149
    void run_trivial_optimization(const WorldGrid& world);                                    // :151-172
    long total_weight(const std::vector<std::pair<PlatformDef, long>>& weights) const;        // :174-183
};
long assignment_total_weight(const Assignment& a, const EncodingVars& vars,
                             const std::vector<std::pair<PlatformDef, long>>& weights);        // encoder.rs:670-692

// ---- CDCL SAT solver: stand-in for rustsat-glucose 0.7.2 (Glucose 4, C++, crates.io, not in the tree).
// Restates the published algorithm: MiniSat-style two-watched-literal CDCL with first-UIP learning and
// clause minimisation, VSIDS + phase saving, and Glucose's LBD clause scoring, dynamic (LBD-queue)
// restarts with trail-size blocking and aggressive learnt-clause reduction (Audemard & Simon, IJCAI'09,
// CP'12).  Results are IPASIR-style: 10 = SAT, 20 = UNSAT, 0 = interrupted / budget exhausted.
struct SolveStats { uint64_t conflicts = 0, decisions = 0, propagations = 0, restarts = 0, learnts = 0; double seconds = 0; };
int solve_cnf(int n_vars, const std::vector<Clause>& clauses, Assignment& out, SolveStats* stats,
              int64_t conflict_budget, const volatile int* interrupt);

// ---- crates/repl/src/main.rs:280-366 solver_loop ------------------------------------
struct LoopStep { long bound; int result; size_t count; bool valid; SolveStats stats; };
struct LoopResult { std::vector<LoopStep> steps; PlatformLayout best; bool proved_optimal = false; };
LoopResult solver_loop(const WorldGrid& world, const std::vector<PlatformDef>& defs, PlatformLimits limits,
                       int64_t conflict_budget_per_solve, const volatile int* interrupt);

}  // namespace tsso
