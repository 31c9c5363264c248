// Host-only C ABI entry points (include/tss.h): world, encoder, layout decode.  No CUDA here except the one call
// that needs the evaluator's support layers (tss_layout_to_assignment, implemented in engine.cu).
#include <cstring>
#include <new>

#include "encoding_handle.hpp"
#include "engine.hpp"
#include "host_model.hpp"

using namespace tss;

extern "C" int tss_layout_to_assignment_impl(tss_engine* e, const Encoding& enc, const uint8_t* grid, const tss_platform* plats, int32_t n_plats,
                                  uint8_t* assignment);

static std::vector<PlatformLimits::Entry> entries(const int32_t* rec, int n) {
    std::vector<PlatformLimits::Entry> v;
    for (int i = 0; i < n; i++) v.push_back({Dims{rec[3 * i], rec[3 * i + 1]}, (long)rec[3 * i + 2]});
    return v;
}

extern "C" {

int tss_world_parse_toml(const char* text, uint8_t* grid, size_t cap, int32_t* w, int32_t* h, int32_t* ragged, char* err, size_t err_cap) {
    if (!text || !w || !h) return TSS_E_INVALID;
    std::vector<uint8_t> g;
    int gw = 0, gh = 0;
    bool rg = false;
    std::string msg = parse_world_toml(text, g, gw, gh, rg);
    if (!msg.empty()) {
        if (err && err_cap) { std::strncpy(err, msg.c_str(), err_cap - 1); err[err_cap - 1] = 0; }
        return TSS_E_PARSE;
    }
    *w = gw; *h = gh;
    if (ragged) *ragged = rg;
    if (g.size() > cap || !grid) return TSS_E_CAPACITY;
    std::memcpy(grid, g.data(), g.size());
    return TSS_OK;
}

int tss_world_to_toml(const uint8_t* grid, int32_t w, int32_t h, char* out, size_t cap) {
    if (!grid || w <= 0 || h <= 0) return TSS_E_INVALID;
    std::string s = world_to_toml(grid, w, h);
    if (!out || s.size() + 1 > cap) return TSS_E_CAPACITY;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

int tss_world_synthetic(int32_t w, int32_t h, uint64_t seed, uint64_t t, uint32_t density_q24, uint8_t* grid) {
    if (!grid || w <= 0 || h <= 0) return TSS_E_INVALID;
    synthetic_world(w, h, seed, t, density_q24, grid);
    return TSS_OK;
}

int tss_encoding_create(const uint8_t* grid, int32_t w, int32_t h, const tss_dims* defs, int32_t n_defs, tss_encoding** out) {
    if (!out || !grid || !defs || n_defs <= 0 || w <= 0 || h <= 0) return TSS_E_INVALID;
    auto data = std::make_shared<tss_encoding_data>();
    std::vector<Dims> d;
    for (int i = 0; i < n_defs; i++) d.push_back(Dims{defs[i].w, defs[i].h});
    std::string msg = Encoding::encode(grid, w, h, d, data->enc);
    if (!msg.empty()) return TSS_E_INVALID;
    data->grid.assign(grid, grid + (size_t)w * h);
    tss_encoding* enc = new (std::nothrow) tss_encoding();
    if (!enc) return TSS_E_INVALID;
    enc->d = std::move(data);
    *out = enc;
    return TSS_OK;
}

void tss_encoding_destroy(tss_encoding* enc) { delete enc; }

int tss_encoding_sizes(const tss_encoding* enc, int32_t* n_vars, int32_t* n_clauses, int64_t* n_lits, int32_t* n_dims) {
    if (!enc) return TSS_E_INVALID;
    if (n_vars) *n_vars = enc->d->enc.base.n_vars;
    if (n_clauses) *n_clauses = enc->d->enc.base.n_clauses();
    if (n_lits) *n_lits = (int64_t)enc->d->enc.base.lits.size();
    if (n_dims) *n_dims = enc->d->enc.K();
    return TSS_OK;
}

int tss_encoding_dims(const tss_encoding* enc, tss_dims* out_dims) {
    if (!enc || !out_dims) return TSS_E_INVALID;
    for (int k = 0; k < enc->d->enc.K(); k++) out_dims[k] = tss_dims{enc->d->enc.keys[k].w, enc->d->enc.keys[k].h};
    return TSS_OK;
}

int tss_encoding_var_maps(const tss_encoding* enc, int32_t* plat_var, int32_t* terr_var) {
    if (!enc) return TSS_E_INVALID;
    if (plat_var) std::memcpy(plat_var, enc->d->enc.plat_var.data(), enc->d->enc.plat_var.size() * sizeof(int32_t));
    if (terr_var) std::memcpy(terr_var, enc->d->enc.terr_var.data(), enc->d->enc.terr_var.size() * sizeof(int32_t));
    return TSS_OK;
}

int tss_encoding_cnf(const tss_encoding* enc, int32_t* lits, uint32_t* offsets) {
    if (!enc || !offsets) return TSS_E_INVALID;
    const Cnf& f = enc->d->enc.base;
    if (lits) std::memcpy(lits, f.lits.data(), f.lits.size() * sizeof(int32_t));
    std::memcpy(offsets, f.offsets.data(), f.offsets.size() * sizeof(uint32_t));
    return TSS_OK;
}

int tss_encoding_with_limits(const tss_encoding* enc, const int32_t* card, int32_t n_card, const int32_t* weights, int32_t n_weights,
                             int32_t has_weight_limit, int64_t weight_limit, int32_t* n_vars, int32_t* n_clauses, int64_t* n_lits,
                             int32_t* lits, uint32_t* offsets) {
    if (!enc || (n_card > 0 && !card) || (n_weights > 0 && !weights)) return TSS_E_INVALID;
    PlatformLimits lim;
    lim.card_limits = entries(card, n_card);
    lim.weights = entries(weights, n_weights);
    lim.has_weight_limit = has_weight_limit != 0;
    lim.weight_limit = (long)weight_limit;
    auto same_entries = [](const std::vector<PlatformLimits::Entry>& a, const std::vector<PlatformLimits::Entry>& b) {
        if (a.size() != b.size()) return false;
        for (size_t i = 0; i < a.size(); i++)
            if (a[i].def.w != b[i].def.w || a[i].def.h != b[i].def.h || a[i].value != b[i].value) return false;
        return true;
    };
    std::shared_ptr<const Cnf> lowered;
    {
        std::lock_guard<std::mutex> lock(enc->d->lowered_mutex);
        const PlatformLimits& c = enc->d->lowered_limits;
        if (!enc->d->lowered || !same_entries(c.card_limits, lim.card_limits) || !same_entries(c.weights, lim.weights) || c.has_weight_limit != lim.has_weight_limit ||
            (lim.has_weight_limit && c.weight_limit != lim.weight_limit)) {
            enc->d->lowered = std::make_shared<const Cnf>(enc->d->enc.with_limits(lim));
            enc->d->lowered_limits = lim;
        }
        lowered = enc->d->lowered;
    }
    const Cnf& f = *lowered;
    if (n_vars) *n_vars = f.n_vars;
    if (n_clauses) *n_clauses = f.n_clauses();
    if (n_lits) *n_lits = (int64_t)f.lits.size();
    if (lits && offsets) {
        std::memcpy(lits, f.lits.data(), f.lits.size() * sizeof(int32_t));
        std::memcpy(offsets, f.offsets.data(), f.offsets.size() * sizeof(uint32_t));
        instance_record(enc->d, lim, lowered);   // the solver that receives these clauses can find its instance again (tss_instance_find)
    }
    return TSS_OK;
}

int tss_layout_from_assignment(const tss_encoding* enc, const uint8_t* assignment, int32_t n_assignment, tss_platform* out, int32_t cap,
                               int32_t* n_out) {
    if (!enc || !assignment || !n_out) return TSS_E_INVALID;
    std::vector<tss_platform> p = enc->d->enc.layout_from_assignment(assignment, n_assignment);
    *n_out = (int32_t)p.size();
    if ((int32_t)p.size() > cap || (!out && !p.empty())) return TSS_E_CAPACITY;
    if (!p.empty()) std::memcpy(out, p.data(), p.size() * sizeof(tss_platform));
    return TSS_OK;
}

int tss_layout_to_assignment(tss_engine* e, const tss_encoding* enc, const tss_platform* plats, int32_t n_plats, uint8_t* assignment) {
    if (!e) return TSS_E_INVALID;
    if (!enc || !assignment || n_plats < 0 || (!plats && n_plats > 0)) return e->fail(TSS_E_INVALID, "tss_layout_to_assignment: bad arguments");
    return tss_layout_to_assignment_impl(e, enc->d->enc, enc->d->grid.data(), plats, n_plats, assignment);
}

int tss_layout_trivial_optimization(const uint8_t* grid, int32_t w, int32_t h, tss_platform* plats, int32_t n) {
    if (!grid || w <= 0 || h <= 0 || n < 0 || (!plats && n > 0)) return TSS_E_INVALID;
    return trivial_optimization(grid, w, h, plats, n);
}

int64_t tss_layout_total_weight(const tss_platform* plats, int32_t n, const int32_t* weights, int32_t n_weights) {
    if ((n > 0 && !plats) || (n_weights > 0 && !weights)) return 0;
    return total_weight(plats, n, entries(weights, n_weights));
}

int tss_platform_overlaps(const tss_platform* a, const tss_platform* b) {
    if (!a || !b) return 0;
    return platform_overlaps(*a, *b);
}

}  // extern "C"
