/* The drop-in boundary used from plain C (no Python, no torch): what a Rust/C host does through include/tss.h.
 * Built and run by tests/test_gpu.py::test_c_abi_from_plain_c with `gcc c_abi_smoke.c -ltss`.
 * Exit code 0 = every check passed; a failed check prints its line and exits 1. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tss.h"

#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) {                                                           \
            fprintf(stderr, "c_abi_smoke: check failed at line %d: %s (%s)\n", __LINE__, #cond, e ? tss_last_error(e) : ""); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(void) {
    tss_engine* e = NULL;
    CHECK(tss_version() == TSS_VERSION);
    CHECK(tss_engine_create(0, &e) == TSS_OK);

    /* test/ex1.toml-shaped project through the loader (src/world.rs:49-79) */
    const char* toml = "[world]\ngrid = [\n \"XXXXX\",\n \"XXXXX\",\n \"X  XX\",\n \"X   X\",\n \"X    \",\n \"XXX  \",\n]\n";
    uint8_t grid[64];
    int32_t w = 0, h = 0, ragged = 0;
    char err[128];
    CHECK(tss_world_parse_toml(toml, grid, sizeof grid, &w, &h, &ragged, err, sizeof err) == TSS_OK);
    CHECK(w == 5 && h == 6 && !ragged);
    CHECK(tss_world_parse_toml("[world]\ngrid = [\"X.X\"]\n", grid, sizeof grid, &w, &h, &ragged, err, sizeof err) == TSS_E_PARSE);
    CHECK(tss_world_parse_toml(toml, grid, sizeof grid, &w, &h, &ragged, err, sizeof err) == TSS_OK);

    /* solve with the REPL's platform set: optimum is one 5x5 */
    const tss_dims defs[8] = {{1, 1}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {3, 3}, {5, 5}};
    tss_platform plats[64];
    int32_t n = 0;
    CHECK(tss_solve_upper_bound(e, grid, w, h, defs, 8, 1, 7, 0, 20000, plats, 64, &n) == TSS_SAT);
    CHECK(n == 1 && plats[0].def_w == 5 && plats[0].def_h == 5);
    /* validate() on the GPU agrees, flags stay clear */
    uint8_t unsupported[64], flags[64];
    CHECK(tss_validate(e, grid, w, h, plats, n, unsupported, flags) == 0 && flags[0] == 0);
    /* capacity error instead of a buffer overrun */
    CHECK(tss_solve_upper_bound(e, grid, w, h, defs, 1, 3, 7, 0, 20000, plats, 2, &n) == TSS_E_CAPACITY && n == 3);
    CHECK(strlen(tss_last_error(e)) > 0);
    /* a platform set without 1x1 is rejected, not UB (src/encoder.rs:564-566) */
    const tss_dims bad[1] = {{3, 3}};
    CHECK(tss_solve_upper_bound(e, grid, w, h, bad, 1, -1, 7, 0, 100, plats, 64, &n) == TSS_E_INVALID);

    /* encoder + CNF check of the witness (kernel c) */
    tss_encoding* enc = NULL;
    CHECK(tss_encoding_create(grid, w, h, defs, 8, &enc) == TSS_OK);
    int32_t n_vars = 0, n_clauses = 0, n_dims = 0;
    int64_t n_lits = 0;
    CHECK(tss_encoding_sizes(enc, &n_vars, &n_clauses, &n_lits, &n_dims) == TSS_OK);
    CHECK(n_vars == 466 && n_clauses == 1625 && n_lits == 3072 && n_dims == 13); /* SURVEY.md §6 */
    int32_t* lits = malloc(sizeof(int32_t) * (size_t)n_lits);
    uint32_t* offsets = malloc(sizeof(uint32_t) * (size_t)(n_clauses + 1));
    uint8_t* assignment = malloc((size_t)n_vars + 1);
    CHECK(lits && offsets && assignment);
    CHECK(tss_encoding_cnf(enc, lits, offsets) == TSS_OK);
    CHECK(tss_solve_upper_bound(e, grid, w, h, defs, 8, 1, 7, 0, 20000, plats, 64, &n) == TSS_SAT);
    CHECK(tss_layout_to_assignment(e, enc, plats, n, assignment) == TSS_OK);
    tss_cnf* cnf = NULL;
    CHECK(tss_cnf_upload(e, lits, offsets, n_clauses, n_vars, &cnf) == TSS_OK);
    int32_t n_falsified = -1, first = 0;
    CHECK(tss_cnf_check(e, cnf, assignment, 1, &n_falsified, &first) == TSS_OK);
    CHECK(n_falsified == 0 && first == -1);
    tss_platform decoded[64];
    int32_t n_dec = 0;
    CHECK(tss_layout_from_assignment(enc, assignment, n_vars + 1, decoded, 64, &n_dec) == TSS_OK);
    CHECK(n_dec == 1 && decoded[0].x == plats[0].x && decoded[0].y == plats[0].y && decoded[0].def_w == 5);

    /* interrupt: a pending interrupt makes the next solve return UNKNOWN, clearing it restores service */
    tss_interrupt(e);
    CHECK(tss_solve_upper_bound(e, grid, w, h, defs, 1, 3, 7, 100, 0, plats, 64, &n) == TSS_UNKNOWN);
    tss_stats st;
    CHECK(tss_get_stats(e, &st) == TSS_OK && st.interrupted == 1);
    tss_clear_interrupt(e);
    CHECK(tss_solve_upper_bound(e, grid, w, h, defs, 1, 3, 7, 0, 20000, plats, 64, &n) == TSS_SAT && n == 3);

    /* the SAT-like call the bound-tightening loop makes: first layout within the bound, one fused launch once the engine
     * holds a workspace; an infeasible bound (ex1 needs 3 supports) with a give-up point comes back UNKNOWN after exactly
     * that many steps per chain — timed here as a plain C host sees it */
    CHECK(tss_solve_upper_bound(e, grid, w, h, defs, 1, 3, 8, 0, 0, plats, 64, &n) == TSS_SAT && n == 3);
    CHECK(tss_solve_upper_bound(e, grid, w, h, defs, 1, 3, 9, 0, 0, plats, 64, &n) == TSS_SAT && n == 3);
    CHECK(tss_get_stats(e, &st) == TSS_OK && st.last_solve_steps > 0 && st.last_solve_steps <= 96);
    CHECK(tss_solve_upper_bound(e, grid, w, h, defs, 1, 2, 9, 0, -500, plats, 64, &n) == TSS_UNKNOWN && n == 0);
    CHECK(tss_get_stats(e, &st) == TSS_OK && st.last_solve_steps == 500);

    /* a persistent portfolio with a warm start: every chain begins from the layout just found */
    tss_search* search = NULL;
    tss_search_params params = {11, 16, 0, -1, TSS_KERNEL_AUTO};
    CHECK(tss_search_create(e, grid, w, h, defs, 1, &params, &search) == TSS_OK);
    CHECK(tss_solve_upper_bound(e, grid, w, h, defs, 1, 3, 9, 0, 0, plats, 64, &n) == TSS_SAT && n == 3);
    uint32_t rows[16 * 32];
    memset(rows, 0, sizeof rows);
    for (int c = 0; c < 16; c++)
        for (int i = 0; i < n; i++) rows[c * 32 + plats[i].y] |= 1u << plats[i].x;
    CHECK(tss_search_write_chains(search, rows) == TSS_OK);
    CHECK(tss_search_run(search, 8, 0) == TSS_OK);
    int32_t best = -1;
    CHECK(tss_search_best_count(search, &best) == TSS_OK && best == 3);   /* complete at once: the warm start is recorded as the best layout */
    rows[5] = 1u << 9;                                                     /* a support outside the 5x6 grid is rejected */
    CHECK(tss_search_write_chains(search, rows) == TSS_E_INVALID);
    tss_search_destroy(search);

    /* ---- round 2: certified lower bounds and the instance bridge, as a plain C host uses them */
    int32_t xy[2 * 64], n_packed = -1;
    CHECK(tss_lower_bound(e, grid, w, h, defs, 1, 1, 0, xy, 64, &n_packed) == TSS_OK && n_packed == 3);      /* ex1, 1x1 supports: 3 = the optimum */
    CHECK(tss_lower_bound(e, grid, w, h, defs, 8, 1, 0, xy, 64, &n_packed) == TSS_OK && n_packed == 1);      /* default-8: one 5x5 covers everything */
    CHECK(tss_lower_bound(e, grid, w, h, defs, 1, 1, 0, xy, 2, &n_packed) == TSS_E_CAPACITY && n_packed == 3);
    CHECK(tss_lower_bound(e, grid, w, h, defs, 1, 1, 0, NULL, 0, &n_packed) == TSS_OK && n_packed == 3);      /* the bound alone */
    CHECK(tss_lower_bound(e, grid, w, h, bad, 1, 1, 0, xy, 64, &n_packed) == TSS_E_INVALID);
    CHECK(tss_lower_bound(e, NULL, w, h, defs, 1, 1, 0, xy, 64, &n_packed) == TSS_E_INVALID);
    int32_t lp_w[64], lp_info[3];
    int64_t lp_total = 0, lp_max = 0, lp_bound = -1;
    CHECK(tss_lower_bound_lp(e, grid, w, h, defs, 1, NULL, 0, 0, 0, lp_w, &lp_total, &lp_max, &lp_bound, lp_info) == TSS_OK);
    CHECK(lp_bound == 3 && lp_max > 0 && lp_bound == (lp_total + lp_max - 1) / lp_max && lp_info[1] == 1 && lp_info[2] == 19);
    int64_t sum = 0;
    for (int i = 0; i < w * h; i++) { CHECK(lp_w[i] >= 0 && (grid[i] || lp_w[i] == 0)); sum += lp_w[i]; }
    CHECK(sum == lp_total);
    CHECK(tss_lower_bound_lp(e, grid, w, h, defs, 1, NULL, 0, 0, 3, NULL, NULL, NULL, &lp_bound, NULL) == TSS_OK && lp_bound == 3);
    CHECK(tss_lower_bound_lp(e, grid, w, h, defs, 1, NULL, 0, 0, 0, NULL, NULL, NULL, NULL, NULL) == TSS_E_INVALID);
    {   /* the GUI's objective (app.rs:53-62): 1x1 = 5, 1xN = 1, 3x3 = 2, 5x5 = 4; one 5x5 costs 5+1+1+1+1+2+4 = 15 and is the optimum */
        const int32_t gui_w[24] = {1, 1, 5, 1, 2, 1, 1, 3, 1, 1, 4, 1, 1, 5, 1, 1, 6, 1, 3, 3, 2, 5, 5, 4};
        int64_t wb = -1;
        CHECK(tss_lower_bound_lp(e, grid, w, h, defs, 8, gui_w, 8, 0, 0, NULL, NULL, NULL, &wb, NULL) == TSS_OK && wb >= 5 && wb <= 15);
    }

    /* Solve::add_cnf + Solve::solve of the Rust shim: the solver is handed clauses only (solver_runner.rs:8-20) */
    {
        const int32_t card[3] = {1, 1, 3};                      /* PlatformLimits.card_limits[1x1] = 3 (main.rs:346) */
        tss_encoding* enc1 = NULL;
        CHECK(tss_encoding_create(grid, w, h, defs, 1, &enc1) == TSS_OK);
        int32_t nv = 0, nc = 0;
        int64_t nl = 0;
        CHECK(tss_encoding_with_limits(enc1, card, 1, NULL, 0, 0, 0, &nv, &nc, &nl, NULL, NULL) == TSS_OK);   /* sizes only: records nothing */
        int32_t* l2 = malloc(sizeof(int32_t) * (size_t)(nl + 1));
        uint32_t* o2 = malloc(sizeof(uint32_t) * (size_t)(nc + 1));
        uint8_t* a2 = malloc((size_t)nv + 1);
        CHECK(l2 && o2 && a2);
        CHECK(tss_encoding_with_limits(enc1, card, 1, NULL, 0, 0, 0, &nv, &nc, &nl, l2, o2) == TSS_OK);       /* ... this one records the instance */
        tss_encoding_destroy(enc1);                              /* the owner may drop its handle: the registry keeps the instance */
        tss_cnf* c2 = NULL;
        CHECK(tss_cnf_upload(e, l2, o2, nc, nv, &c2) == TSS_OK && tss_cnf_num_vars(c2) == nv);
        tss_encoding* inst = NULL;
        tss_instance_info info;
        CHECK(tss_instance_find(l2, o2, nc, nv, &inst, &info, NULL, 0) == TSS_SAT);
        CHECK(info.exact == 1 && info.w == w && info.h == h && info.n_defs == 1 && info.card_limit_1x1 == 3 && info.n_other_card_limits == 0);
        uint8_t g2[64];
        int32_t w2 = 0, h2 = 0;
        CHECK(tss_encoding_terrain(inst, g2, sizeof g2, &w2, &h2) == TSS_OK && w2 == w && h2 == h && memcmp(g2, grid, (size_t)(w * h)) == 0);
        CHECK(tss_solve_instance(e, c2, inst, &info, NULL, 5, 2000, a2) == TSS_SAT);                          /* a verified model of exactly these clauses */
        int32_t nf = -1;
        CHECK(tss_cnf_check(e, c2, a2, 1, &nf, NULL) == TSS_OK && nf == 0);
        CHECK(tss_layout_from_assignment(inst, a2, nv + 1, decoded, 64, &n_dec) == TSS_OK && n_dec == 3);
        {   /* the completion step on its own: platform and terrain-layer variables from the layout, the totalizer left open */
            int32_t n_base = 0, conflict = 0, n_fals = -1;
            tss_encoding_sizes(inst, &n_base, NULL, NULL, NULL);
            CHECK(tss_layout_to_assignment(e, inst, decoded, n_dec, a2) == TSS_OK);
            for (int v = n_base + 1; v <= nv; v++) a2[v] = 2;
            CHECK(tss_cnf_complete(e, c2, a2, &conflict, &n_fals) == TSS_OK && conflict == -1 && n_fals == 0);
            for (int v = 1; v <= nv; v++) CHECK(a2[v] == 0 || a2[v] == 1);
        }
        CHECK(tss_witness_for_cnf(e, c2, inst, decoded, 2, a2) == TSS_UNKNOWN);                               /* one support short: not a model */
        {   /* the loop's next question, "at most 2 platforms" (main.rs:346): below the certified lower bound of 3 -> UNSAT, no exact solver */
            const int32_t card2[3] = {1, 1, 2};
            tss_encoding* enc2 = NULL;
            CHECK(tss_encoding_create(grid, w, h, defs, 1, &enc2) == TSS_OK);
            int32_t nv3 = 0, nc3 = 0;
            int64_t nl3 = 0;
            CHECK(tss_encoding_with_limits(enc2, card2, 1, NULL, 0, 0, 0, &nv3, &nc3, &nl3, NULL, NULL) == TSS_OK);
            int32_t* l3 = malloc(sizeof(int32_t) * (size_t)(nl3 + 1));
            uint32_t* o3 = malloc(sizeof(uint32_t) * (size_t)(nc3 + 1));
            uint8_t* a3 = malloc((size_t)nv3 + 1);
            CHECK(l3 && o3 && a3);
            CHECK(tss_encoding_with_limits(enc2, card2, 1, NULL, 0, 0, 0, &nv3, &nc3, &nl3, l3, o3) == TSS_OK);
            tss_cnf* c3 = NULL;
            tss_encoding* inst3 = NULL;
            tss_instance_info info3;
            CHECK(tss_cnf_upload(e, l3, o3, nc3, nv3, &c3) == TSS_OK);
            CHECK(tss_instance_find(l3, o3, nc3, nv3, &inst3, &info3, NULL, 0) == TSS_SAT && info3.card_limit_1x1 == 2);
            CHECK(tss_engine_certified_unsat(e, 0) == TSS_OK);
            CHECK(tss_solve_instance(e, c3, inst3, &info3, NULL, 5, 2000, a3) == TSS_UNKNOWN);   /* a search alone proves nothing */
            CHECK(tss_engine_certified_unsat(e, 1) == TSS_OK);
            CHECK(tss_solve_instance(e, c3, inst3, &info3, NULL, 5, 2000, a3) == TSS_UNSAT);
            CHECK(tss_solve_instance(e, c3, inst3, &info3, NULL, 6, 2000, a3) == TSS_UNSAT);     /* from the cached bound */
            tss_encoding_destroy(inst3);
            tss_encoding_destroy(enc2);
            tss_cnf_destroy(c3);
            free(l3); free(o3); free(a3);
        }
        l2[0] = -l2[0];                                          /* other clauses: not a recorded instance */
        CHECK(tss_instance_find(l2, o2, nc, nv, &enc1, &info, NULL, 0) == TSS_UNKNOWN && enc1 == NULL);
        tss_encoding_destroy(inst);
        tss_cnf_destroy(c2);
        free(l2); free(o2); free(a2);
    }
    CHECK(tss_debug_smem_violations() == -1);                    /* the shipped build carries no bounds checks */

    tss_cnf_destroy(cnf);
    tss_encoding_destroy(enc);
    free(lits); free(offsets); free(assignment);
    tss_engine_destroy(e);
    printf("c_abi_smoke ok\n");
    return 0;
}
