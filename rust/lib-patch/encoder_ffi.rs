//! The lib crate's side of the boundary (SOURCE ONLY, see ../README.md): `Encoding::encode` / `Encoding::with_limits`
//! (src/encoder.rs:435,619) backed by libtss's host-side encoder.  Same public API, same clause families (the oracle's
//! restatement and libtss agree clause for clause, tests/test_host.py), deterministic variable numbering instead of the
//! reference's HashMap order (src/encoder.rs:191-194).  `with_limits` is the call that RECORDS the instance — encoding,
//! limits and the lowered clauses — in libtss's registry, which is how a solver that later receives only the `Cnf`
//! (crates/repl/src/solver_runner.rs:8-20) finds terrain, platform set and limits again without any change to the drivers.
use rustsat::instances::{Cnf, SatInstance};
use rustsat::types::{Clause, Lit};
use tss_sys as ffi;

pub struct Encoding {
    handle: *mut ffi::tss_encoding,
    vars: EncodingVars, // built from tss_encoding_var_maps: plat_var[tile * K + k], terr_var[tile * 4 + layer]
}

impl Encoding {
    pub fn encode(platform_defs: &[PlatformDef], terrain: &WorldGrid) -> Self {
        let defs: Vec<ffi::tss_dims> = platform_defs.iter().map(|d| ffi::tss_dims { w: d.dims().width as i32, h: d.dims().height as i32 }).collect();
        let grid: Vec<u8> = terrain.iter().map(|&b| b as u8).collect(); // Grid<bool>.data, row-major (src/math/grid.rs:66-68)
        let mut handle = std::ptr::null_mut();
        let rc = unsafe { ffi::tss_encoding_create(grid.as_ptr(), terrain.dims().width as i32, terrain.dims().height as i32, defs.as_ptr(), defs.len() as i32, &mut handle) };
        assert_eq!(rc, ffi::TSS_OK, "platform set without 1x1 or empty grid (the reference unwraps, src/encoder.rs:564-566)");
        Encoding { handle, vars: EncodingVars::from_maps(handle, terrain.dims(), platform_defs) }
    }

    pub fn vars(&self) -> &EncodingVars {
        &self.vars
    }

    /// src/encoder.rs:619-667 + `into_cnf()`: the returned instance already holds the lowered cardinality / PB clauses, so the
    /// drivers' `instance.into_cnf()` (crates/repl/src/main.rs:293) is the identity on it.
    pub fn with_limits(&self, limits: &PlatformLimits) -> SatInstance {
        let card: Vec<i32> = limits.card_limits.iter().flat_map(|(d, n)| [d.dims().width as i32, d.dims().height as i32, *n as i32]).collect();
        let wts: Vec<i32> = limits.weights.iter().flat_map(|(d, w)| [d.dims().width as i32, d.dims().height as i32, *w as i32]).collect();
        let (mut n_vars, mut n_clauses, mut n_lits) = (0i32, 0i32, 0i64);
        let wl = limits.weight_limit;
        unsafe {
            ffi::tss_encoding_with_limits(self.handle, card.as_ptr(), (card.len() / 3) as i32, wts.as_ptr(), (wts.len() / 3) as i32, wl.is_some() as i32,
                                          wl.unwrap_or(0) as i64, &mut n_vars, &mut n_clauses, &mut n_lits, std::ptr::null_mut(), std::ptr::null_mut());
        }
        let (mut lits, mut offsets) = (vec![0i32; n_lits as usize], vec![0u32; n_clauses as usize + 1]);
        unsafe {
            // the call with buffers is the one that records (encoding, limits, clauses) for tss_instance_find
            ffi::tss_encoding_with_limits(self.handle, card.as_ptr(), (card.len() / 3) as i32, wts.as_ptr(), (wts.len() / 3) as i32, wl.is_some() as i32,
                                          wl.unwrap_or(0) as i64, &mut n_vars, &mut n_clauses, &mut n_lits, lits.as_mut_ptr(), offsets.as_mut_ptr());
        }
        let mut cnf = Cnf::new();
        for c in 0..n_clauses as usize {
            let clause: Clause = lits[offsets[c] as usize..offsets[c + 1] as usize]
                .iter()
                .map(|&l| if l > 0 { Lit::positive(l as u32 - 1) } else { Lit::negative((-l) as u32 - 1) })
                .collect();
            cnf.add_clause(clause);
        }
        SatInstance::from(cnf) // clause order and variable indices are preserved: add_cnf sees exactly the recorded CSR
    }
}

impl Drop for Encoding {
    fn drop(&mut self) {
        unsafe { ffi::tss_encoding_destroy(self.handle) }; // the registry keeps its own reference to the instance
    }
}
