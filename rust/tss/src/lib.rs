//! `GpuSeeded<S>` — the drop-in the reference's drivers can use in place of `GlucoseSimp`
//! (`crates/repl/src/main.rs:17,295`, `crates/gui/src/main.rs:2,26`): a type that is
//! `Solve + Interrupt + SolveStats + Default + Send`, owns an exact solver `S` and a `tss_engine`, answers
//! `solve()` from the GPU when it can and from `S` otherwise.  `S` stays the only prover of UNSAT.
//!
//! SOURCE ONLY — not compiled in this repository (no Rust toolchain in the build image); the rustsat 0.7 trait
//! surface is written from memory and must be checked against the pinned crate.  The same call sequence is
//! exercised, tested and timed from Python (`timberborn_support_solver_b200/api.py`: `GpuBoundSolver`,
//! `solver_loop`) and from C (`tests/c_abi_smoke.c`).
use std::ffi::CStr;
use std::ptr;
use std::sync::atomic::{AtomicPtr, Ordering};
use std::sync::Arc;

use anyhow::{anyhow, Result};
use rustsat::instances::Cnf;
use rustsat::solvers::{Interrupt, InterruptSolver, Solve, SolveStats, SolverResult, SolverStats};
use rustsat::types::{Assignment, Clause, Lit, TernaryVal};
use tss_sys as ffi;

/// What a bare CNF does not carry: the terrain and the platform set.  Handed over once, next to
/// `Encoding::encode` (`crates/repl/src/main.rs:254`, `crates/gui/src/app.rs:180-184`).
#[derive(Clone, Default)]
pub struct Terrain {
    pub grid: Vec<u8>, // Grid<bool>.data, row-major x + y * width (src/math/grid.rs:66-68)
    pub w: i32,
    pub h: i32,
    pub defs: Vec<ffi::tss_dims>,   // PlatformDef dims in the order given to the encoder
    pub plat_var_1x1: Vec<u32>,     // variable of P_1x1 at every tile, row-major (EncodingVars, src/encoder.rs:184-206)
    pub terr_vars: Vec<[u32; 4]>,   // T_0..T_3 per tile, 0 = absent
}

pub struct GpuSeeded<S = rustsat_glucose::simp::Glucose> {
    inner: S,
    engine: *mut ffi::tss_engine,
    shared: Arc<AtomicPtr<ffi::tss_engine>>, // for the interrupter (another thread, main.rs:298-323)
    cnf: *mut ffi::tss_cnf,
    n_vars: i32,
    terrain: Terrain,
    bound: Option<i32>,            // n of "sum P_1x1 <= n" (Encoding::with_limits, src/encoder.rs:643-652)
    witness: Option<Assignment>,
    give_up: i64,                  // steps per chain before the exact solver takes over
    seed: u64,
}

// The engine is only ever driven from the thread that owns the solver; tss_interrupt is the one call made from
// elsewhere and is documented as thread safe (include/tss.h).
unsafe impl<S: Send> Send for GpuSeeded<S> {}

fn last_error(e: *const ffi::tss_engine) -> String {
    unsafe { CStr::from_ptr(ffi::tss_last_error(e)).to_string_lossy().into_owned() }
}

impl<S: Default> Default for GpuSeeded<S> {
    fn default() -> Self {
        let mut engine = ptr::null_mut();
        // no usable CUDA device: engine stays null and every solve() goes straight to the exact solver (logged, never a panic —
        // the drivers log solver errors and carry on, crates/gui/src/app.rs:160-173)
        if unsafe { ffi::tss_engine_create(-1, &mut engine) } != ffi::TSS_OK {
            log::warn!("tss: no CUDA device, solving on the CPU only");
            engine = ptr::null_mut();
        }
        GpuSeeded {
            inner: S::default(),
            engine,
            shared: Arc::new(AtomicPtr::new(engine)),
            cnf: ptr::null_mut(),
            n_vars: 0,
            terrain: Terrain::default(),
            bound: None,
            witness: None,
            give_up: 1024,
            seed: 0,
        }
    }
}

impl<S> Drop for GpuSeeded<S> {
    fn drop(&mut self) {
        self.shared.store(ptr::null_mut(), Ordering::SeqCst);
        unsafe {
            if !self.cnf.is_null() {
                ffi::tss_cnf_destroy(self.cnf);
            }
            if !self.engine.is_null() {
                ffi::tss_engine_destroy(self.engine);
            }
        }
    }
}

impl<S> GpuSeeded<S> {
    /// Terrain, platform set, variable map and the current cardinality bound of the instance about to be added.
    pub fn set_terrain(&mut self, terrain: Terrain, bound: Option<usize>) {
        self.terrain = terrain;
        self.bound = bound.map(|b| b as i32);
    }

    /// GPU layout -> full assignment: platform variables from the layout, terrain-layer variables from validate()'s
    /// dilation rounds, auxiliary cardinality variables by unit propagation on the uploaded CNF (kernel (c)); the
    /// result is checked against every clause the exact solver received before it is trusted.
    fn verified_assignment(&mut self, plats: &[ffi::tss_platform]) -> Result<Option<Assignment>> {
        let n = (self.n_vars + 1) as usize;
        let mut a = vec![2u8; n]; // 0 = false, 1 = true, 2 = unassigned
        for v in &self.terrain.plat_var_1x1 {
            a[*v as usize + 1] = 0;
        }
        for p in plats {
            // 1x1 supports only in this sketch; larger platforms set P_dims at the anchor and, through the encoder's DAG
            // implications (src/encoder.rs:450-458), every smaller dims variable: see tss_layout_to_assignment
            let tile = (p.y * self.terrain.w + p.x) as usize;
            a[self.terrain.plat_var_1x1[tile] as usize + 1] = 1;
        }
        let (mut conflict, mut rounds) = (-1i32, 0i32);
        let rc = unsafe { ffi::tss_cnf_propagate(self.engine, self.cnf, a.as_mut_ptr(), 1, &mut conflict, &mut rounds) };
        if rc < 0 {
            return Err(anyhow!("tss_cnf_propagate: {}", last_error(self.engine)));
        }
        if conflict >= 0 {
            return Ok(None);
        }
        for v in a.iter_mut() {
            if *v == 2 {
                *v = 0;
            }
        }
        let (mut n_false, mut first) = (0i32, -1i32);
        let rc = unsafe { ffi::tss_cnf_check(self.engine, self.cnf, a.as_ptr(), 1, &mut n_false, &mut first) };
        if rc < 0 || n_false != 0 {
            return Ok(None);
        }
        let vals: Vec<TernaryVal> = a[1..].iter().map(|&b| if b == 1 { TernaryVal::True } else { TernaryVal::False }).collect();
        Ok(Some(Assignment::from(vals)))
    }
}

impl<S: Solve> Solve for GpuSeeded<S> {
    fn signature(&self) -> &'static str {
        "tss GpuSeeded (B200 upper-bound engine + exact solver)"
    }

    fn add_cnf(&mut self, cnf: Cnf) -> Result<()> {
        if !self.engine.is_null() {
            // CSR with DIMACS-signed literals: Lit(idx, negated) -> +/-(idx + 1)
            let (mut lits, mut offsets, mut n_vars) = (Vec::<i32>::new(), vec![0u32], 0i32);
            for clause in cnf.iter() {
                for l in clause.iter() {
                    let v = l.vidx32() as i32 + 1;
                    n_vars = n_vars.max(v);
                    lits.push(if l.is_neg() { -v } else { v });
                }
                offsets.push(lits.len() as u32);
            }
            self.n_vars = n_vars;
            let rc = unsafe {
                ffi::tss_cnf_upload(self.engine, lits.as_ptr(), offsets.as_ptr(), (offsets.len() - 1) as i32, n_vars, &mut self.cnf)
            };
            if rc != ffi::TSS_OK {
                log::warn!("tss_cnf_upload: {}", last_error(self.engine));
                self.cnf = ptr::null_mut();
            }
        }
        self.inner.add_cnf(cnf)
    }

    fn add_clause_ref<C>(&mut self, clause: &C) -> Result<()>
    where
        C: AsRef<rustsat::types::Cl> + ?Sized,
    {
        self.inner.add_clause_ref(clause)
    }

    fn solve(&mut self) -> Result<SolverResult> {
        self.witness = None;
        if !self.engine.is_null() && !self.cnf.is_null() && !self.terrain.grid.is_empty() {
            let t = &self.terrain;
            let mut plats = vec![ffi::tss_platform::default(); t.grid.len() + 1];
            let mut n = 0i32;
            // ONE SAT-like call: the first layout within the bound, give up after `give_up` steps per chain
            let rc = unsafe {
                ffi::tss_solve_upper_bound(
                    self.engine, t.grid.as_ptr(), t.w, t.h, t.defs.as_ptr(), t.defs.len() as i32, self.bound.unwrap_or(-1), self.seed, 0,
                    -self.give_up, plats.as_mut_ptr(), plats.len() as i32, &mut n,
                )
            };
            self.seed = self.seed.wrapping_add(1);
            if rc == ffi::TSS_SAT {
                let mut st = ffi::tss_stats::default();
                unsafe { ffi::tss_get_stats(self.engine, &mut st) };
                self.give_up = (32 * st.last_solve_steps).max(1024);
                if let Some(a) = self.verified_assignment(&plats[..n as usize])? {
                    self.witness = Some(a);
                    return Ok(SolverResult::Sat);
                }
            } else if rc < 0 {
                log::warn!("tss_solve_upper_bound: {}", last_error(self.engine)); // logged, never fatal (app.rs:160-173)
            }
        }
        self.inner.solve() // the exact solver: every UNSAT answer comes from here
    }

    fn lit_val(&self, lit: Lit) -> Result<TernaryVal> {
        match &self.witness {
            Some(a) => Ok(a.lit_value(lit)),
            None => self.inner.lit_val(lit),
        }
    }

    fn full_solution(&self) -> Result<Assignment> {
        match &self.witness {
            Some(a) => Ok(a.clone()),
            None => self.inner.full_solution(),
        }
    }
}

/// Interrupts both sides: `tss_interrupt` is safe from any thread while a solve runs (include/tss.h).
pub struct Both<I> {
    engine: Arc<AtomicPtr<ffi::tss_engine>>,
    inner: I,
}
impl<I: InterruptSolver> InterruptSolver for Both<I> {
    fn interrupt(&mut self) {
        let e = self.engine.load(Ordering::SeqCst);
        if !e.is_null() {
            unsafe { ffi::tss_interrupt(e) };
        }
        self.inner.interrupt();
    }
}
unsafe impl<I: Send> Send for Both<I> {}

impl<S: Interrupt> Interrupt for GpuSeeded<S> {
    type Interrupter = Both<S::Interrupter>;
    fn interrupter(&mut self) -> Self::Interrupter {
        Both { engine: self.shared.clone(), inner: self.inner.interrupter() }
    }
}

impl<S: SolveStats> SolveStats for GpuSeeded<S> {
    fn stats(&self) -> SolverStats {
        self.inner.stats() // engine counters are available through tss_get_stats
    }
}

#[allow(unused)]
fn _clause_type_is_used(_: Clause) {}
