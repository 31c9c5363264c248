//! `GpuSeeded<S>` — the drop-in the reference's drivers use in place of `GlucoseSimp`
//! (`crates/repl/src/main.rs:17,295`, `crates/gui/src/main.rs:2,26`): a type that is
//! `Solve + Interrupt + SolveStats + Default + Send`, owns an exact solver `S` and a `tss_engine`, answers
//! `solve()` from the GPU when it can and from `S` otherwise.  UNSAT comes from `S`, or from a certified lower bound of the
//! instance (`tss_solve_instance` returning `TSS_UNSAT`), never from a search that merely found nothing.
//!
//! The drivers change ONE line (`use tss::GpuSeeded as GlucoseSimp;`): the solver is handed a bare `Cnf`
//! (`crates/repl/src/solver_runner.rs:8-20`) and finds terrain, platform set, variable map and limits again through the
//! instance registry that the lib crate's `Encoding::with_limits` fills (`tss_instance_find`, see `lib-patch/encoder_ffi.rs`).
//!
//! SOURCE ONLY — not compiled in this repository (no Rust toolchain in the build image); the rustsat 0.7 trait
//! surface is written from memory and must be checked against the pinned crate.  The call sequence below is EXACTLY the
//! one `tools/tss_repl.cpp` performs (tss_cnf_upload -> tss_instance_find -> tss_solve_instance), which the GPU test suite
//! runs on test/ex1-3 (tests/test_gpu.py::test_cpp_repl_driver_*); `tss_solve_instance` itself is host C++ over the public
//! C ABI (csrc/instance.cpp), so the recipe is exercised without a Rust toolchain.
//!
//! Without a usable CUDA device `Default` logs a warning and the value is "exact solver only" — acceptable degradation of
//! the wrapper (the drivers must keep working on a laptop); libtss itself has no CPU fallback.
use std::ffi::CStr;
use std::ptr;
use std::sync::Arc;

use anyhow::Result;
use rustsat::instances::Cnf;
use rustsat::solvers::{Interrupt, InterruptSolver, Solve, SolveStats, SolverResult, SolverStats};
use rustsat::types::{Assignment, Lit, TernaryVal};
use tss_sys as ffi;

/// Owns the engine.  Shared (`Arc`) between the solver and every interrupter it handed out, so an interrupter that fires
/// from another thread (main.rs:298-323) can never outlive the engine it points at: the engine is destroyed when the last
/// holder drops.  `tss_interrupt` is the one entry point documented as callable concurrently with a running solve.
struct EngineHandle(*mut ffi::tss_engine);
unsafe impl Send for EngineHandle {}
unsafe impl Sync for EngineHandle {}
impl Drop for EngineHandle {
    fn drop(&mut self) {
        if !self.0.is_null() {
            unsafe { ffi::tss_engine_destroy(self.0) };
        }
    }
}

pub struct GpuSeeded<S = rustsat_glucose::simp::Glucose> {
    inner: S,
    engine: Option<Arc<EngineHandle>>,
    cnf: *mut ffi::tss_cnf,              // the clauses as uploaded in add_cnf (kernel (c) verifies witnesses against them)
    enc: *mut ffi::tss_encoding,         // the instance those clauses belong to (tss_instance_find), or null
    info: ffi::tss_instance_info,
    weights: Vec<i32>,                   // (def_w, def_h, weight) records of PlatformLimits.weights
    n_vars: i32,
    witness: Option<Assignment>,
    give_up: i64,                        // steps per chain before the exact solver takes over
    seed: u64,
}

// The engine is only ever driven from the thread that owns the solver (tokio moves the solver into the blocking task and
// back, solver_runner.rs:15-17); tss_interrupt is the one call made from elsewhere.
unsafe impl<S: Send> Send for GpuSeeded<S> {}

fn last_error(e: *const ffi::tss_engine) -> String {
    unsafe { CStr::from_ptr(ffi::tss_last_error(e)).to_string_lossy().into_owned() }
}

impl<S: Default> Default for GpuSeeded<S> {
    fn default() -> Self {
        let mut raw = ptr::null_mut();
        let engine = if unsafe { ffi::tss_engine_create(-1, &mut raw) } == ffi::TSS_OK {
            Some(Arc::new(EngineHandle(raw)))
        } else {
            log::warn!("tss: no usable CUDA device, solving with the exact solver only");
            None
        };
        GpuSeeded {
            inner: S::default(),
            engine,
            cnf: ptr::null_mut(),
            enc: ptr::null_mut(),
            info: ffi::tss_instance_info::default(),
            weights: vec![0; 3 * 16],
            n_vars: 0,
            witness: None,
            give_up: 1024,
            seed: 0,
        }
    }
}

impl<S> Drop for GpuSeeded<S> {
    fn drop(&mut self) {
        unsafe {
            if !self.cnf.is_null() {
                ffi::tss_cnf_destroy(self.cnf);
            }
            if !self.enc.is_null() {
                ffi::tss_encoding_destroy(self.enc);
            }
        }
        // the engine goes when the last Arc (this one or an interrupter's) is dropped
    }
}

impl<S: Solve> Solve for GpuSeeded<S> {
    fn signature(&self) -> &'static str {
        "tss GpuSeeded (B200 upper-bound engine + exact solver)"
    }

    fn add_cnf(&mut self, cnf: Cnf) -> Result<()> {
        if let Some(engine) = &self.engine {
            unsafe {
                // (a second add_cnf on the same solver replaces the uploaded clauses; the drivers create one solver per iteration)
                if !self.cnf.is_null() {
                    ffi::tss_cnf_destroy(self.cnf);
                    self.cnf = ptr::null_mut();
                }
                if !self.enc.is_null() {
                    ffi::tss_encoding_destroy(self.enc);
                    self.enc = ptr::null_mut();
                }
            }
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
            let n_clauses = (offsets.len() - 1) as i32;
            let rc = unsafe { ffi::tss_cnf_upload(engine.0, lits.as_ptr(), offsets.as_ptr(), n_clauses, n_vars, &mut self.cnf) };
            if rc != ffi::TSS_OK {
                log::warn!("tss_cnf_upload: {}", last_error(engine.0)); // logged, never fatal (crates/gui/src/app.rs:160-173)
                self.cnf = ptr::null_mut();
            } else {
                // which instance is this?  (recorded by Encoding::with_limits, lib-patch/encoder_ffi.rs)
                let rc = unsafe {
                    ffi::tss_instance_find(lits.as_ptr(), offsets.as_ptr(), n_clauses, n_vars, &mut self.enc, &mut self.info, self.weights.as_mut_ptr(), 16)
                };
                if rc != ffi::TSS_SAT {
                    self.enc = ptr::null_mut(); // not one of ours: the exact solver handles it alone
                }
            }
        }
        self.inner.add_cnf(cnf)
    }

    fn add_clause_ref<C>(&mut self, clause: &C) -> Result<()>
    where
        C: AsRef<rustsat::types::Cl> + ?Sized,
    {
        if !self.enc.is_null() {
            unsafe { ffi::tss_encoding_destroy(self.enc) }; // clauses beyond the recorded instance: the GPU's view would be stale
            self.enc = ptr::null_mut();
        }
        self.inner.add_clause_ref(clause)
    }

    fn solve(&mut self) -> Result<SolverResult> {
        self.witness = None;
        if let (Some(engine), false, false) = (&self.engine, self.cnf.is_null(), self.enc.is_null()) {
            let mut a = vec![2u8; self.n_vars as usize + 1]; // 0 = false, 1 = true, 2 = unassigned; slot 0 unused
            unsafe { ffi::tss_clear_interrupt(engine.0) };
            // ONE SAT-like call within the instance's limit, giving up after `give_up` steps per chain; the witness is built
            // with tss_layout_to_assignment (platform + terrain-layer variables), completed by unit propagation (totalizer
            // auxiliaries) and checked against every uploaded clause before it is trusted — all inside tss_solve_instance
            let rc = unsafe { ffi::tss_solve_instance(engine.0, self.cnf, self.enc, &self.info, self.weights.as_ptr(), self.seed, self.give_up, a.as_mut_ptr()) };
            self.seed = self.seed.wrapping_add(1);
            if rc == ffi::TSS_SAT {
                let mut st = ffi::tss_stats::default();
                unsafe { ffi::tss_get_stats(engine.0, &mut st) };
                self.give_up = (32 * st.last_solve_steps).max(1024);
                let vals: Vec<TernaryVal> = a[1..].iter().map(|&b| if b == 1 { TernaryVal::True } else { TernaryVal::False }).collect();
                self.witness = Some(Assignment::from(vals));
                return Ok(SolverResult::Sat);
            } else if rc == ffi::TSS_UNSAT {
                // the limit lies below a CERTIFIED lower bound of the instance (packing / fractional LP, include/tss.h): the
                // drivers' loops end here ("No solution found for the current constraints", main.rs:331-334) without the exact solver
                return Ok(SolverResult::Unsat);
            } else if rc < 0 {
                log::warn!("tss_solve_instance: {}", last_error(engine.0));
            }
        }
        self.inner.solve() // the exact solver: every UNSAT answer by SEARCH comes from here
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

/// Interrupts both sides.  Holds its own reference to the engine, so it stays valid whatever happens to the solver.
pub struct Both<I> {
    engine: Option<Arc<EngineHandle>>,
    inner: I,
}
impl<I: InterruptSolver> InterruptSolver for Both<I> {
    fn interrupt(&mut self) {
        if let Some(e) = &self.engine {
            unsafe { ffi::tss_interrupt(e.0) };
        }
        self.inner.interrupt();
    }
}

impl<S: Interrupt> Interrupt for GpuSeeded<S> {
    type Interrupter = Both<S::Interrupter>;
    fn interrupter(&mut self) -> Self::Interrupter {
        Both { engine: self.engine.clone(), inner: self.inner.interrupter() }
    }
}

impl<S: SolveStats> SolveStats for GpuSeeded<S> {
    fn stats(&self) -> SolverStats {
        self.inner.stats() // engine counters are available through tss_get_stats
    }
}
