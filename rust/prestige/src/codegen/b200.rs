//! prestige::codegen::b200 -- sibling of simple_cpu (prestige/src/codegen/simple_cpu.rs:3).
//! Where generate_simple_cpu(&FusedEquations) returns the all-pairs loop as text, this back-end RUNS the
//! fused set on the GPU through prestige_b200_sys.  UNCOMPILED in this repository (no Rust toolchain).
//!
//! Needs `pub names: Vec<String>` on FusedEquations (see ../../fuse.patch.md): the reference's fuse()
//! (fuse.rs:14-40) keeps bodies but drops equation names, and a back-end that dispatches to hand-written
//! kernels selects them by name.
use crate::equations::fuse::FusedEquations;
use prestige_b200_sys as sys;
use std::ffi::{CStr, CString};

#[derive(Debug)]
pub struct B200Error {
    pub status: i32,
    pub message: String,
}

/// Owns one pst_ctx (one GPU).  Host slices are borrowed per call, exactly like `&mut [f64]` / `&[f64]` in
/// prestige/src/lib.rs:8; arrays are identified by the same names the EquationIR carries.
pub struct B200Context {
    raw: *mut sys::pst_ctx,
    n: usize,
}

const KERNELS: [&str; 7] = ["eq1", "tait_eos", "wall_pressure", "continuity", "momentum", "dem_contact", "body_reduce"];

impl B200Context {
    pub fn new(cfg: &sys::pst_config) -> Result<Self, B200Error> {
        let mut raw = std::ptr::null_mut();
        let st = unsafe { sys::pst_create(cfg, &mut raw) };
        if st != sys::PST_OK {
            let msg = unsafe { CStr::from_ptr(sys::pst_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(B200Error { status: st, message: msg });
        }
        Ok(Self { raw, n: 0 })
    }

    fn check(&self, st: i32) -> Result<(), B200Error> {
        if st == sys::PST_OK {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(sys::pst_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(B200Error { status: st, message: msg })
    }

    pub fn set_count(&mut self, n: usize) -> Result<(), B200Error> {
        self.n = n;
        self.check(unsafe { sys::pst_set_count(self.raw, n as u64) })
    }

    pub fn set_param(&mut self, name: &str, v: f64) -> Result<(), B200Error> {
        let c = CString::new(name).unwrap();
        self.check(unsafe { sys::pst_set_param(self.raw, c.as_ptr(), v) })
    }

    pub fn array_create(&mut self, name: &str) -> Result<(), B200Error> {
        let c = CString::new(name).unwrap();
        self.check(unsafe { sys::pst_array_create(self.raw, c.as_ptr(), sys::PST_REAL, sys::PST_ARRAY_PERSISTENT) })
    }

    pub fn upload(&mut self, name: &str, host: &[f64]) -> Result<(), B200Error> {
        let c = CString::new(name).unwrap();
        self.check(unsafe { sys::pst_upload(self.raw, c.as_ptr(), host.as_ptr().cast(), host.len()) })
    }

    pub fn download(&mut self, name: &str, host: &mut [f64]) -> Result<(), B200Error> {
        let c = CString::new(name).unwrap();
        self.check(unsafe { sys::pst_download(self.raw, c.as_ptr(), host.as_mut_ptr().cast(), host.len()) })
    }

    pub fn build_neighbours(&mut self) -> Result<(), B200Error> {
        self.check(unsafe { sys::pst_build_neighbours(self.raw) })
    }

    /// The counterpart of generate_simple_cpu: execute the fused bodies (one fused pair kernel).
    pub fn run(&mut self, ir: &FusedEquations) -> Result<(), B200Error> {
        let plan = generate_b200(ir).map_err(|m| B200Error { status: sys::PST_EINVAL, message: m })?;
        let c: Vec<CString> = plan.iter().map(|s| CString::new(s.as_str()).unwrap()).collect();
        let p: Vec<*const std::os::raw::c_char> = c.iter().map(|s| s.as_ptr()).collect();
        self.check(unsafe { sys::pst_apply(self.raw, p.as_ptr(), p.len() as i32) })
    }
}

impl Drop for B200Context {
    fn drop(&mut self) {
        unsafe { sys::pst_destroy(self.raw) }
    }
}

/// The launch plan for a fused set: the equation names pst_apply receives, in body order.
pub fn generate_b200(ir: &FusedEquations) -> Result<Vec<String>, String> {
    if ir.names.is_empty() {
        return Err("FusedEquations.names is empty".into());
    }
    for n in &ir.names {
        if !KERNELS.contains(&n.as_str()) {
            return Err(format!("no hand-written kernel for equation '{n}'"));
        }
    }
    Ok(ir.names.clone())
}
