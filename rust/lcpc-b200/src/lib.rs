//! Safe wrapper: `GpuLigeroEncoding<F>` implements `lcpc_2d::LcEncoding`, so
//! `LcCommit::prove` and `LcEvalProof::verify` from lcpc-2d work unchanged, and
//! `commit_gpu` replaces lines 665-697 of lcpc-2d/src/lib.rs (pad, encode, hash, Merkle)
//! while returning a genuine `LcCommit` with host `Vec`s.
//!
//! Written against include/lcpc_b200.h ABI version 1.  This crate has not been compiled:
//! the build image has no Rust toolchain.
use std::ffi::CStr;
use std::marker::PhantomData;
use std::ptr;

use blake3::Hasher as Blake3;
use digest::Output;
use ff::PrimeField;
use fffft::{FFTError, FieldFFT};
use lcpc_2d::{def_labels, n_degree_tests, FieldHash, LcCommit, LcEncoding, ProverError, SizedField};
use lcpc_b200_sys as sys;

/// Fields whose in-memory form is `[u64; LIMBS]` Montgomery limbs (every ff_derive field).
pub trait GpuField: PrimeField + FieldFFT + FieldHash + SizedField {
    const FIELD_ID: i32;
    const LIMBS: usize;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::lcpc_last_error()).to_string_lossy().into_owned() }
}

struct Plan(*mut sys::lcpc_plan, *mut sys::lcpc_ctx);
unsafe impl Send for Plan {}
unsafe impl Sync for Plan {} // handles are internally locked (header, "Conventions")
impl Drop for Plan {
    fn drop(&mut self) {
        unsafe {
            sys::lcpc_plan_destroy(self.0);
            sys::lcpc_ctx_destroy(self.1);
        }
    }
}

#[derive(Clone)]
pub struct GpuLigeroEncoding<F> {
    n_per_row: usize,
    n_cols: usize,
    plan: std::sync::Arc<Plan>,
    _p: PhantomData<F>,
}

impl<F> std::fmt::Debug for GpuLigeroEncoding<F> {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "GpuLigeroEncoding({} -> {})", self.n_per_row, self.n_cols)
    }
}

impl<F: GpuField> GpuLigeroEncoding<F> {
    /// LigeroEncodingRho::new_from_dims (lcpc-ligero-pc/src/lib.rs:138-148)
    pub fn new_from_dims(n_per_row: usize, n_cols: usize, device: i32) -> Self {
        assert!(n_per_row < n_cols && n_cols.is_power_of_two());
        let log_len = n_cols.trailing_zeros();
        // the n_cols-th root fffft would use: ROOT_OF_UNITY^(2^(S - log_len))
        let mut root = F::ROOT_OF_UNITY;
        for _ in 0..(<F as FieldFFT>::S - log_len) {
            root *= root;
        }
        let mut ctx = ptr::null_mut();
        let mut plan = ptr::null_mut();
        unsafe {
            assert_eq!(sys::lcpc_ctx_create(device, &mut ctx), sys::LCPC_OK, "{}", last_error());
            let rc = sys::lcpc_plan_ligero(ctx, F::FIELD_ID, n_per_row, n_cols, &root as *const F as *const u64, &mut plan);
            assert_eq!(rc, sys::LCPC_OK, "{}", last_error());
        }
        Self { n_per_row, n_cols, plan: std::sync::Arc::new(Plan(plan, ctx)), _p: PhantomData }
    }
}

impl<F: GpuField> LcEncoding for GpuLigeroEncoding<F> {
    type F = F;
    type Err = FFTError;
    def_labels!(ligero_pc);

    /// one row in place; lcpc-2d calls this from rayon workers, the plan serialises them
    fn encode<T: AsMut<[F]>>(&self, mut inp: T) -> Result<(), FFTError> {
        let row = inp.as_mut();
        if row.len() != self.n_cols {
            return Err(FFTError::TooBig); // closest fffft error for a length mismatch
        }
        let rc = unsafe { sys::lcpc_encode_rows(self.plan.0, row.as_mut_ptr() as *mut u64, 1) };
        if rc == sys::LCPC_OK { Ok(()) } else { Err(FFTError::TooBig) }
    }
    fn get_dims(&self, len: usize) -> (usize, usize, usize) {
        ((len + self.n_per_row - 1) / self.n_per_row, self.n_per_row, self.n_cols)
    }
    fn dims_ok(&self, n_per_row: usize, n_cols: usize) -> bool {
        n_per_row < n_cols && n_cols.is_power_of_two() && n_per_row == self.n_per_row && n_cols == self.n_cols
    }
    fn get_n_col_opens(&self) -> usize {
        // rho = 1/2: ceil(-128 / log2(3/4)) (lcpc-ligero-pc/src/lib.rs:61-64)
        (-(128f64) / ((1f64 + 0.5f64) / 2f64).log2()).ceil() as usize
    }
    fn get_n_degree_tests(&self) -> usize {
        n_degree_tests(128, self.n_cols, F::FLOG2 as usize)
    }
}

/// Drop-in for `LcCommit::<Blake3, _>::commit(coeffs, enc)` (lcpc-2d/src/lib.rs:314).
pub fn commit_gpu<F: GpuField>(coeffs_in: &[F], enc: &GpuLigeroEncoding<F>)
    -> Result<LcCommit<Blake3, GpuLigeroEncoding<F>>, ProverError<FFTError>>
{
    let (n_rows, n_per_row, n_cols) = enc.get_dims(coeffs_in.len());
    assert!(n_rows * n_per_row >= coeffs_in.len());
    assert!((n_rows - 1) * n_per_row < coeffs_in.len());
    let np2 = n_cols.checked_next_power_of_two().ok_or(ProverError::TooBig)?;
    let mut coeffs = vec![F::ZERO; n_rows * n_per_row];
    let mut comm = vec![F::ZERO; n_rows * n_cols];
    let mut hashes = vec![<Output<Blake3> as Default>::default(); 2 * np2 - 1];
    let rc = unsafe {
        sys::lcpc_commit_host(enc.plan.0, coeffs_in.as_ptr() as *const u64, coeffs_in.len(),
                              coeffs.as_mut_ptr() as *mut u64, comm.as_mut_ptr() as *mut u64,
                              hashes.as_mut_ptr() as *mut u8, ptr::null_mut())
    };
    match rc {
        sys::LCPC_OK => Ok(LcCommit { comm, coeffs, n_rows, n_cols, n_per_row, hashes }),
        sys::LCPC_ERR_TOO_BIG => Err(ProverError::TooBig),
        _ => Err(ProverError::Commit),
    }
}
