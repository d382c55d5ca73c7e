//! Safe wrapper over liblcpc_b200 (include/lcpc_b200.h) for the lcpc crates.
//!
//! * `GpuLigeroEncodingRho<F, Rn, Rd>` / `GpuLigeroEncoding<F>` mirror `LigeroEncodingRho` / `LigeroEncoding`
//!   (lcpc-ligero-pc/src/lib.rs:31-186), `GpuSdigEncodingS<F, S>` / `GpuSdigEncoding<F>` mirror `SdigEncodingS` /
//!   `SdigEncoding` (lcpc-brakedown-pc/src/lib.rs:38-176).  All implement `lcpc_2d::LcEncoding`, so
//!   `LcCommit::prove`, `LcEvalProof::verify`, `open_column`, ... from lcpc-2d work on them unchanged.
//! * `commit_gpu` replaces `LcCommit::commit` (lcpc-2d/src/lib.rs:314 -> :651-700) and returns a genuine `LcCommit`
//!   with host `Vec`s; `GpuCommit` keeps the commitment resident in HBM for the server's repeated requests
//!   (proof-of-storage/src/networking/server.rs:360,530,600,...): `prove_gpu` (a real `LcEvalProof`), `fold`,
//!   `open_columns`, `leaves`, `update_rows` / `append_rows`.
//! * `GpuContext::multi(&devices)` shards commitments over the GPUs of one process inside the library.
//!
//! Written against include/lcpc_b200.h.  This crate has NOT been compiled: the build image has no Rust toolchain
//! (rust/lcpc-b200-sys/src/lib.rs is generated from the header and checked by the CPU test-suite; this file is not).
use std::ffi::CStr;
use std::marker::PhantomData;
use std::ptr;
use std::sync::Arc;

use blake3::Hasher as Blake3;
use digest::Output;
use ff::{Field, PrimeField};
use fffft::{FFTError, FieldFFT};
use lcpc_2d::{
    def_labels, n_degree_tests, FieldHash, LcColumn, LcCommit, LcEncoding, LcEvalProof, ProverError, SizedField,
};
use lcpc_b200_sys as sys;
use lcpc_brakedown_pc::codespec::{SdigCode3, SdigSpecification};
use merlin::Transcript;
use sprs::CsMat;
use typenum::{Unsigned, U1, U2};

/// Fields whose in-memory form is `[u64; LIMBS]` Montgomery limbs, least-significant limb first (every ff_derive field):
/// a `&[F]` crosses the boundary as `*const u64` with `len * LIMBS` words.
pub trait GpuField: PrimeField + FieldHash + SizedField {
    const FIELD_ID: i32;
    const LIMBS: usize;
}

/// `impl GpuField` for a concrete field type: `gpu_field!(Ft63, LCPC_FT63, 1);`
#[macro_export]
macro_rules! gpu_field {
    ($t:ty, $id:expr, $limbs:expr) => {
        impl $crate::GpuField for $t {
            const FIELD_ID: i32 = $id;
            const LIMBS: usize = $limbs;
        }
    };
}

/// The five fields of the reference (lcpc-test-fields/src/lib.rs:18-70; proof-of-storage/src/fields).
#[cfg(feature = "test-fields")]
mod test_field_impls {
    gpu_field!(lcpc_test_fields::ft63::Ft63, lcpc_b200_sys::LCPC_FT63, 1);
    gpu_field!(lcpc_test_fields::ft127::Ft127, lcpc_b200_sys::LCPC_FT127, 2);
    gpu_field!(lcpc_test_fields::ft191::Ft191, lcpc_b200_sys::LCPC_FT191, 3);
    gpu_field!(lcpc_test_fields::ft255::Ft255, lcpc_b200_sys::LCPC_FT255, 4);
}
#[cfg(feature = "proof-of-storage")]
mod pos_field_impls {
    gpu_field!(proof_of_storage::fields::WriteableFt63, lcpc_b200_sys::LCPC_FT63, 1);
    gpu_field!(proof_of_storage::fields::ft253_192::Ft253_192, lcpc_b200_sys::LCPC_FT253_192, 4);
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::lcpc_last_error()).to_string_lossy().into_owned() }
}

fn words<F: GpuField>(s: &[F]) -> *const u64 {
    debug_assert_eq!(std::mem::size_of::<F>(), 8 * F::LIMBS);
    s.as_ptr() as *const u64
}
fn words_mut<F: GpuField>(s: &mut [F]) -> *mut u64 {
    s.as_mut_ptr() as *mut u64
}

// ------------------------------------------------------------------------------------------------ context / plan

struct CtxInner(*mut sys::lcpc_ctx);
unsafe impl Send for CtxInner {}
unsafe impl Sync for CtxInner {} // handles are internally locked (header, "Conventions")
impl Drop for CtxInner {
    fn drop(&mut self) {
        unsafe { sys::lcpc_ctx_destroy(self.0) }
    }
}

/// One device + stream, or several devices of this process (`multi`).
#[derive(Clone)]
pub struct GpuContext(Arc<CtxInner>);

impl GpuContext {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut ctx = ptr::null_mut();
        match unsafe { sys::lcpc_ctx_create(device, &mut ctx) } {
            sys::LCPC_OK => Ok(Self(Arc::new(CtxInner(ctx)))),
            _ => Err(last_error()),
        }
    }
    /// lcpc_ctx_create_multi: commitments made through encodings on this context are sharded over `devices` inside the
    /// library (rows for encoding, column blocks of the leaf range for the Merkle subtrees); a power of two of them, <= 16.
    pub fn multi(devices: &[i32]) -> Result<Self, String> {
        let mut ctx = ptr::null_mut();
        match unsafe { sys::lcpc_ctx_create_multi(devices.as_ptr(), devices.len() as i32, &mut ctx) } {
            sys::LCPC_OK => Ok(Self(Arc::new(CtxInner(ctx)))),
            _ => Err(last_error()),
        }
    }
    pub fn n_devices(&self) -> usize {
        unsafe { sys::lcpc_ctx_device_count(self.0 .0) as usize }
    }
}

struct PlanInner {
    plan: *mut sys::lcpc_plan,
    _ctx: GpuContext, // the library reference-counts this too; kept for clarity
}
unsafe impl Send for PlanInner {}
unsafe impl Sync for PlanInner {}
impl Drop for PlanInner {
    fn drop(&mut self) {
        unsafe { sys::lcpc_plan_destroy(self.plan) }
    }
}

/// What `commit_gpu` / `GpuCommit` need from an encoding besides `LcEncoding`.
pub trait GpuEncoding: LcEncoding {
    fn plan(&self) -> *mut sys::lcpc_plan;
    /// maps a failure of the library's encode step onto `Self::Err`
    fn encode_error(msg: String) -> Self::Err;
}

// ------------------------------------------------------------------------------------------------ Ligero

/// `LigeroEncodingRho<Ft, Rn, Rd>` on the GPU (lcpc-ligero-pc/src/lib.rs:31-186).
#[derive(Clone)]
pub struct GpuLigeroEncodingRho<Ft, Rn, Rd> {
    n_per_row: usize,
    n_cols: usize,
    plan: Arc<PlanInner>,
    _p: PhantomData<(Ft, Rn, Rd)>,
}

impl<Ft, Rn, Rd> std::fmt::Debug for GpuLigeroEncodingRho<Ft, Rn, Rd> {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "GpuLigeroEncodingRho({} -> {})", self.n_per_row, self.n_cols)
    }
}

impl<Ft, Rn, Rd> GpuLigeroEncodingRho<Ft, Rn, Rd>
where
    Ft: GpuField + FieldFFT,
    Rn: Unsigned + std::fmt::Debug + Sync,
    Rd: Unsigned + std::fmt::Debug + Sync,
{
    const LAMBDA: usize = 128;

    fn _rho() -> f64 {
        assert!(Rn::to_usize() < Rd::to_usize());
        Rn::to_usize() as f64 / Rd::to_usize() as f64
    }
    // lcpc-ligero-pc/src/lib.rs:61-64
    fn _n_col_opens() -> usize {
        let den = ((1f64 + Self::_rho()) / 2f64).log2();
        (-(Self::LAMBDA as f64) / den).ceil() as usize
    }
    // :66-68
    fn _n_degree_tests(n_cols: usize) -> usize {
        n_degree_tests(Self::LAMBDA, n_cols, Ft::FLOG2 as usize)
    }
    // :70-112, unchanged
    fn _get_dims(len: usize) -> Option<(usize, usize, usize)> {
        let n_col_opens = Self::_n_col_opens();
        let lncf = (n_col_opens * len) as f64;
        let ndt = Self::_n_degree_tests((lncf.sqrt() / Self::_rho()).ceil() as usize) as f64;
        let nc1 = (((lncf / ndt).sqrt() / Self::_rho()).ceil() as usize)
            .checked_next_power_of_two()
            .and_then(|nc| if nc > (1 << <Ft as FieldFFT>::S) { None } else { Some(nc) })?;
        let np1 = nc1 * Rn::to_usize() / Rd::to_usize();
        let nr1 = (len + np1 - 1) / np1;
        let nd1 = Self::_n_degree_tests(nc1);
        let nc2 = nc1 / 2;
        let np2 = np1 / 2;
        let nr2 = (len + np2 - 1) / np2;
        let nd2 = Self::_n_degree_tests(nc2);
        let sz1 = n_col_opens * nr1 + (1 + nd1) * np1;
        let sz2 = n_col_opens * nr2 + (1 + nd2) * np2;
        Some(if sz1 < sz2 { (nr1, np1, nc1) } else { (nr2, np2, nc2) })
    }
    fn _dims_ok(n_per_row: usize, n_cols: usize) -> bool {
        n_per_row < n_cols && n_cols.is_power_of_two()
    }

    pub fn new(len: usize, ctx: &GpuContext) -> Self {
        let (_, n_per_row, n_cols) = Self::_get_dims(len).unwrap();
        Self::new_from_dims(n_per_row, n_cols, ctx)
    }
    pub fn new_ml(n_vars: usize, ctx: &GpuContext) -> Self {
        let n_monomials = 1 << n_vars;
        let (n_rows, n_per_row, n_cols) = Self::_get_dims(n_monomials).unwrap();
        assert!(n_rows.is_power_of_two() && n_per_row.is_power_of_two());
        assert_eq!(n_rows * n_per_row, n_monomials);
        Self::new_from_dims(n_per_row, n_cols, ctx)
    }
    /// :138-148.  The n_cols-th root of unity is handed to the library from `F::ROOT_OF_UNITY`, so the C side guesses no
    /// constant: w = ROOT_OF_UNITY^(2^(S - log2 n_cols)), what `precomp_fft(n_cols)` derives.
    pub fn new_from_dims(n_per_row: usize, n_cols: usize, ctx: &GpuContext) -> Self {
        assert!(Self::_dims_ok(n_per_row, n_cols));
        let log_len = n_cols.trailing_zeros();
        let mut root = Ft::ROOT_OF_UNITY;
        for _ in 0..(<Ft as FieldFFT>::S - log_len) {
            root *= root;
        }
        let mut plan = ptr::null_mut();
        let rc = unsafe {
            sys::lcpc_plan_ligero(ctx.0 .0, Ft::FIELD_ID, n_per_row, n_cols, &root as *const Ft as *const u64, &mut plan)
        };
        assert_eq!(rc, sys::LCPC_OK, "{}", last_error());
        Self { n_per_row, n_cols, plan: Arc::new(PlanInner { plan, _ctx: ctx.clone() }), _p: PhantomData }
    }

    /// proof-of-storage `decode_row` (lcpc_online.rs:568-573): fffft `ifft_oi` on whole encoded rows, in place.
    pub fn decode_rows(&self, rows: &mut [Ft]) -> Result<(), FFTError> {
        if rows.len() % self.n_cols != 0 {
            return Err(FFTError::TooBig);
        }
        let rc = unsafe { sys::lcpc_decode_rows(self.plan.plan, words_mut(rows), rows.len() / self.n_cols) };
        if rc == sys::LCPC_OK { Ok(()) } else { Err(FFTError::TooBig) }
    }
}

impl<Ft, Rn, Rd> LcEncoding for GpuLigeroEncodingRho<Ft, Rn, Rd>
where
    Ft: GpuField + FieldFFT,
    Rn: Unsigned + std::fmt::Debug + Sync + Clone,
    Rd: Unsigned + std::fmt::Debug + Sync + Clone,
{
    type F = Ft;
    type Err = FFTError;
    def_labels!(ligero_pc);

    /// one row (or several whole rows) in place; lcpc-2d calls this from rayon workers, the plan serialises them
    fn encode<T: AsMut<[Ft]>>(&self, mut inp: T) -> Result<(), FFTError> {
        let row = inp.as_mut();
        if row.is_empty() || row.len() % self.n_cols != 0 {
            return Err(FFTError::TooBig); // closest fffft error for a length mismatch
        }
        let rc = unsafe { sys::lcpc_encode_rows(self.plan.plan, words_mut(row), row.len() / self.n_cols) };
        if rc == sys::LCPC_OK { Ok(()) } else { Err(FFTError::TooBig) }
    }
    fn get_dims(&self, len: usize) -> (usize, usize, usize) {
        ((len + self.n_per_row - 1) / self.n_per_row, self.n_per_row, self.n_cols)
    }
    fn dims_ok(&self, n_per_row: usize, n_cols: usize) -> bool {
        Self::_dims_ok(n_per_row, n_cols) && n_per_row == self.n_per_row && n_cols == self.n_cols
    }
    fn get_n_col_opens(&self) -> usize {
        Self::_n_col_opens()
    }
    fn get_n_degree_tests(&self) -> usize {
        Self::_n_degree_tests(self.n_cols)
    }
}

impl<Ft, Rn, Rd> GpuEncoding for GpuLigeroEncodingRho<Ft, Rn, Rd>
where
    Ft: GpuField + FieldFFT,
    Rn: Unsigned + std::fmt::Debug + Sync + Clone,
    Rd: Unsigned + std::fmt::Debug + Sync + Clone,
{
    fn plan(&self) -> *mut sys::lcpc_plan {
        self.plan.plan
    }
    fn encode_error(_msg: String) -> FFTError {
        FFTError::TooBig
    }
}

/// rho = 1/2, lambda = 128 (lcpc-ligero-pc/src/lib.rs:189)
pub type GpuLigeroEncoding<F> = GpuLigeroEncodingRho<F, U1, U2>;

// ------------------------------------------------------------------------------------------------ Brakedown

/// `SdigEncodingS<Ft, S>` on the GPU (lcpc-brakedown-pc/src/lib.rs:38-176).  The code matrices are generated by the
/// reference's own `matgen::generate` (host, Rust) and handed to the library as CSC arrays.
#[derive(Clone)]
pub struct GpuSdigEncodingS<Ft, S> {
    n_per_row: usize,
    n_cols: usize,
    n_pre0_cols: usize,
    plan: Arc<PlanInner>,
    _p: PhantomData<(Ft, S)>,
}

impl<Ft, S> std::fmt::Debug for GpuSdigEncodingS<Ft, S> {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "GpuSdigEncodingS({} -> {})", self.n_per_row, self.n_cols)
    }
}

fn csc_of<F: GpuField>(m: &CsMat<F>, indptr: &mut Vec<u64>, indices: &mut Vec<u64>) -> sys::lcpc_csc {
    assert!(m.is_csc());
    // sprs keeps usize index arrays; the ABI takes u64 (identical on the 64-bit targets this runs on, copied to be safe)
    indptr.clear();
    indptr.extend(m.indptr().raw_storage().iter().map(|&v| v as u64));
    indices.clear();
    indices.extend(m.indices().iter().map(|&v| v as u64));
    sys::lcpc_csc {
        rows: m.rows() as u64,
        cols: m.cols() as u64,
        indptr: indptr.as_ptr(),
        indices: indices.as_ptr(),
        data: words(m.data()),
    }
}

impl<Ft, S> GpuSdigEncodingS<Ft, S>
where
    Ft: GpuField + num_traits::Num,
    S: SdigSpecification,
{
    const LAMBDA: usize = 128;

    // lcpc-brakedown-pc/src/lib.rs:57-61
    fn _n_col_opens() -> usize {
        let den = (1f64 - S::dist() / 3f64).log2();
        (-(Self::LAMBDA as f64) / den).ceil() as usize
    }
    // :64-66
    fn _n_degree_tests(n_cols: usize) -> usize {
        n_degree_tests(Self::LAMBDA, n_cols, Ft::FLOG2 as usize)
    }

    fn from_matrices(n_per_row: usize, precodes: &[CsMat<Ft>], postcodes: &[CsMat<Ft>], ctx: &GpuContext) -> Self {
        assert_eq!(n_per_row, precodes[0].cols());
        let n_cols = lcpc_brakedown_pc::encode::codeword_length(precodes, postcodes);
        let n = precodes.len();
        let (mut ip, mut ix): (Vec<Vec<u64>>, Vec<Vec<u64>>) = (vec![Vec::new(); 2 * n], vec![Vec::new(); 2 * n]);
        let mut pre = Vec::with_capacity(n);
        let mut post = Vec::with_capacity(n);
        for (i, m) in precodes.iter().enumerate() {
            let (a, b) = (&mut ip[i] as *mut Vec<u64>, &mut ix[i] as *mut Vec<u64>);
            pre.push(unsafe { csc_of(m, &mut *a, &mut *b) });
        }
        for (i, m) in postcodes.iter().enumerate() {
            let (a, b) = (&mut ip[n + i] as *mut Vec<u64>, &mut ix[n + i] as *mut Vec<u64>);
            post.push(unsafe { csc_of(m, &mut *a, &mut *b) });
        }
        let mut plan = ptr::null_mut();
        let rc = unsafe {
            sys::lcpc_plan_brakedown(ctx.0 .0, Ft::FIELD_ID, n_per_row, n_cols, n, pre.as_ptr(), post.as_ptr(), &mut plan)
        };
        assert_eq!(rc, sys::LCPC_OK, "{}", last_error());
        Self {
            n_per_row,
            n_cols,
            n_pre0_cols: precodes[0].cols(),
            plan: Arc::new(PlanInner { plan, _ctx: ctx.clone() }),
            _p: PhantomData,
        }
    }

    // :69-110, unchanged except that the matrices go to the device
    fn _new_from_np1(len: usize, np1: usize, seed: u64, ctx: &GpuContext) -> Self {
        let np1 = if np1 > len { len } else { np1 };
        let n_col_opens = Self::_n_col_opens();
        let nr1 = (len + np1 - 1) / np1;
        let nd1 = Self::_n_degree_tests(np1 * 2);
        let np2 = np1 / 2;
        let nr2 = (len + np2 - 1) / np2;
        let nd2 = Self::_n_degree_tests(np2 * 2);
        let sz1 = n_col_opens * nr1 + (1 + nd1) * np1;
        let sz2 = n_col_opens * nr2 + (1 + nd2) * np2;
        let n_per_row = if sz1 < sz2 { np1 } else { np2 };
        let (precodes, postcodes) = lcpc_brakedown_pc::matgen::generate::<Ft, S>(n_per_row, seed);
        Self::from_matrices(n_per_row, &precodes, &postcodes, ctx)
    }
    pub fn new(len: usize, seed: u64, ctx: &GpuContext) -> Self {
        let lncf = (Self::_n_col_opens() * len) as f64;
        let ndt = Self::_n_degree_tests(lncf.sqrt().ceil() as usize * 2) as f64;
        let np1 = (lncf / ndt).sqrt().ceil() as usize;
        Self::_new_from_np1(len, np1, seed, ctx)
    }
    pub fn new_ml(n_vars: usize, seed: u64, ctx: &GpuContext) -> Self {
        let n_monomials = 1 << n_vars;
        let lncf = (Self::_n_col_opens() * n_monomials) as f64;
        let ndt = Self::_n_degree_tests(lncf.sqrt().ceil() as usize * 2) as f64;
        let np1 = ((lncf / ndt).sqrt().ceil() as usize).checked_next_power_of_two().unwrap();
        Self::_new_from_np1(n_monomials, np1, seed, ctx)
    }
    pub fn new_from_dims(n_per_row: usize, n_cols: usize, seed: u64, ctx: &GpuContext) -> Self {
        let (precodes, postcodes) = lcpc_brakedown_pc::matgen::generate::<Ft, S>(n_per_row, seed);
        let enc = Self::from_matrices(n_per_row, &precodes, &postcodes, ctx);
        assert_eq!(n_cols, enc.n_cols);
        enc
    }
}

impl<Ft, S> LcEncoding for GpuSdigEncodingS<Ft, S>
where
    Ft: GpuField + num_traits::Num,
    S: SdigSpecification + std::fmt::Debug + Clone + Sync,
{
    type F = Ft;
    type Err = std::io::Error;
    def_labels!(sdig_pc);

    fn encode<T: AsMut<[Ft]>>(&self, mut inp: T) -> Result<(), Self::Err> {
        let row = inp.as_mut();
        if row.is_empty() || row.len() % self.n_cols != 0 {
            return Err(std::io::Error::new(std::io::ErrorKind::InvalidInput, "row length must be n_cols"));
        }
        match unsafe { sys::lcpc_encode_rows(self.plan.plan, words_mut(row), row.len() / self.n_cols) } {
            sys::LCPC_OK => Ok(()),
            _ => Err(std::io::Error::new(std::io::ErrorKind::Other, last_error())),
        }
    }
    fn get_dims(&self, len: usize) -> (usize, usize, usize) {
        ((len + self.n_per_row - 1) / self.n_per_row, self.n_per_row, self.n_cols)
    }
    fn dims_ok(&self, n_per_row: usize, n_cols: usize) -> bool {
        n_per_row < n_cols && n_per_row == self.n_per_row && n_per_row == self.n_pre0_cols && n_cols == self.n_cols
    }
    fn get_n_col_opens(&self) -> usize {
        Self::_n_col_opens()
    }
    fn get_n_degree_tests(&self) -> usize {
        Self::_n_degree_tests(self.n_cols)
    }
}

impl<Ft, S> GpuEncoding for GpuSdigEncodingS<Ft, S>
where
    Ft: GpuField + num_traits::Num,
    S: SdigSpecification + std::fmt::Debug + Clone + Sync,
{
    fn plan(&self) -> *mut sys::lcpc_plan {
        self.plan.plan
    }
    fn encode_error(msg: String) -> std::io::Error {
        std::io::Error::new(std::io::ErrorKind::Other, msg)
    }
}

/// default code (lcpc-brakedown-pc/src/lib.rs:19)
pub type GpuSdigEncoding<F> = GpuSdigEncodingS<F, SdigCode3>;

// ------------------------------------------------------------------------------------------------ commit

fn prover_error<E: GpuEncoding>(rc: i32) -> ProverError<E::Err> {
    match rc {
        sys::LCPC_ERR_TOO_BIG => ProverError::TooBig,
        sys::LCPC_ERR_ENCODE => ProverError::Encode(E::encode_error(last_error())),
        sys::LCPC_ERR_COLUMN_NUMBER => ProverError::ColumnNumber,
        sys::LCPC_ERR_OUTER_TENSOR => ProverError::OuterTensor,
        _ => ProverError::Commit,
    }
}

/// Drop-in for `LcCommit::<Blake3, E>::commit(coeffs, enc)` (lcpc-2d/src/lib.rs:314): pad, encode every row, hash every
/// column, build the Merkle tree -- on the GPU(s) of `enc`'s context; the result holds host `Vec`s like the reference's.
pub fn commit_gpu<E>(coeffs_in: &[E::F], enc: &E) -> Result<LcCommit<Blake3, E>, ProverError<E::Err>>
where
    E: GpuEncoding,
    E::F: GpuField,
{
    let (n_rows, n_per_row, n_cols) = enc.get_dims(coeffs_in.len());
    // lib.rs:659-661
    assert!(n_rows * n_per_row >= coeffs_in.len());
    assert!((n_rows - 1) * n_per_row < coeffs_in.len());
    assert!(enc.dims_ok(n_per_row, n_cols));
    let np2 = n_cols.checked_next_power_of_two().ok_or(ProverError::TooBig)?;
    let mut coeffs = vec![<E::F as Field>::ZERO; n_rows * n_per_row];
    let mut comm = vec![<E::F as Field>::ZERO; n_rows * n_cols];
    let mut hashes = vec![<Output<Blake3> as Default>::default(); 2 * np2 - 1];
    let rc = unsafe {
        sys::lcpc_commit_host(enc.plan(), words(coeffs_in), coeffs_in.len(), words_mut(&mut coeffs), words_mut(&mut comm),
                              hashes.as_mut_ptr() as *mut u8, ptr::null_mut())
    };
    match rc {
        sys::LCPC_OK => Ok(LcCommit { comm, coeffs, n_rows, n_cols, n_per_row, hashes }),
        rc => Err(prover_error::<E>(rc)),
    }
}

/// A commitment kept resident in HBM (one GPU, or sharded over the context's GPUs): what a server answering openings,
/// proofs and evaluations from the same file needs instead of recommitting per request.
pub struct GpuCommit<E: GpuEncoding> {
    h: *mut sys::lcpc_commit,
    pub n_rows: usize,
    pub n_per_row: usize,
    pub n_cols: usize,
    _enc: E,
}
unsafe impl<E: GpuEncoding + Send> Send for GpuCommit<E> {}
unsafe impl<E: GpuEncoding + Sync> Sync for GpuCommit<E> {}
impl<E: GpuEncoding> Drop for GpuCommit<E> {
    fn drop(&mut self) {
        unsafe { sys::lcpc_commit_free(self.h) }
    }
}

impl<E> GpuCommit<E>
where
    E: GpuEncoding + Clone,
    E::F: GpuField,
{
    pub fn commit(coeffs_in: &[E::F], enc: &E) -> Result<Self, ProverError<E::Err>> {
        let (n_rows, n_per_row, n_cols) = enc.get_dims(coeffs_in.len());
        let mut h = ptr::null_mut();
        let rc = unsafe {
            sys::lcpc_commit_host(enc.plan(), words(coeffs_in), coeffs_in.len(), ptr::null_mut(), ptr::null_mut(),
                                  ptr::null_mut(), &mut h)
        };
        if rc != sys::LCPC_OK {
            return Err(prover_error::<E>(rc));
        }
        Ok(Self { h, n_rows, n_per_row, n_cols, _enc: enc.clone() })
    }
    /// proof-of-storage `convert_file_data_to_commit` (lcpc_online.rs:81-143): the 7-byte (WriteableFt63) or 31-byte
    /// (Ft253_192) packing runs on the device in front of the commit.
    pub fn commit_bytes(file: &[u8], bytes_per_elem: usize, enc: &E) -> Result<Self, ProverError<E::Err>> {
        let n_elems = (file.len() + bytes_per_elem - 1) / bytes_per_elem;
        let (n_rows, n_per_row, n_cols) = enc.get_dims(n_elems);
        let mut h = ptr::null_mut();
        let rc = unsafe {
            sys::lcpc_commit_bytes_host(enc.plan(), file.as_ptr(), file.len(), ptr::null_mut(), ptr::null_mut(), ptr::null_mut(), &mut h)
        };
        if rc != sys::LCPC_OK {
            return Err(prover_error::<E>(rc));
        }
        Ok(Self { h, n_rows, n_per_row, n_cols, _enc: enc.clone() })
    }
    /// LcCommit::get_root (lib.rs:291-296)
    pub fn get_root(&self) -> Output<Blake3> {
        let mut root = <Output<Blake3> as Default>::default();
        let rc = unsafe { sys::lcpc_commit_root(self.h, root.as_mut_ptr()) };
        assert_eq!(rc, sys::LCPC_OK, "{}", last_error());
        root
    }
    /// The whole `LcCommit` (coeffs, comm, hashes) copied back to the host.
    pub fn download(&self) -> LcCommit<Blake3, E> {
        let np2 = self.n_cols.next_power_of_two();
        let mut coeffs = vec![<E::F as Field>::ZERO; self.n_rows * self.n_per_row];
        let mut comm = vec![<E::F as Field>::ZERO; self.n_rows * self.n_cols];
        let mut hashes = vec![<Output<Blake3> as Default>::default(); 2 * np2 - 1];
        let rc = unsafe {
            sys::lcpc_commit_download(self.h, words_mut(&mut coeffs), words_mut(&mut comm), hashes.as_mut_ptr() as *mut u8)
        };
        assert_eq!(rc, sys::LCPC_OK, "{}", last_error());
        LcCommit { comm, coeffs, n_rows: self.n_rows, n_cols: self.n_cols, n_per_row: self.n_per_row, hashes }
    }
    /// collapse_columns (lib.rs:1126-1154) for a batch of tensors; `encoded` folds the encoded matrix instead
    /// (proof-of-storage verifiable_polynomial_evaluation, lcpc_online.rs:454-484).
    pub fn fold(&self, tensors: &[&[E::F]], encoded: bool) -> Result<Vec<Vec<E::F>>, ProverError<E::Err>> {
        let width = if encoded { self.n_cols } else { self.n_per_row };
        let mut flat = Vec::with_capacity(tensors.len() * self.n_rows);
        for t in tensors {
            if t.len() != self.n_rows {
                return Err(ProverError::OuterTensor);
            }
            flat.extend_from_slice(t);
        }
        let mut out = vec![<E::F as Field>::ZERO; tensors.len() * width];
        let rc = unsafe { sys::lcpc_fold_host(self.h, encoded as i32, words(&flat), tensors.len(), words_mut(&mut out)) };
        if rc != sys::LCPC_OK {
            return Err(prover_error::<E>(rc));
        }
        Ok(out.chunks(width).map(|c| c.to_vec()).collect())
    }
    /// open_column (lib.rs:818-855) for every index: values and Merkle paths.
    pub fn open_columns(&self, cols: &[usize]) -> Result<Vec<LcColumn<Blake3, E>>, ProverError<E::Err>> {
        let depth = self.n_cols.next_power_of_two().trailing_zeros() as usize;
        let idx: Vec<u64> = cols.iter().map(|&c| c as u64).collect();
        let mut vals = vec![<E::F as Field>::ZERO; cols.len() * self.n_rows];
        let mut paths = vec![<Output<Blake3> as Default>::default(); cols.len() * depth];
        let rc = unsafe {
            sys::lcpc_open_columns_host(self.h, idx.as_ptr(), idx.len(), words_mut(&mut vals), paths.as_mut_ptr() as *mut u8)
        };
        if rc != sys::LCPC_OK {
            return Err(prover_error::<E>(rc));
        }
        Ok((0..cols.len())
            .map(|i| LcColumn {
                col: vals[i * self.n_rows..(i + 1) * self.n_rows].to_vec(),
                path: paths[i * depth..(i + 1) * depth].to_vec(),
            })
            .collect())
    }
    /// CommitRequestType::Leaves (lcpc_online.rs:144-190)
    pub fn leaves(&self, cols: &[usize]) -> Result<Vec<Output<Blake3>>, ProverError<E::Err>> {
        let idx: Vec<u64> = cols.iter().map(|&c| c as u64).collect();
        let mut out = vec![<Output<Blake3> as Default>::default(); cols.len()];
        let rc = unsafe { sys::lcpc_leaves_host(self.h, idx.as_ptr(), idx.len(), out.as_mut_ptr() as *mut u8) };
        if rc != sys::LCPC_OK {
            return Err(prover_error::<E>(rc));
        }
        Ok(out)
    }
    /// FileHandler::edit_bytes -> reencode_row -> recalculate_merkle_tree (file_handler.rs:279-402, 474-481): whole
    /// coefficient rows replaced, only their BLAKE3 chunks of the leaves re-hashed.  Returns the new root.
    pub fn update_rows(&mut self, row0: usize, coeff_rows: &[E::F]) -> Result<Output<Blake3>, ProverError<E::Err>> {
        assert_eq!(coeff_rows.len() % self.n_per_row, 0);
        let rc = unsafe {
            sys::lcpc_commit_update_rows_host(self.h, row0, coeff_rows.len() / self.n_per_row, words(coeff_rows), ptr::null_mut(),
                                              ptr::null_mut())
        };
        if rc != sys::LCPC_OK {
            return Err(prover_error::<E>(rc));
        }
        Ok(self.get_root())
    }
    /// FileHandler::append_bytes (file_handler.rs:336-402)
    pub fn append_rows(&mut self, row0: usize, coeff_rows: &[E::F]) -> Result<Output<Blake3>, ProverError<E::Err>> {
        assert_eq!(coeff_rows.len() % self.n_per_row, 0);
        let n = coeff_rows.len() / self.n_per_row;
        let rc = unsafe {
            sys::lcpc_commit_append_rows_host(self.h, row0, n, words(coeff_rows), ptr::null_mut(), ptr::null_mut())
        };
        if rc != sys::LCPC_OK {
            return Err(prover_error::<E>(rc));
        }
        self.n_rows = self.n_rows.max(row0 + n);
        Ok(self.get_root())
    }
}

// ------------------------------------------------------------------------------------------------ prove

/// merlin's transcript lives in Rust; the library has its own copy of the same construction (STROBE-128 / Keccak) for
/// `lcpc_prove`.  The two are kept in step by replaying, into the library's transcript, exactly the messages the Rust
/// transcript has received so far -- which the caller knows (lcpc tests: one `append_message(b"polycommit", root)`).
pub struct MirroredTranscript {
    pub rust: Transcript,
    lib: *mut sys::lcpc_transcript,
}
impl MirroredTranscript {
    pub fn new(label: &'static [u8]) -> Self {
        let mut lib = ptr::null_mut();
        let rc = unsafe { sys::lcpc_transcript_new(label.as_ptr(), label.len(), &mut lib) };
        assert_eq!(rc, sys::LCPC_OK, "{}", last_error());
        Self { rust: Transcript::new(label), lib }
    }
    pub fn append_message(&mut self, label: &'static [u8], msg: &[u8]) {
        self.rust.append_message(label, msg);
        let rc = unsafe { sys::lcpc_transcript_append_message(self.lib, label.as_ptr(), label.len(), msg.as_ptr(), msg.len()) };
        assert_eq!(rc, sys::LCPC_OK, "{}", last_error());
    }
}
impl Drop for MirroredTranscript {
    fn drop(&mut self) {
        unsafe { sys::lcpc_transcript_free(self.lib) }
    }
}

/// Drop-in for `comm.prove(outer_tensor, enc, tr)` (lcpc-2d/src/lib.rs:319 -> :1034-1123) on a resident commitment:
/// degree-test folds, the evaluation fold and the column openings run on the GPU, the transcript (merlin, ChaCha20
/// challenge expansion, `F::random`, `Uniform`) inside the library's host code, and the result is a real `LcEvalProof`
/// that `verify` from lcpc-2d accepts.  `tr.rust` is advanced by the same messages, so prover and verifier stay in step.
pub fn prove_gpu<E>(comm: &GpuCommit<E>, outer_tensor: &[E::F], enc: &E, tr: &mut MirroredTranscript)
    -> Result<LcEvalProof<Blake3, E>, ProverError<E::Err>>
where
    E: GpuEncoding + Clone,
    E::F: GpuField,
{
    let (n_dt, n_open) = (enc.get_n_degree_tests(), enc.get_n_col_opens());
    let depth = comm.n_cols.next_power_of_two().trailing_zeros() as usize;
    let mut p_eval = vec![<E::F as Field>::ZERO; comm.n_per_row];
    let mut p_random = vec![<E::F as Field>::ZERO; n_dt * comm.n_per_row];
    let mut cols = vec![<E::F as Field>::ZERO; n_open * comm.n_rows];
    let mut paths = vec![<Output<Blake3> as Default>::default(); n_open * depth];
    let rc = unsafe {
        sys::lcpc_prove(comm.h, words(outer_tensor), outer_tensor.len(), n_dt, n_open, tr.lib, words_mut(&mut p_eval),
                        words_mut(&mut p_random), ptr::null_mut(), words_mut(&mut cols), paths.as_mut_ptr() as *mut u8)
    };
    if rc != sys::LCPC_OK {
        return Err(prover_error::<E>(rc));
    }
    // replay on the Rust transcript what prove() feeds it (lib.rs:1054-1110), so that `tr.rust` ends in the same state
    for i in 0..n_dt {
        let mut key = [0u8; 32];
        tr.rust.challenge_bytes(E::LABEL_DT, &mut key);
        p_random[i * comm.n_per_row..(i + 1) * comm.n_per_row].iter().for_each(|c| c.transcript_update(&mut tr.rust, E::LABEL_PR));
    }
    p_eval.iter().for_each(|c| c.transcript_update(&mut tr.rust, E::LABEL_PE));
    let mut key = [0u8; 32];
    tr.rust.challenge_bytes(E::LABEL_CO, &mut key);
    Ok(LcEvalProof {
        n_cols: comm.n_cols,
        p_eval,
        p_random_vec: p_random.chunks(comm.n_per_row).map(|c| c.to_vec()).collect(),
        columns: (0..n_open)
            .map(|i| LcColumn {
                col: cols[i * comm.n_rows..(i + 1) * comm.n_rows].to_vec(),
                path: paths[i * depth..(i + 1) * depth].to_vec(),
            })
            .collect(),
    })
}
