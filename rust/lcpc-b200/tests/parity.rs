//! Differential test to run on a machine with cargo + a B200: pins the one item the
//! in-container oracle cannot (the fffft NTT convention) and the whole commit.
use blake3::Hasher as Blake3;
use ff::Field;
use fffft::FieldFFT;
use lcpc_2d::{LcCommit, LcEncoding};
use lcpc_b200::{commit_gpu, GpuField, GpuLigeroEncoding};
use lcpc_ligero_pc::LigeroEncoding;
use lcpc_test_fields::ft63::Ft63;

impl GpuField for Ft63 {
    const FIELD_ID: i32 = 0;
    const LIMBS: usize = 1;
}

#[test]
fn encode_matches_fft_io_pc() {
    let mut rng = rand::thread_rng();
    for log_len in 1..=18u32 {
        let n = 1usize << log_len;
        let enc = GpuLigeroEncoding::<Ft63>::new_from_dims(n / 2, n, 0);
        let mut a: Vec<Ft63> = (0..n).map(|_| Ft63::random(&mut rng)).collect();
        let mut b = a.clone();
        enc.encode(&mut a).unwrap();
        <Ft63 as FieldFFT>::fft_io(&mut b).unwrap();
        assert_eq!(a, b, "log_len {log_len}");
    }
}

#[test]
fn commit_roots_match() {
    let mut rng = rand::thread_rng();
    let coeffs: Vec<Ft63> = (0..(1usize << 16)).map(|_| Ft63::random(&mut rng)).collect();
    let cpu_enc = LigeroEncoding::<Ft63>::new(coeffs.len());
    let (_, n_per_row, n_cols) = cpu_enc.get_dims(coeffs.len());
    let cpu = LcCommit::<Blake3, _>::commit(&coeffs, &cpu_enc).unwrap();
    let gpu = commit_gpu(&coeffs, &GpuLigeroEncoding::<Ft63>::new_from_dims(n_per_row, n_cols, 0)).unwrap();
    assert_eq!(cpu.get_root().as_ref(), gpu.get_root().as_ref());
    assert_eq!(cpu.comm, gpu.comm);
}
