//! Differential tests to run on a machine with cargo + a B200: ONE `cargo test --release` pins every convention the
//! in-container oracle could only restate from memory ("parity unpinned"):
//!
//!  * the fffft NTT (`fft_io` root and output permutation, `ifft_oi`) for 2^1..2^18;
//!  * whole commitments (encoded matrix, every digest of the tree) for Ligero and Brakedown;
//!  * the host-side randomness of the library against rand 0.8 / rand_chacha 0.3 / ff 0.13: `F::random`
//!    (`lcpc_random_field_vec`), `Uniform::new(0, n).sample` (`lcpc_random_columns`), `ChaCha8Rng::seed_from_u64(1337)` +
//!    `choose_multiple` (`lcpc_pos_choose_columns`), and `matgen::generate` for seeds 0 and 1 (`lcpc_sdig_gen_level`);
//!  * one full evaluation proof: `prove_gpu` == `LcCommit::prove`, and lcpc-2d's `verify` accepts it.
use blake3::Hasher as Blake3;
use ff::Field;
use fffft::FieldFFT;
use lcpc_2d::{LcCommit, LcEncoding};
use lcpc_b200::{commit_gpu, prove_gpu, GpuCommit, GpuContext, GpuLigeroEncoding, GpuSdigEncoding, MirroredTranscript};
use lcpc_b200_sys as sys;
use lcpc_brakedown_pc::{codespec::SdigCode3, matgen, SdigEncoding};
use lcpc_ligero_pc::LigeroEncoding;
use lcpc_test_fields::{ft255::Ft255, ft63::Ft63};
use merlin::Transcript;
use rand::{distributions::{Distribution, Uniform}, seq::IteratorRandom, SeedableRng};
use rand_chacha::{ChaCha20Rng, ChaCha8Rng};

fn ctx() -> GpuContext {
    GpuContext::new(0).unwrap()
}

#[test]
fn encode_matches_fft_io_and_decode_matches_ifft_oi() {
    let mut rng = rand::thread_rng();
    let ctx = ctx();
    for log_len in 1..=18u32 {
        let n = 1usize << log_len;
        let enc = GpuLigeroEncoding::<Ft63>::new_from_dims(n / 2, n, &ctx);
        let mut a: Vec<Ft63> = (0..n).map(|_| Ft63::random(&mut rng)).collect();
        let mut b = a.clone();
        enc.encode(&mut a).unwrap();
        <Ft63 as FieldFFT>::fft_io(&mut b).unwrap();
        assert_eq!(a, b, "fft_io, log_len {log_len}");
        enc.decode_rows(&mut a).unwrap();
        <Ft63 as FieldFFT>::ifft_oi(&mut b).unwrap();
        assert_eq!(a, b, "ifft_oi, log_len {log_len}");
    }
    let enc = GpuLigeroEncoding::<Ft255>::new_from_dims(1 << 11, 1 << 12, &ctx);
    let mut a: Vec<Ft255> = (0..(1 << 12)).map(|_| Ft255::random(&mut rng)).collect();
    let mut b = a.clone();
    enc.encode(&mut a).unwrap();
    <Ft255 as FieldFFT>::fft_io(&mut b).unwrap();
    assert_eq!(a, b);
}

#[test]
fn ligero_commit_matches() {
    let mut rng = rand::thread_rng();
    let coeffs: Vec<Ft63> = (0..(1usize << 16)).map(|_| Ft63::random(&mut rng)).collect();
    let cpu_enc = LigeroEncoding::<Ft63>::new(coeffs.len());
    let gpu_enc = GpuLigeroEncoding::<Ft63>::new(coeffs.len(), &ctx());
    assert_eq!(cpu_enc.get_dims(coeffs.len()), gpu_enc.get_dims(coeffs.len()));
    assert_eq!(cpu_enc.get_n_col_opens(), gpu_enc.get_n_col_opens());
    assert_eq!(cpu_enc.get_n_degree_tests(), gpu_enc.get_n_degree_tests());
    let cpu = LcCommit::<Blake3, _>::commit(&coeffs, &cpu_enc).unwrap();
    let gpu = commit_gpu(&coeffs, &gpu_enc).unwrap();
    assert_eq!(cpu.comm, gpu.comm);
    assert_eq!(cpu.coeffs, gpu.coeffs);
    assert_eq!(cpu.hashes, gpu.hashes);
    assert_eq!(cpu.get_root().as_ref(), gpu.get_root().as_ref());
}

#[test]
fn brakedown_matgen_and_commit_match() {
    let mut rng = rand::thread_rng();
    for seed in [0u64, 1u64] {
        let len = 60_000usize;
        let coeffs: Vec<Ft255> = (0..len).map(|_| Ft255::random(&mut rng)).collect();
        let cpu_enc = SdigEncoding::<Ft255>::new(len, seed);
        let gpu_enc = GpuSdigEncoding::<Ft255>::new(len, seed, &ctx());
        let (_, n_per_row, n_cols) = cpu_enc.get_dims(len);
        assert_eq!((n_per_row, n_cols), { let d = gpu_enc.get_dims(len); (d.1, d.2) });
        let cpu = LcCommit::<Blake3, _>::commit(&coeffs, &cpu_enc).unwrap();
        let gpu = commit_gpu(&coeffs, &gpu_enc).unwrap();
        assert_eq!(cpu.comm, gpu.comm, "seed {seed}");
        assert_eq!(cpu.hashes, gpu.hashes, "seed {seed}");
        // the library's own host-side code generation (what the C++ / Python hosts use) against matgen::generate
        let (pre, post) = matgen::generate::<Ft255, SdigCode3>(n_per_row, seed);
        let mut pre_dims = vec![0u64; 3 * 64];
        let mut post_dims = vec![0u64; 3 * 64];
        let mut levels = 0i32;
        let rc = unsafe { sys::lcpc_sdig_get_dims(3, n_per_row as u64, sys::LCPC_FT255, pre_dims.as_mut_ptr(), post_dims.as_mut_ptr(), 64, &mut levels) };
        assert_eq!(rc, sys::LCPC_OK);
        assert_eq!(levels as usize, pre.len());
        for l in 0..pre.len() {
            let (pd, qd) = (&pre_dims[3 * l..3 * l + 3], &post_dims[3 * l..3 * l + 3]);
            assert_eq!((pd[0] as usize, pd[1] as usize), (pre[l].cols(), pre[l].rows()));
            assert_eq!((qd[0] as usize, qd[1] as usize), (post[l].cols(), post[l].rows()));
            let (mut a_ip, mut a_ix, mut a_d) = (vec![0u64; pd[0] as usize + 1], vec![0u64; (pd[0] * pd[2]) as usize], vec![0u64; (pd[0] * pd[2] * 4) as usize]);
            let (mut b_ip, mut b_ix, mut b_d) = (vec![0u64; qd[0] as usize + 1], vec![0u64; (qd[0] * qd[2]) as usize], vec![0u64; (qd[0] * qd[2] * 4) as usize]);
            let rc = unsafe {
                sys::lcpc_sdig_gen_level(sys::LCPC_FT255, seed, l as u64, pd.as_ptr(), qd.as_ptr(), a_ip.as_mut_ptr(), a_ix.as_mut_ptr(),
                                         a_d.as_mut_ptr(), b_ip.as_mut_ptr(), b_ix.as_mut_ptr(), b_d.as_mut_ptr())
            };
            assert_eq!(rc, sys::LCPC_OK);
            let as_u64 = |v: &[usize]| v.iter().map(|&x| x as u64).collect::<Vec<_>>();
            assert_eq!(a_ix, as_u64(pre[l].indices()), "precode indices, level {l}, seed {seed}");
            assert_eq!(b_ix, as_u64(post[l].indices()), "postcode indices, level {l}, seed {seed}");
            let limbs = |m: &[Ft255]| unsafe { std::slice::from_raw_parts(m.as_ptr() as *const u64, m.len() * 4).to_vec() };
            assert_eq!(a_d, limbs(pre[l].data()), "precode values, level {l}, seed {seed}");
            assert_eq!(b_d, limbs(post[l].data()), "postcode values, level {l}, seed {seed}");
        }
    }
}

#[test]
fn host_randomness_matches_rand_and_ff() {
    let key = [7u8; 32];
    // F::random x n from ChaCha20Rng::from_seed(key) (lcpc-2d/src/lib.rs:1058-1062)
    let mut rng = ChaCha20Rng::from_seed(key);
    let want: Vec<Ft63> = (0..1000).map(|_| Ft63::random(&mut rng)).collect();
    let mut got = vec![Ft63::ZERO; 1000];
    assert_eq!(unsafe { sys::lcpc_random_field_vec(sys::LCPC_FT63, key.as_ptr(), got.as_mut_ptr() as *mut u64, 1000) }, sys::LCPC_OK);
    assert_eq!(got, want);
    let mut rng = ChaCha20Rng::from_seed(key);
    let want: Vec<Ft255> = (0..1000).map(|_| Ft255::random(&mut rng)).collect();
    let mut got = vec![Ft255::ZERO; 1000];
    assert_eq!(unsafe { sys::lcpc_random_field_vec(sys::LCPC_FT255, key.as_ptr(), got.as_mut_ptr() as *mut u64, 1000) }, sys::LCPC_OK);
    assert_eq!(got, want);
    // Uniform::new(0, n_cols) x n (lib.rs:1105-1110)
    for n_cols in [4096usize, 252_931, 1 << 20] {
        let mut rng = ChaCha20Rng::from_seed(key);
        let range = Uniform::new(0usize, n_cols);
        let want: Vec<u64> = (0..6593).map(|_| range.sample(&mut rng) as u64).collect();
        let mut got = vec![0u64; 6593];
        assert_eq!(unsafe { sys::lcpc_random_columns(key.as_ptr(), n_cols as u64, got.as_mut_ptr(), 6593) }, sys::LCPC_OK);
        assert_eq!(got, want, "n_cols {n_cols}");
    }
    // proof-of-storage column choice (networking/client.rs:443-456)
    let mut rng = ChaCha8Rng::seed_from_u64(1337);
    let want: Vec<u64> = (0..65536usize).choose_multiple(&mut rng, 309).into_iter().map(|c| c as u64).collect();
    let mut got = vec![0u64; 309];
    let mut n = 0usize;
    assert_eq!(unsafe { sys::lcpc_pos_choose_columns(1337, 309, 65536, got.as_mut_ptr(), &mut n) }, sys::LCPC_OK);
    assert_eq!((n, got), (309, want));
}

#[test]
fn full_proof_matches_and_verifies() {
    let mut rng = rand::thread_rng();
    let len = 1usize << 16;
    let coeffs: Vec<Ft63> = (0..len).map(|_| Ft63::random(&mut rng)).collect();
    let cpu_enc = LigeroEncoding::<Ft63>::new(len);
    let gpu_enc = GpuLigeroEncoding::<Ft63>::new(len, &ctx());
    let cpu = LcCommit::<Blake3, _>::commit(&coeffs, &cpu_enc).unwrap();
    let gpu = GpuCommit::commit(&coeffs, &gpu_enc).unwrap();
    let root = cpu.get_root();
    assert_eq!(root.as_ref(), gpu.get_root().as_ref());
    // evaluation point and tensors as in lcpc-ligero-pc/src/tests.rs:234-242
    let x = Ft63::random(&mut rng);
    let (n_rows, n_per_row) = (cpu.get_n_rows(), cpu.get_n_per_row());
    let inner: Vec<Ft63> = std::iter::successors(Some(Ft63::ONE), |p| Some(*p * x)).take(n_per_row).collect();
    let xr = x * inner.last().unwrap();
    let outer: Vec<Ft63> = std::iter::successors(Some(Ft63::ONE), |p| Some(*p * xr)).take(n_rows).collect();
    let mut tr_cpu = Transcript::new(b"test transcript");
    tr_cpu.append_message(b"polycommit", root.as_ref());
    let pf_cpu = cpu.prove(&outer, &cpu_enc, &mut tr_cpu).unwrap();
    let mut tr_gpu = MirroredTranscript::new(b"test transcript");
    tr_gpu.append_message(b"polycommit", root.as_ref());
    let pf_gpu = prove_gpu(&gpu, &outer, &gpu_enc, &mut tr_gpu).unwrap();
    assert_eq!(pf_cpu.p_eval, pf_gpu.p_eval);
    assert_eq!(pf_cpu.p_random_vec, pf_gpu.p_random_vec);
    for (a, b) in pf_cpu.columns.iter().zip(&pf_gpu.columns) {
        assert_eq!(a.col, b.col);
        assert_eq!(a.path, b.path);
    }
    // lcpc-2d's own verifier accepts the GPU proof (the GPU encoding is an LcEncoding: verify re-encodes on the GPU)
    let mut tr_v = Transcript::new(b"test transcript");
    tr_v.append_message(b"polycommit", root.as_ref());
    let res = pf_gpu.verify(&root, &outer, &inner, &gpu_enc, &mut tr_v).unwrap();
    let direct = coeffs.iter().rev().fold(Ft63::ZERO, |acc, c| acc * x + c);
    assert_eq!(res, direct);
}
