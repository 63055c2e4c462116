//! GPU path of `PolynomialBatch::from_coeffs / from_values` and `MerkleTree::new` (cargo feature `cuda`).
//!
//! Drop this file at `plonky2/src/fri/gpu.rs` of Lain-Iwakuro/Plonky2-Demo and apply `patches/*.patch`
//! (rust/plonky2-gpu-shim/README.md).  It uses only items that exist in the reference:
//!   * `TimingTree::push / pop`                         plonky2/src/util/timing.rs:80,107
//!   * the public fields of `MerkleTree` / `MerkleCap`  plonky2/src/hash/merkle_tree.rs:18,39-55
//!   * `GenericHashOut::from_bytes`                     plonky2/src/plonk/config.rs:22, hash_types.rs:91-101
//!   * `Field::from_canonical_u64`, `PrimeField64::to_noncanonical_u64`
//! and the C ABI of libpcs.so through the generated `pcs-sys` crate (include/pcs.h).
//!
//! Dispatch: the engine exists for F = GoldilocksField and H = PoseidonHash only.  `applicable::<F, H>()` compares
//! type names (no `'static` bound on `Hasher`), every other instantiation keeps the reference's CPU code, and
//! inside the GPU path a `Vec<F>` is read as `*const u64` -- sound because `GoldilocksField` is
//! `#[repr(transparent)]` over `u64` (field/src/goldilocks_field.rs:23-25).
//!
//! Storage: "compatible mode".  The rows (`MerkleTree.leaves`) and `digests` are copied back once per commit, so every
//! consumer that indexes the public fields (`get_lde_values`, `compute_quotient_polys`, `fri_prover_query_round`,
//! serialisation) works unchanged.  The device-resident batch handle is kept next to them (`MerkleTree.device`) for the
//! consumers that have a device entry point (`OpeningSet::new` -> pcs_batch_eval_ext, `prove_openings` ->
//! pcs_fri_final_poly, bulk natural-order rows -> pcs_batch_lde_natural); dropping the tree frees it.
use alloc::sync::Arc;
use alloc::vec;
use alloc::vec::Vec;
use core::ffi::CStr;

use pcs_sys as sys;

use crate::field::goldilocks_field::GoldilocksField;
use crate::field::polynomial::{PolynomialCoeffs, PolynomialValues};
use crate::field::types::{Field, PrimeField64};
use crate::hash::hash_types::RichField;
use crate::hash::merkle_tree::{MerkleCap, MerkleTree};
use crate::hash::poseidon::PoseidonHash;
use crate::plonk::config::{GenericHashOut, Hasher};
use crate::util::log2_strict;
use crate::util::timing::TimingTree;

/// Commitments with fewer leaves than this stay on the reference's CPU path: a device commit has a fixed cost of launches and
/// one synchronisation; measured through the C ABI (profiles/r02_small_commits.md) a 135-column commit costs 329 us at 64 leaves
/// (CPU port: 272 us) and 350 us at 128 leaves (CPU port: 486 us), so the device path takes over from 128 leaves up.
pub const MIN_GPU_LEAVES: usize = 1 << 7;

/// Owner of a `pcs_batch*` (device-resident LDE rows, digests, cap, coefficients).
pub struct DeviceBatch(pub(crate) *mut sys::pcs_batch);

// The engine serialises calls per context (include/pcs.h, "Contexts"); the handle itself is an opaque pointer.
unsafe impl Send for DeviceBatch {}
unsafe impl Sync for DeviceBatch {}

impl Drop for DeviceBatch {
    fn drop(&mut self) {
        unsafe { sys::pcs_batch_free(self.0) }
    }
}

impl core::fmt::Debug for DeviceBatch {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        write!(f, "DeviceBatch({:p})", self.0)
    }
}

/// `MerkleTree` derives `Eq`: two trees are equal when their host-visible contents are; the device copy is a cache.
impl PartialEq for DeviceBatch {
    fn eq(&self, _other: &Self) -> bool {
        true
    }
}
impl Eq for DeviceBatch {}

/// Is this instantiation the one the engine implements?
pub(crate) fn applicable<F: RichField, H: Hasher<F>>() -> bool {
    core::any::type_name::<F>() == core::any::type_name::<GoldilocksField>()
        && core::any::type_name::<H>() == core::any::type_name::<PoseidonHash>()
}

fn last_error() -> alloc::string::String {
    unsafe { CStr::from_ptr(sys::pcs_last_error()) }
        .to_string_lossy()
        .into_owned()
}

/// The reference panics on invalid input (oracle.rs:76,114; merkle_tree.rs:136-142; util/src/lib.rs:37); the C ABI
/// returns a status and keeps the reference's message.
fn check(rc: core::ffi::c_int) {
    if rc != 0 {
        panic!("{}", last_error());
    }
}

fn hashes_from_words<F: RichField, H: Hasher<F>>(words: &[u64]) -> Vec<H::Hash> {
    // HashOut<F> == [u64; 4], canonical little-endian words (hash_types.rs:83-101)
    words
        .chunks_exact(4)
        .map(|h| {
            let mut bytes = [0u8; 32];
            for (k, x) in h.iter().enumerate() {
                bytes[8 * k..8 * k + 8].copy_from_slice(&x.to_le_bytes());
            }
            H::Hash::from_bytes(&bytes)
        })
        .collect()
}

/// Everything `from_coeffs` produces besides `polynomials` (oracle.rs:76-97), from one device commit.
pub(crate) struct GpuCommit<F: RichField, H: Hasher<F>> {
    pub merkle_tree: MerkleTree<F, H>,
    /// [0] IFFT, [1] FFT + blinding, [2] transpose LDEs (always 0: fused away), [3] leaf hashing, [4] node levels, in ms
    pub phase_ms: [f32; 5],
}

/// Reports the device phases under the reference's scope names (oracle.rs:51-89).  `TimingTree` has no way to record
/// an externally measured duration, so the scopes are opened and closed around nothing and the measured CUDA-event
/// times are logged next to them at the same level.
fn report_timing(timing: &mut TimingTree, from_values: bool, ms: &[f32; 5]) {
    if from_values {
        timing.push("IFFT", log::Level::Debug);
        timing.pop();
    }
    for name in ["FFT + blinding", "transpose LDEs", "build Merkle tree"] {
        timing.push(name, log::Level::Debug);
        timing.pop();
    }
    log::debug!(
        "pcs (GPU): IFFT {:.3} ms, FFT + blinding {:.3} ms, transpose LDEs {:.3} ms (fused), build Merkle tree {:.3} ms",
        ms[0],
        ms[1],
        ms[2],
        ms[3] + ms[4]
    );
}

fn read_back<F: RichField, H: Hasher<F>>(batch: *mut sys::pcs_batch, cap_words: &[u64]) -> MerkleTree<F, H> {
    let (mut n, mut width, mut n_digests, mut cap_height) = (0usize, 0usize, 0usize, 0u32);
    check(unsafe { sys::pcs_batch_shape(batch, &mut n, &mut width, &mut n_digests, &mut cap_height) });
    // merkle_tree.leaves: N heap rows, as the reference's transpose() makes them (plonky2/src/util/mod.rs:22-28)
    let mut flat = vec![0u64; n * width];
    check(unsafe { sys::pcs_batch_leaves(batch, 0, n, flat.as_mut_ptr()) });
    let leaves: Vec<Vec<F>> = flat
        .chunks_exact(width)
        .map(|row| row.iter().map(|&x| F::from_canonical_u64(x)).collect())
        .collect();
    let mut dig = vec![0u64; 4 * n_digests];
    check(unsafe { sys::pcs_batch_digests(batch, dig.as_mut_ptr()) });
    MerkleTree {
        leaves,
        digests: hashes_from_words::<F, H>(&dig),
        cap: MerkleCap(hashes_from_words::<F, H>(cap_words)),
        device: Some(Arc::new(DeviceBatch(batch))),
    }
}

/// Body of `PolynomialBatch::from_coeffs` (oracle.rs:68-98) for F = Goldilocks, H = Poseidon.
pub(crate) fn commit_from_coeffs<F: RichField, H: Hasher<F>>(
    polynomials: &[PolynomialCoeffs<F>],
    rate_bits: usize,
    blinding: bool,
    cap_height: usize,
    timing: &mut TimingTree,
) -> GpuCommit<F, H> {
    debug_assert!(applicable::<F, H>());
    let degree = polynomials[0].len(); // panics on an empty batch like oracle.rs:76
    for p in polynomials {
        assert_eq!(p.len(), degree, "Polynomial degrees inconsistent"); // oracle.rs:114
    }
    let lg_d = log2_strict(degree); // "Not a power of two: {n}" (util/src/lib.rs:37)
    let n = degree << rate_bits;
    // GoldilocksField is repr(transparent) over u64: a Vec<F> is a *const u64 (checked by `applicable`)
    let ptrs: Vec<*const u64> = polynomials.iter().map(|p| p.coeffs.as_ptr() as *const u64).collect();
    // SALT_SIZE random columns when blinding (oracle.rs:26,119-123); drawn here so the commit stays a function of its inputs
    let salts: Vec<Vec<F>> = if blinding {
        (0..crate::fri::oracle::SALT_SIZE).map(|_| F::rand_vec(n)).collect()
    } else {
        Vec::new()
    };
    let salt_ptrs: Vec<*const u64> = salts.iter().map(|s| s.as_ptr() as *const u64).collect();
    let mut cap_words = vec![0u64; 4 << cap_height];
    let mut batch: *mut sys::pcs_batch = core::ptr::null_mut();
    check(unsafe {
        sys::pcs_commit_from_coeffs(
            ptrs.as_ptr(),
            ptrs.len(),
            lg_d as u32,
            rate_bits as u32,
            cap_height as u32, // "cap_height={} should be at most log2(leaves.len())={}" comes back as the error text
            if blinding { salt_ptrs.as_ptr() } else { core::ptr::null() },
            salt_ptrs.len(),
            sys::PCS_KEEP_COEFFS, // OpeningSet::new / prove_openings read `polynomials` on the device
            cap_words.as_mut_ptr(),
            &mut batch,
        )
    });
    let mut ms = [0f32; 5];
    check(unsafe { sys::pcs_batch_timings(batch, ms.as_mut_ptr()) });
    report_timing(timing, false, &ms);
    GpuCommit {
        merkle_tree: read_back::<F, H>(batch, &cap_words),
        phase_ms: ms,
    }
}

/// Body of `PolynomialBatch::from_values` (oracle.rs:43-65): the IFFT runs on the device too and the coefficients come back
/// into freshly allocated `PolynomialCoeffs` (the reference keeps them as `polynomials`).
pub(crate) fn commit_from_values<F: RichField, H: Hasher<F>>(
    values: &[PolynomialValues<F>],
    rate_bits: usize,
    blinding: bool,
    cap_height: usize,
    timing: &mut TimingTree,
) -> (Vec<PolynomialCoeffs<F>>, GpuCommit<F, H>) {
    debug_assert!(applicable::<F, H>());
    let degree = values[0].len();
    for v in values {
        assert_eq!(v.len(), degree, "Polynomial degrees inconsistent");
    }
    let lg_d = log2_strict(degree);
    let n = degree << rate_bits;
    let ptrs: Vec<*const u64> = values.iter().map(|v| v.values.as_ptr() as *const u64).collect();
    let mut coeffs: Vec<PolynomialCoeffs<F>> = (0..values.len())
        .map(|_| PolynomialCoeffs::new(vec![F::ZERO; degree]))
        .collect();
    let out_ptrs: Vec<*mut u64> = coeffs.iter_mut().map(|p| p.coeffs.as_mut_ptr() as *mut u64).collect();
    let salts: Vec<Vec<F>> = if blinding {
        (0..crate::fri::oracle::SALT_SIZE).map(|_| F::rand_vec(n)).collect()
    } else {
        Vec::new()
    };
    let salt_ptrs: Vec<*const u64> = salts.iter().map(|s| s.as_ptr() as *const u64).collect();
    let mut cap_words = vec![0u64; 4 << cap_height];
    let mut batch: *mut sys::pcs_batch = core::ptr::null_mut();
    check(unsafe {
        sys::pcs_commit_from_values(
            ptrs.as_ptr(),
            ptrs.len(),
            lg_d as u32,
            rate_bits as u32,
            cap_height as u32,
            if blinding { salt_ptrs.as_ptr() } else { core::ptr::null() },
            salt_ptrs.len(),
            0, // from_values keeps the coefficients on the device by itself
            out_ptrs.as_ptr(),
            cap_words.as_mut_ptr(),
            &mut batch,
        )
    });
    let mut ms = [0f32; 5];
    check(unsafe { sys::pcs_batch_timings(batch, ms.as_mut_ptr()) });
    report_timing(timing, true, &ms);
    (
        coeffs,
        GpuCommit {
            merkle_tree: read_back::<F, H>(batch, &cap_words),
            phase_ms: ms,
        },
    )
}

/// `MerkleTree::new(leaves, cap_height)` (merkle_tree.rs:135-166) for the FRI commit-phase trees (fri/prover.rs:81-87) and
/// every other caller: rows are flattened, hashed on the device, `digests` / `cap` come back in the reference layout.
pub(crate) fn merkle_tree_new<F: RichField, H: Hasher<F>>(leaves: Vec<Vec<F>>, cap_height: usize) -> MerkleTree<F, H> {
    debug_assert!(applicable::<F, H>());
    let n = leaves.len();
    let log2_leaves_len = log2_strict(n);
    assert!(
        cap_height <= log2_leaves_len,
        "cap_height={} should be at most log2(leaves.len())={}",
        cap_height,
        log2_leaves_len
    );
    let width = leaves[0].len();
    let mut flat = Vec::with_capacity(n * width);
    for row in &leaves {
        assert_eq!(row.len(), width, "ragged leaves are not supported by the device path");
        flat.extend(row.iter().map(|x| x.to_noncanonical_u64()));
    }
    let n_digests = 2 * (n - (1 << cap_height));
    let mut dig = vec![0u64; 4 * n_digests];
    let mut cap = vec![0u64; 4 << cap_height];
    check(unsafe {
        sys::pcs_merkle_build(flat.as_ptr(), n, width, cap_height as u32, dig.as_mut_ptr(), cap.as_mut_ptr())
    });
    MerkleTree {
        leaves,
        digests: hashes_from_words::<F, H>(&dig),
        cap: MerkleCap(hashes_from_words::<F, H>(&cap)),
        device: None,
    }
}

// ------------------------------------------------------------------------------------------------
// The batch's consumers in prove(): openings and the opening proof's composition polynomial
// (SURVEY 8f N3 / N2).  D = 2 only: F::Extension = QuadraticExtension<GoldilocksField> = [u64; 2].
// ------------------------------------------------------------------------------------------------
use crate::field::extension::{Extendable, FieldExtension};
use crate::fri::oracle::PolynomialBatch;
use crate::fri::structure::FriInstanceInfo;
use crate::plonk::config::GenericConfig;

fn ext_to_words<F: RichField + Extendable<D>, const D: usize>(x: F::Extension) -> [u64; 2] {
    let a = x.to_basefield_array();
    [a[0].to_canonical_u64(), a[1].to_canonical_u64()]
}

fn ext_from_words<F: RichField + Extendable<D>, const D: usize>(w: &[u64]) -> F::Extension {
    let mut a = [F::ZERO; D];
    a[0] = F::from_canonical_u64(w[0]);
    a[1] = F::from_canonical_u64(w[1]);
    F::Extension::from_basefield_array(a)
}

/// Does this batch carry a device copy (with coefficients) that the opening kernels can read?
pub(crate) fn has_device<F: RichField + Extendable<D>, C: GenericConfig<D, F = F>, const D: usize>(
    c: &PolynomialBatch<F, C, D>,
) -> bool {
    D == 2 && c.merkle_tree.device.is_some()
}

/// `OpeningSet::new`'s `eval_commitment(z, c)`: `c.polynomials[j].to_extension().eval(z)` for every j
/// (plonk/proof.rs:316-322) in one pass over the coefficients in HBM (pcs_batch_eval_ext, 3.8 TB/s).
pub(crate) fn eval_commitment<F: RichField + Extendable<D>, C: GenericConfig<D, F = F>, const D: usize>(
    z: F::Extension,
    c: &PolynomialBatch<F, C, D>,
) -> Vec<F::Extension> {
    let dev = c.merkle_tree.device.as_ref().expect("has_device() was checked");
    let point = ext_to_words::<F, D>(z);
    let mut out = vec![0u64; 2 * c.polynomials.len()];
    check(unsafe { sys::pcs_batch_eval_ext(dev.0, point.as_ptr(), out.as_mut_ptr()) });
    out.chunks_exact(2).map(|w| ext_from_words::<F, D>(w)).collect()
}

/// The final polynomial of `PolynomialBatch::prove_openings` (fri/oracle.rs:171-200): per batch the alpha-reduction of its
/// polynomials (`ReducingFactor::reduce_polys_base`), the quotient by (X - z_i) padded back to a power of two, and
/// `shift_poly` -- computed on the device over the oracles' coefficients, d extension coefficients come back.
pub(crate) fn final_poly<F: RichField + Extendable<D>, C: GenericConfig<D, F = F>, const D: usize>(
    instance: &FriInstanceInfo<F, D>,
    oracles: &[&PolynomialBatch<F, C, D>],
    alpha: F::Extension,
) -> PolynomialCoeffs<F::Extension> {
    let handles: Vec<*const sys::pcs_batch> = oracles
        .iter()
        .map(|o| o.merkle_tree.device.as_ref().expect("has_device() was checked").0 as *const sys::pcs_batch)
        .collect();
    let mut points = Vec::with_capacity(2 * instance.batches.len());
    let mut batch_len = Vec::with_capacity(instance.batches.len());
    let (mut oracle_index, mut poly_index) = (Vec::new(), Vec::new());
    for b in &instance.batches {
        points.extend_from_slice(&ext_to_words::<F, D>(b.point));
        batch_len.push(b.polynomials.len());
        for p in &b.polynomials {
            oracle_index.push(p.oracle_index as u32);
            poly_index.push(p.polynomial_index as u32);
        }
    }
    let alpha_w = ext_to_words::<F, D>(alpha);
    let mut poly: *mut sys::pcs_ext_poly = core::ptr::null_mut();
    check(unsafe {
        sys::pcs_fri_final_poly(
            handles.as_ptr(),
            handles.len(),
            instance.batches.len(),
            points.as_ptr(),
            batch_len.as_ptr(),
            oracle_index.as_ptr(),
            poly_index.as_ptr(),
            alpha_w.as_ptr(),
            &mut poly,
        )
    });
    let mut len = 0usize;
    check(unsafe { sys::pcs_ext_poly_len(poly, &mut len) });
    let mut words = vec![0u64; 2 * len];
    check(unsafe { sys::pcs_ext_poly_read(poly, words.as_mut_ptr()) });
    unsafe { sys::pcs_ext_poly_free(poly) };
    PolynomialCoeffs::new(words.chunks_exact(2).map(|w| ext_from_words::<F, D>(w)).collect())
}
