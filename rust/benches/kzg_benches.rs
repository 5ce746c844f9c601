//! The reference's criterion benchmarks under the same names (reference `benches/kzg_benches.rs:46-126`), plus the
//! batched entry points the GPU path adds.  `KZG_TRUSTED_SETUP` points at `trusted_setup.txt`.
use criterion::{criterion_group, criterion_main, BenchmarkId, Criterion, Throughput};
use kzg_rust::*;
use rand::{rngs::ThreadRng, Rng};

fn random_field_element(rng: &mut ThreadRng) -> Bytes32 {
    let mut arr = [0u8; BYTES_PER_FIELD_ELEMENT];
    rng.fill(&mut arr[..]);
    arr[0] = 0;
    arr.into()
}

/// Canonical by construction: the top byte of every field element is zero (reference `benches/kzg_benches.rs:14-23`).
fn random_blob(rng: &mut ThreadRng) -> Blob {
    let mut bytes = vec![0u8; BYTES_PER_BLOB];
    rng.fill(&mut bytes[..]);
    for i in 0..FIELD_ELEMENTS_PER_BLOB {
        bytes[i * BYTES_PER_FIELD_ELEMENT] = 0;
    }
    Blob::from_bytes(&bytes).unwrap()
}

pub fn criterion_benchmark(c: &mut Criterion) {
    let max_count: usize = 64;
    let mut rng = rand::thread_rng();
    let setup = std::env::var("KZG_TRUSTED_SETUP").unwrap_or_else(|_| "trusted_setup.txt".to_string());
    let kzg_settings = Kzg::load_trusted_setup_file(setup).unwrap();

    let blobs: Vec<Blob> = (0..max_count).map(|_| random_blob(&mut rng)).collect();
    let commitments = Kzg::blob_to_kzg_commitment_batch(&blobs, &kzg_settings).unwrap();
    let proofs = Kzg::compute_blob_kzg_proof_batch(&blobs, &commitments, &kzg_settings).unwrap();
    let fields: Vec<Bytes32> = (0..max_count).map(|_| random_field_element(&mut rng)).collect();

    c.bench_function("blob_to_kzg_commitment", |b| {
        b.iter(|| Kzg::blob_to_kzg_commitment(&blobs[0], &kzg_settings))
    });
    c.bench_function("compute_kzg_proof", |b| {
        b.iter(|| Kzg::compute_kzg_proof(&blobs[0], &fields[0], &kzg_settings))
    });
    c.bench_function("compute_blob_kzg_proof", |b| {
        b.iter(|| Kzg::compute_blob_kzg_proof(&blobs[0], &commitments[0], &kzg_settings))
    });
    c.bench_function("verify_kzg_proof", |b| {
        b.iter(|| Kzg::verify_kzg_proof(&commitments[0], &fields[0], &fields[0], &proofs[0], &kzg_settings))
    });
    c.bench_function("verify_blob_kzg_proof", |b| {
        b.iter(|| Kzg::verify_blob_kzg_proof(&blobs[0], &commitments[0], &proofs[0], &kzg_settings))
    });

    let mut group = c.benchmark_group("verify_blob_kzg_proof_batch");
    for count in [1, 2, 4, 8, 16, 32, 64] {
        group.throughput(Throughput::Elements(count as u64));
        group.bench_with_input(BenchmarkId::from_parameter(count), &count, |b, &count| {
            b.iter(|| Kzg::verify_blob_kzg_proof_batch(&blobs[..count], &commitments[..count], &proofs[..count], &kzg_settings))
        });
    }
    group.finish();

    // the batched entry points
    let mut group = c.benchmark_group("blob_to_kzg_commitment_batch");
    for count in [1, 8, 64] {
        group.throughput(Throughput::Elements(count as u64));
        group.bench_with_input(BenchmarkId::from_parameter(count), &count, |b, &count| {
            b.iter(|| Kzg::blob_to_kzg_commitment_batch(&blobs[..count], &kzg_settings))
        });
    }
    group.finish();
    let mut group = c.benchmark_group("compute_blob_kzg_proof_batch");
    for count in [1, 8, 64] {
        group.throughput(Throughput::Elements(count as u64));
        group.bench_with_input(BenchmarkId::from_parameter(count), &count, |b, &count| {
            b.iter(|| Kzg::compute_blob_kzg_proof_batch(&blobs[..count], &commitments[..count], &kzg_settings))
        });
    }
    group.finish();
}

criterion_group!(benches, criterion_benchmark);
criterion_main!(benches);
