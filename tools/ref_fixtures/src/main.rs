//! Emits golden vectors from the real arkworks stack (ark-ff / ark-ec / ark-poly-commit at the versions the reference
//! depends on, `/root/reference/Cargo.toml:15-19,33-39`) as one JSON object on stdout.  Everything is derived from
//! `ark_std::test_rng()` -- the generator the reference's own tests use (`src/lib.rs:344`) -- so the Python side
//! (accumulation_b200/wire.py: TestRng, rand_fe, rand_point) can re-derive the inputs and only has to trust the outputs.
//!
//! Sections: "rng" (raw u64 / field / point draws: pins the generator + UniformRand rule), "msm" (VariableBaseMSM on the
//! scalar distributions of SURVEY 8d, incl. the scalar == 1 shortcut and zero filter), "pedersen" (commit with and without
//! randomizer), "coeffs" (SuccinctCheckPolynomial::compute_coeffs / evaluate), "ipa" (commit, open, check at degree 15 and
//! 1023 with the Poseidon sponge of the branch: every l, r, final_comm_key, c and the serialized proof bytes).
use ark_ec::msm::VariableBaseMSM;
use ark_ec::{AffineCurve, ProjectiveCurve};
use ark_ff::{to_bytes, PrimeField, UniformRand, Zero};
use ark_pallas::{Affine, Fq, Fr, Projective};
use ark_poly::polynomial::univariate::DensePolynomial;
use ark_poly::UVPolynomial;
use ark_poly_commit::ipa_pc::{InnerProductArgPC, SuccinctCheckPolynomial};
use ark_poly_commit::trivial_pc::PedersenCommitment;
use ark_poly_commit::{LabeledPolynomial, PolynomialCommitment};
use ark_serialize::CanonicalSerialize;
use ark_sponge::poseidon::PoseidonSponge;
use ark_std::rand::RngCore;

fn hex_ser<T: CanonicalSerialize>(t: &T) -> String {
    let mut v = Vec::new();
    t.serialize(&mut v).unwrap();
    hex::encode(v)
}
fn hex_fe<F: PrimeField>(f: &F) -> String {
    hex::encode(to_bytes![f.into_repr()].unwrap())       // 32 B little-endian canonical
}
fn jarr(items: Vec<String>) -> String {
    format!("[{}]", items.iter().map(|s| format!("\"{}\"", s)).collect::<Vec<_>>().join(","))
}

fn main() {
    let mut out = Vec::<String>::new();

    // ---- rng: raw stream, field draws, point draws (in this order from ONE test_rng())
    {
        let mut rng = ark_std::test_rng();
        let raw: Vec<String> = (0..8).map(|_| format!("{:016x}", rng.next_u64())).collect();
        let fr: Vec<String> = (0..8).map(|_| hex_fe(&Fr::rand(&mut rng))).collect();
        let fq: Vec<String> = (0..4).map(|_| hex_fe(&Fq::rand(&mut rng))).collect();
        let pts: Vec<String> = (0..6).map(|_| hex_ser(&Projective::rand(&mut rng).into_affine())).collect();
        out.push(format!("\"rng\":{{\"u64\":{},\"fr\":{},\"fq\":{},\"points\":{}}}", jarr(raw), jarr(fr), jarr(fq), jarr(pts)));
    }

    // ---- msm: bases and scalars from a fresh test_rng(); distributions by name
    {
        let mut rng = ark_std::test_rng();
        let n = 300usize;
        let bases: Vec<Affine> = (0..n).map(|_| Projective::rand(&mut rng).into_affine()).collect();
        let uniform: Vec<Fr> = (0..n).map(|_| Fr::rand(&mut rng)).collect();
        let a = Fr::rand(&mut rng);
        let mut cases = Vec::<String>::new();
        let mut run = |name: &str, sc: Vec<Fr>, cases: &mut Vec<String>| {
            let big: Vec<_> = sc.iter().map(|s| s.into_repr()).collect();
            let res = VariableBaseMSM::multi_scalar_mul(&bases[..sc.len().min(n)], &big).into_affine();
            cases.push(format!("{{\"name\":\"{}\",\"scalars\":{},\"result\":\"{}\"}}", name, jarr(sc.iter().map(hex_fe).collect()), hex_ser(&res)));
        };
        run("uniform", uniform.clone(), &mut cases);
        run("uniform_33", uniform[..33].to_vec(), &mut cases);
        run("uniform_31", uniform[..31].to_vec(), &mut cases);          // window rule: c = 3 below 32 pairs
        run("constant", vec![a; n], &mut cases);
        run("a_a_a_0", { let mut v = vec![a; n]; v[n - 1] = Fr::zero(); v }, &mut cases);
        run("ones", vec![Fr::from(1u64); n], &mut cases);               // the scalar == 1 shortcut
        run("zeros", vec![Fr::zero(); n], &mut cases);
        run("minus_one", vec![-Fr::from(1u64); n], &mut cases);
        run("longer_than_bases", { let mut v = uniform.clone(); v.extend_from_slice(&uniform[..7]); v }, &mut cases);
        out.push(format!("\"msm\":{{\"bases\":{},\"cases\":[{}]}}", jarr(bases.iter().map(hex_ser).collect()), cases.join(",")));
    }

    // ---- pedersen: setup is deterministic (hash-to-curve), so the generators themselves are fixtures
    {
        let mut rng = ark_std::test_rng();
        let pp = PedersenCommitment::<Affine>::setup(64);
        let ck = PedersenCommitment::<Affine>::trim(&pp, 64);
        let elems: Vec<Fr> = (0..64).map(|_| Fr::rand(&mut rng)).collect();
        let r = Fr::rand(&mut rng);
        let c0 = PedersenCommitment::<Affine>::commit(&ck, &elems, None);
        let c1 = PedersenCommitment::<Affine>::commit(&ck, &elems, Some(r));
        out.push(format!("\"pedersen\":{{\"generators\":{},\"hiding_generator\":\"{}\",\"elems\":{},\"randomizer\":\"{}\",\"commit\":\"{}\",\"commit_hiding\":\"{}\"}}",
                         jarr(ck.generators.iter().map(hex_ser).collect()), hex_ser(&ck.hiding_generator), jarr(elems.iter().map(hex_fe).collect()),
                         hex_fe(&r), hex_ser(&c0), hex_ser(&c1)));
    }

    // ---- succinct-check polynomial
    {
        let mut rng = ark_std::test_rng();
        let ch: Vec<Fr> = (0..6).map(|_| Fr::rand(&mut rng)).collect();
        let z = Fr::rand(&mut rng);
        let h = SuccinctCheckPolynomial(ch.clone());
        out.push(format!("\"coeffs\":{{\"challenges\":{},\"coeffs\":{},\"point\":\"{}\",\"evaluation\":\"{}\"}}",
                         jarr(ch.iter().map(hex_fe).collect()), jarr(h.compute_coeffs().iter().map(hex_fe).collect()), hex_fe(&z), hex_fe(&h.evaluate(z))));
    }

    // ---- IpaPC commit / open / check (the type alias of src/ipa_pc_as/mod.rs:33-39 with the Poseidon sponge of the branch)
    {
        type PC = InnerProductArgPC<Affine, blake2::Blake2s, DensePolynomial<Fr>, Fq, PoseidonSponge<Fq>>;
        let mut cases = Vec::<String>::new();
        for &degree in &[15usize, 1023] {
            let mut rng = ark_std::test_rng();
            let pp = PC::setup(degree, None, &mut rng).unwrap();
            let (ck, vk) = PC::trim(&pp, degree, 0, None).unwrap();
            let poly = DensePolynomial::<Fr>::rand(degree, &mut rng);
            let lp = LabeledPolynomial::new("p".to_string(), poly.clone(), None, None);
            let (comms, rands) = PC::commit(&ck, vec![&lp], None).unwrap();
            let point = Fr::rand(&mut rng);
            let value = poly.evaluate(&point);
            let proof = PC::open_individual_opening_challenges(&ck, vec![&lp], &comms, &point, &|_| Fr::from(1u64), &rands, None).unwrap();
            let ok = PC::check_individual_opening_challenges(&vk, &comms, &point, vec![value], &proof, &|_| Fr::from(1u64), None).unwrap();
            cases.push(format!("{{\"degree\":{},\"comm_key\":{},\"h\":\"{}\",\"s\":\"{}\",\"coeffs\":{},\"commitment\":\"{}\",\"point\":\"{}\",\"value\":\"{}\",\"l_vec\":{},\"r_vec\":{},\"final_comm_key\":\"{}\",\"c\":\"{}\",\"proof_bytes\":\"{}\",\"check\":{}}}",
                               degree, jarr(ck.comm_key.iter().map(hex_ser).collect()), hex_ser(&ck.h), hex_ser(&ck.s), jarr(poly.coeffs.iter().map(hex_fe).collect()),
                               hex_ser(comms[0].commitment()), hex_fe(&point), hex_fe(&value), jarr(proof.l_vec.iter().map(hex_ser).collect()),
                               jarr(proof.r_vec.iter().map(hex_ser).collect()), hex_ser(&proof.final_comm_key), hex_fe(&proof.c), hex_ser(&proof), ok));
        }
        out.push(format!("\"ipa\":[{}]", cases.join(",")));
    }
    println!("{{{}}}", out.join(","));
}
