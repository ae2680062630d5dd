"""zigz_b200 — B200-native proving hot path of the zigz zkVM (BabyBear MLE kernels, sumcheck rounds, Lasso
lookup prover, SHA3-256 Merkle commitments) behind a C ABI.  See DESIGN.md / INTEGRATION.md.

The package holds the sm_100a kernels + C ABI (csrc/, built into libzigz_b200.so) and this thin Python mirror of
the reference's prover-side API (src/lib.zig:9-17).  Importing the package never compiles anything and never falls
back to a CPU path: creating a `Context` without a CUDA device raises `ZigzError(NoCudaDevice)`.
"""
from ._cabi import LIB_PATH, ZigzError, build, declared_prototypes, lib  # noqa: F401
from .api import (  # noqa: F401
    BABYBEAR_P, TABLE_ADD, TABLE_AND, TABLE_XOR, CommitmentScheme, Context, FiatShamirTranscript, LassoProof, LassoProver,
    MerkleOpeningProof, MerklePath, MerkleTree, Multilinear, OpeningProof, PolynomialCommitment, ProductSumcheckProver, EqProductSumcheckProver, SimpleMerkleTree,
    SumcheckProof, SumcheckProver, CommitmentOpenings, WITNESS_COLUMNS, generate_commitments, witness_pack, witness_pack_commit, prove_from_trace, verify_proof, build_add_table, build_and_table, build_xor_table, eval_univariate_coeffs, sha3_256,
)
