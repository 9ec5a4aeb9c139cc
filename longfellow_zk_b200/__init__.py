"""longfellow_zk_b200 -- B200-native back end for the Longfellow-ZK prover hot
path (Reed-Solomon row encoding, SHA-256 Merkle column commitment, sumcheck
layer prover, Ligero prove) as hand-written sm_100a CUDA behind a C ABI
(include/longfellow_b200.h).  This package is the thin host-side mirror of the
reference's prover interfaces over that ABI."""
from .api import (Context, Circuit, LCH14ReedSolomonFactory, ReedSolomonFactory, MerkleCommitment, ZkProver, ZkVerifier,  # noqa: F401
                  FIELD_GF2_128, FIELD_P256, FIELD_BN254, FIELD_FP128, FIELD_GOLDILOCKS, LongfellowError)
