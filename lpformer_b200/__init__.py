"""lpformer_b200 — B200-native (sm_100a) implementation of LPFormer's per-link
pairwise-encoding path behind the reference's LinkTransformer module API."""
from .model import MLP, GCN, LinkTransformer, mlp_score  # noqa: F401
from .graph import CSR, csr_from_coo, csr_from_sparse, gcn_normalise  # noqa: F401

__all__ = ["LinkTransformer", "mlp_score", "MLP", "GCN", "CSR", "csr_from_coo", "csr_from_sparse", "gcn_normalise"]
