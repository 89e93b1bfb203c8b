"""Fused CEMLP-block / EGCL path (csrc/block_fused.cu).  Placeholder switch until the kernels land."""
import os


def enabled(algebra) -> bool:
    return False
