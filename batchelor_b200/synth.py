"""Synthetic Gaussian-mixture batches of the shapes BASELINE.md section 3 names (used by bench.py and the tests).

Values are generated in float32 and widened to float64, so the GPU path and the fp64 CPU oracle consume bit-identical
inputs.  Seeds: 7 for the mixture centres / loadings, 1000+b for batch b.
"""
from __future__ import annotations

import numpy as np


def pc_batches(nbatches: int, ncells, d: int = 50, ncomp: int = 32, shift_norm: float = 5.0, seed_offset: int = 0):
    """PC-space batches: `ncomp` components with centres ~ N(0, 3^2 I_d), unit noise, Dirichlet(2) weights per batch,
    batch b > 0 shifted by a random vector of norm `shift_norm`."""
    if np.isscalar(ncells):
        ncells = [int(ncells)] * nbatches
    crng = np.random.default_rng(7)
    centres = crng.normal(scale=3.0, size=(ncomp, d)).astype(np.float32)
    out = []
    for b in range(nbatches):
        rng = np.random.default_rng(1000 + b + seed_offset)
        w = rng.dirichlet(np.full(ncomp, 2.0))
        comp = rng.choice(ncomp, size=ncells[b], p=w)
        x = centres[comp] + rng.standard_normal(size=(ncells[b], d), dtype=np.float32)
        if b > 0:
            s = rng.standard_normal(size=d).astype(np.float32)
            s *= np.float32(shift_norm) / np.linalg.norm(s)
            x = x + s
        out.append(x.astype(np.float32).astype(np.float64))
    return out


def gene_batches(nbatches: int, ncells, G: int = 2000, latent: int = 20, ncomp: int = 32, seed_offset: int = 0):
    """Gene-space batches [G x cells]: latent mixture -> fixed N(0, 1/sqrt(latent)) loading + N(0,1) noise."""
    if np.isscalar(ncells):
        ncells = [int(ncells)] * nbatches
    crng = np.random.default_rng(7)
    loading = (crng.normal(size=(latent, G)) / np.sqrt(latent)).astype(np.float32)
    lat = pc_batches(nbatches, ncells, d=latent, ncomp=ncomp, shift_norm=2.0, seed_offset=seed_offset)
    out = []
    for b in range(nbatches):
        rng = np.random.default_rng(5000 + b + seed_offset)
        x = lat[b].astype(np.float32) @ loading + rng.standard_normal(size=(ncells[b], G), dtype=np.float32)
        out.append(np.ascontiguousarray(x.T).astype(np.float32).astype(np.float64))
    return out
