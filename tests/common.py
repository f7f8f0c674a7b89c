"""Shared helpers for the parity tests: seeded synthetic world + checker handles."""
from __future__ import annotations

import os
import numpy as np

from rawhash_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def model_for(kind="r9.4", tmpdir="/tmp"):
    p = synth.model_path(kind)
    if p is None:
        k = 6 if kind == "r9.4" else 9
        p = os.path.join(tmpdir, f"synthetic_{kind}.model")
        if not os.path.isfile(p):
            synth.write_synthetic_model(p, k)
    return p


class World:
    """Genome + reads + everything needed to build the same index in every implementation."""

    def __init__(self, n_contigs=2, genome_len=400_000, n_reads=24, read_bp=3000, seed=1, kind="r9.4",
                 sample_rate=4000.0, bp_per_sec=450.0, tmpdir="/tmp"):
        self.kind = kind
        self.k = 6 if kind == "r9.4" else 9
        self.model = model_for(kind, tmpdir)
        self.means, self.stdv = synth.load_model_pa(self.model, self.k)
        self.genome = synth.make_genome(n_contigs, genome_len, seed=seed)
        self.fasta = os.path.join(tmpdir, f"rh_world_{seed}_{n_contigs}_{genome_len}.fa")
        synth.write_fasta(self.fasta, self.genome)
        self.reads = synth.make_reads(self.genome, n_reads, read_bp, self.k, self.means, self.stdv,
                                      sample_rate=sample_rate, bp_per_sec=bp_per_sec, seed=seed + 100)
        self.names = self.reads["names"]

    def pa(self, i):
        return synth.raw_to_pa(self.reads["raw"][i], synth.OFFSET, synth.RANGE, synth.DIGITISATION)

    def genome_strings(self):
        gs = synth.genome_to_strings(self.genome)
        return [n for n, _ in gs], [s for _, s in gs]


def tap_equal(a, b, keys=("events", "seeds", "anchors", "u", "chain_a", "prev_a", "regs")):
    """Compare two per-chunk tap lists bit for bit; returns list of (chunk, key) mismatches."""
    bad = []
    if len(a) != len(b):
        return [("n_chunks", len(a), len(b))]
    for ci, (x, y) in enumerate(zip(a, b)):
        for k in keys:
            ax, ay = x[k], y[k]
            if k == "events":
                ax, ay = ax.view(np.uint32), ay.view(np.uint32)
            if ax.shape != ay.shape or not np.array_equal(ax, ay):
                bad.append((ci, k, ax.shape, ay.shape))
                break
    return bad
