"""Loader for tests/golden/*.npz (fixtures generated from the compiled reference by
tests/golden/make_golden.py).  Rebuilds the exact inputs of a case from the fixture alone."""
from __future__ import annotations

import hashlib
import os

import numpy as np

from rawhash_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["r94_sensitive", "r94_fast", "r10_sensitive"]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def unpack2(b: np.ndarray, n: int) -> np.ndarray:
    out = np.empty((len(b), 4), dtype=np.uint8)
    for j in range(4):
        out[:, j] = (b >> (2 * j)) & 3
    return out.reshape(-1)[:n]


class GoldenCase:
    def __init__(self, name: str, tmpdir: str = "/tmp"):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
        z = self.z
        self.preset = str(z["preset"]); self.r10 = bool(int(z["r10"])); self.k = int(z["k"])
        self.sample_rate = int(z["sample_rate"]); self.bp_per_sec = int(z["bp_per_sec"])
        self.model = os.path.join(tmpdir, f"rh_golden_model_{self.k}.tsv")
        synth.write_synthetic_model(self.model, self.k, int(z["model_seed"]))
        self.genome = [(str(z[f"contig{i}_name"]), unpack2(z[f"contig{i}_2bit"], int(z[f"contig{i}_len"]))) for i in range(int(z["n_contigs"]))]
        self.fasta = os.path.join(tmpdir, f"rh_golden_{name}.fa")
        synth.write_fasta(self.fasta, self.genome)
        self.names = [str(s) for s in z["names"]]
        self.raws = [z[f"raw{i}"] for i in range(len(self.names))]
        self.cal = (float(z["offset"]), float(z["range"]), float(z["digitisation"]))
        self.paf = str(z["paf"])
        self.mid_occ = int(z["mid_occ"])

    def non_default_sampling(self):
        return self.sample_rate != 4000 or self.bp_per_sec != 450

    def genome_strings(self):
        gs = synth.genome_to_strings(self.genome)
        return [n for n, _ in gs], [s for _, s in gs]

    def chunks(self, i):
        """Expected per-chunk stage outputs of read i (anchors only as a hash)."""
        z = self.z
        out = []
        for ci in range(int(z[f"nchunks{i}"])):
            pre = f"r{i}c{ci}_"
            out.append({k: z[pre + k] for k in ("cnt", "events", "seeds", "u", "chain_a", "prev_a", "regs")} | {"anchors_sha": str(z[pre + "anchors_sha"])})
        return out


def compare_tap(got, exp):
    """got: list of per-chunk dicts from a tap_read; exp: GoldenCase.chunks(i).  Returns mismatch list."""
    bad = []
    if len(got) != len(exp):
        return [("n_chunks", len(got), len(exp))]
    for ci, (g, e) in enumerate(zip(got, exp)):
        if not np.array_equal(g["cnt"][1:], e["cnt"][1:]):  # cnt[0] (n_sig) is internal to detect_events: the reference tap leaves it 0
            bad.append((ci, "cnt", g["cnt"].tolist(), e["cnt"].tolist()))
            continue
        if not np.array_equal(g["events"].view(np.uint32), e["events"]):
            bad.append((ci, "events"))
        for k in ("seeds", "u", "chain_a", "prev_a", "regs"):
            if g[k].shape != e[k].shape or not np.array_equal(g[k], e[k]):
                bad.append((ci, k))
        if sha(g["anchors"]) != e["anchors_sha"]:
            bad.append((ci, "anchors"))
    return bad
