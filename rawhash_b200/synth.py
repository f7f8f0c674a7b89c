"""Seeded synthetic inputs for tests and bench (SURVEY.md §8d recipe).

No genome or signal files exist on the build/GPU boxes and there is no network, so every
input is generated: an i.i.d. ACGT genome and nanopore-like raw int16 reads sampled from it
through a k-mer pore model (level mean per k-mer held for an exponential dwell, Gaussian
noise, ADC quantisation with digitisation 8192 / range 1400 / offset 10).

The pore model is the ONT table the reference vendors (extern/kmer_models); build() stages a
copy under data/models/ (git-ignored).  When it is absent a seeded synthetic table of the
same shape is used and reported as such.
"""
from __future__ import annotations

import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

DIGITISATION = 8192.0
RANGE = 1400.0
OFFSET = 10.0


def model_path(kind: str = "r9.4") -> str | None:
    """Path of the staged ONT k-mer model, or None."""
    name = {"r9.4": "r9.4_6mer.model", "r10.4.1": "r10.4.1_9mer.txt"}[kind]
    for d in (os.path.join(_ROOT, "data", "models"),
              "/root/reference/extern/kmer_models/legacy/legacy_r9.4_180mv_450bps_6mer" if kind == "r9.4"
              else "/root/reference/extern/kmer_models/dna_r10.4.1_e8.2_400bps"):
        for cand in (name, "template_median68pA.model", "9mer_levels_v1.txt"):
            p = os.path.join(d, cand)
            if os.path.isfile(p):
                return p
    return None


def write_synthetic_model(path: str, k: int = 6, seed: int = 7) -> str:
    """Seeded stand-in for an ONT model file (same columns as the R9.4 table)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = 4 ** k
    mean = rng.normal(90.0, 12.0, n).clip(55.0, 130.0)
    stdv = np.full(n, 1.5)
    with open(path, "w") as f:
        f.write("kmer\tlevel_mean\tlevel_stdv\n")
        for i in range(n):
            kmer = "".join("ACGT"[(i >> (2 * (k - 1 - j))) & 3] for j in range(k))
            f.write(f"{kmer}\t{mean[i]:.6f}\t{stdv[i]:.6f}\n")
    return path


def load_model_pa(path: str, k: int, r10_mean: float = 90.0, r10_sd: float = 13.0):
    """Returns (level_mean_pA[4^k], level_stdv_pA[4^k]) indexed by 2-bit packed k-mer (A=0..T=3).

    The R10.4.1 table holds normalised levels without a stdv column; it is de-normalised as
    90 + 13*level pA with a fixed 2 pA noise."""
    means = np.zeros(4 ** k, dtype=np.float64)
    stdv = np.zeros(4 ** k, dtype=np.float64)
    i = 0
    with open(path) as f:
        for line in f:
            if line.startswith("kmer"):
                continue
            t = line.rstrip("\n").split("\t")
            if len(t) < 2:
                continue
            means[i] = float(t[1])
            stdv[i] = float(t[2]) if len(t) > 2 else -1.0
            i += 1
    assert i == 4 ** k, (i, k)
    if stdv[0] < 0:  # normalised table (R10)
        means = r10_mean + r10_sd * means
        stdv[:] = 2.0
    return means, stdv


def make_genome(n_contigs: int, total_len: int, seed: int = 1):
    """List of (name, uint8 array of 0..3) contigs, i.i.d. uniform."""
    rng = np.random.Generator(np.random.PCG64(seed))
    per = total_len // n_contigs
    out = []
    for c in range(n_contigs):
        ln = per if c < n_contigs - 1 else total_len - per * (n_contigs - 1)
        out.append((f"chr{c + 1}", rng.integers(0, 4, ln, dtype=np.uint8)))
    return out


def genome_to_strings(genome):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    return [(name, lut[seq].tobytes().decode()) for name, seq in genome]


def write_fasta(path: str, genome, width: int = 80):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as f:
        for name, seq in genome:
            f.write(b">" + name.encode() + b"\n")
            s = lut[seq]
            for i in range(0, len(s), width * 1000):
                blk = s[i:i + width * 1000]
                nfull = len(blk) // width
                if nfull:
                    a = np.empty((nfull, width + 1), dtype=np.uint8)
                    a[:, :width] = blk[:nfull * width].reshape(nfull, width)
                    a[:, width] = 10
                    f.write(a.tobytes())
                if len(blk) % width:
                    f.write(blk[nfull * width:].tobytes() + b"\n")


def _kmer_codes(seq: np.ndarray, k: int) -> np.ndarray:
    """2-bit packed k-mer code at every position (len - k + 1 codes)."""
    n = len(seq) - k + 1
    code = np.zeros(n, dtype=np.int64)
    for j in range(k):
        code = (code << 2) | seq[j:j + n].astype(np.int64)
    return code


def make_reads(genome, n_reads: int, read_len_bp: int, k: int, means, stdv,
               sample_rate: float = 4000.0, bp_per_sec: float = 450.0, seed: int = 2,
               max_samples: int | None = None):
    """Synthetic raw reads.

    Returns dict with: raw (list of int16 arrays), names, truth (contig idx, start, strand),
    offset/range/digitisation (float64 arrays)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    scale = RANGE / DIGITISATION
    mean_dwell = sample_rate / bp_per_sec
    raws, names, truth = [], [], []
    lens = np.array([len(s) for _, s in genome], dtype=np.float64)
    prob = lens / lens.sum()
    for r in range(n_reads):
        ci = int(rng.choice(len(genome), p=prob))
        seq = genome[ci][1]
        ln = min(read_len_bp, len(seq))
        st = int(rng.integers(0, len(seq) - ln + 1))
        strand = int(rng.integers(0, 2))
        frag = seq[st:st + ln]
        if strand:
            frag = (3 - frag)[::-1]
        codes = _kmer_codes(frag, k)
        dwell = np.maximum(1, np.rint(rng.exponential(mean_dwell, len(codes)))).astype(np.int64)
        if max_samples is not None:
            cs = np.cumsum(dwell)
            keep = int(np.searchsorted(cs, max_samples)) + 1
            codes, dwell = codes[:keep], dwell[:keep]
        lv = np.repeat(means[codes], dwell)
        sd = np.repeat(stdv[codes], dwell)
        pa = lv + rng.standard_normal(len(lv)) * sd
        raw = np.rint(pa / scale - OFFSET).clip(-32768, 32767).astype(np.int16)
        raws.append(raw)
        names.append(f"read_{r:07d}")
        truth.append((ci, st, strand))
    n = len(raws)
    return {
        "raw": raws, "names": names, "truth": truth,
        "offset": np.full(n, OFFSET), "range": np.full(n, RANGE), "digitisation": np.full(n, DIGITISATION),
    }


def raw_to_pa(raw: np.ndarray, offset: float, rng_: float, digitisation: float) -> np.ndarray:
    """Host restatement of the slow5 pA conversion + (30,200) drop (src/rsig.c:488-503);
    used only to feed the CPU checkers, never the product path."""
    scale = np.float32(rng_ / digitisation)
    pa = ((raw.astype(np.float64) + offset) * np.float64(scale)).astype(np.float32)
    return pa[(pa > np.float32(30.0)) & (pa < np.float32(200.0))]


# GRCh38 primary assembly chromosome lengths (chr1..22, X, Y): the contig layout of the human-size synthetic genome
GRCH38_LENS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
               133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
               58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
GRCH38_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]


class DeviceGenome:
    """i.i.d. uniform ACGT genome generated in device memory (2-bit codes, one per byte, contigs back to back).

    Human-size inputs (3.1 Gb) are generated where they are consumed; nothing of this size exists on the boxes."""

    def __init__(self, names, lens, device="cuda", seed: int = 1):
        import torch
        self.names = list(names)
        self.lens = np.asarray(lens, dtype=np.int64)
        self.total = int(self.lens.sum())
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        self.codes = torch.empty(self.total, dtype=torch.uint8, device=device)
        step = 1 << 28
        for a in range(0, self.total, step):  # bounded temporaries
            b = min(a + step, self.total)
            self.codes[a:b] = torch.randint(0, 4, (b - a,), dtype=torch.uint8, device=device, generator=gen)
        self.starts = np.concatenate([[0], np.cumsum(self.lens)[:-1]]).astype(np.int64)

    def host_contigs(self):
        """[(name, uint8 codes)] on the host (small genomes only: parity checks against the host builders)."""
        h = self.codes.cpu().numpy()
        return [(n, h[s:s + l]) for n, s, l in zip(self.names, self.starts, self.lens)]


def make_reads_torch(genome, n_reads: int, read_len_bp: int, k: int, means, stdv, device="cuda",
                     sample_rate: float = 4000.0, bp_per_sec: float = 450.0, seed: int = 2, batch: int = 4000):
    """Same recipe as make_reads, generated on the GPU with torch (bench-sized inputs: 10^5 reads).

    `genome` is a list of (name, uint8 codes) or a DeviceGenome.
    Returns (raw int16 tensor on `device`, reads concatenated back to back; raw_off uint64 numpy
    [n+1]; lens uint64 numpy [n]; truth list).  Read i is raw[raw_off[i] : raw_off[i+1]]."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    if isinstance(genome, DeviceGenome):
        G = genome.codes
        clen = genome.lens.copy()
        genome = [None] * len(clen)
    else:
        G = torch.from_numpy(np.concatenate([s for _, s in genome])).to(device)
        clen = np.array([len(s) for _, s in genome], dtype=np.int64)
    cstart = np.concatenate([[0], np.cumsum(clen)[:-1]])
    M = torch.from_numpy(np.asarray(means, dtype=np.float32)).to(device)
    SD = torch.from_numpy(np.asarray(stdv, dtype=np.float32)).to(device)
    scale = RANGE / DIGITISATION
    mean_dwell = sample_rate / bp_per_sec
    rng = np.random.Generator(np.random.PCG64(seed))
    chunks, lens_all, truth = [], [], []
    for b0 in range(0, n_reads, batch):
        B = min(batch, n_reads - b0)
        ci = rng.choice(len(genome), size=B, p=clen / clen.sum())
        L = int(min(read_len_bp, clen.min()))
        st = (rng.random(B) * (clen[ci] - L + 1)).astype(np.int64)
        strand = rng.integers(0, 2, B)
        truth += list(zip(ci.tolist(), st.tolist(), strand.tolist()))
        base = torch.from_numpy(cstart[ci] + st).to(device)[:, None]
        j = torch.arange(L, device=device)[None, :]
        sd_t = torch.from_numpy(strand).to(device)[:, None]
        pos = torch.where(sd_t == 1, base + (L - 1 - j), base + j)
        bases = G[pos].to(torch.int64)
        bases = torch.where(sd_t == 1, 3 - bases, bases)
        nk = L - k + 1
        code = torch.zeros((B, nk), dtype=torch.int64, device=device)
        for t in range(k):
            code = (code << 2) | bases[:, t:t + nk]
        dwell = torch.empty((B, nk), device=device, dtype=torch.float32).exponential_(1.0 / mean_dwell, generator=gen)
        dwell = torch.clamp(torch.round(dwell), min=1).to(torch.int64)
        n_per = dwell.sum(dim=1)
        flat_code = torch.repeat_interleave(code.flatten(), dwell.flatten())
        pa = M[flat_code] + torch.randn(flat_code.shape, device=device, generator=gen) * SD[flat_code]
        raw = torch.clamp(torch.round(pa / scale - OFFSET), -32768, 32767).to(torch.int16)
        chunks.append((raw, n_per.cpu().numpy()))
        lens_all.append(n_per.cpu().numpy())
    lens = np.concatenate(lens_all).astype(np.uint64)
    raw_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    out = torch.cat([c for c, _ in chunks] + [torch.zeros(8, dtype=torch.int16, device=device)])
    return out, raw_off, lens, truth
