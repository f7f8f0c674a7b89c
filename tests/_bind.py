"""ctypes bindings for the two CPU checkers (test infrastructure):

  RefLib    -> oracle/_ref/libref_tap.so  (the unmodified reference + tap harness)
  OracleLib -> oracle/librh_oracle.so     (our restatement)

Both expose the same Python surface so a test can run either against the GPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_tap.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "librh_oracle.so")

TAP_NCNT = 8
TAP_REG_NF = 14
(TAP_NSIG, TAP_NEVENTS, TAP_NSEEDS, TAP_NANCHORS, TAP_NU, TAP_NV, TAP_NREGS, TAP_REPLEN) = range(8)


class Params(C.Structure):
    _fields_ = [
        ("w", C.c_int32), ("e", C.c_int32), ("n", C.c_int32), ("q", C.c_int32), ("k", C.c_int32),
        ("idx_flag", C.c_int32), ("lev_col", C.c_int32),
        ("diff", C.c_float), ("fine_min", C.c_float), ("fine_max", C.c_float), ("fine_range", C.c_float),
        ("window_length1", C.c_uint32), ("window_length2", C.c_uint32),
        ("threshold1", C.c_float), ("threshold2", C.c_float), ("peak_height", C.c_float),
        ("bp_per_sec", C.c_uint32), ("sample_rate", C.c_uint32), ("chunk_size", C.c_uint32),
        ("sample_per_base", C.c_float),
        ("mid_occ_frac", C.c_float), ("min_mid_occ", C.c_int32), ("max_mid_occ", C.c_int32), ("mid_occ", C.c_int32),
        ("min_events", C.c_uint32), ("bw", C.c_int32),
        ("max_target_gap_length", C.c_int32), ("max_query_gap_length", C.c_int32), ("max_chain_iter", C.c_int32),
        ("max_num_skips", C.c_int32), ("min_num_anchors", C.c_int32),
        ("min_chaining_score", C.c_int32), ("min_chaining_score2", C.c_int32),
        ("chain_gap_scale", C.c_float), ("chain_skip_scale", C.c_float),
        ("mask_level", C.c_float), ("mask_len", C.c_int32), ("pri_ratio", C.c_float), ("best_n", C.c_int32),
        ("alt_drop", C.c_float),
        ("w_bestq", C.c_float), ("w_bestmq", C.c_float), ("w_bestmc", C.c_float), ("w_threshold", C.c_float),
        ("max_num_chunk", C.c_uint32), ("min_mapq", C.c_int32),
        ("map_flag", C.c_int64),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class MapRec(C.Structure):
    _fields_ = [
        ("read_idx", C.c_uint32), ("c_id", C.c_uint32), ("read_length", C.c_uint32), ("ref_id", C.c_uint32),
        ("read_start_position", C.c_uint32), ("read_end_position", C.c_uint32),
        ("fragment_start_position", C.c_uint32), ("fragment_length", C.c_uint32),
        ("mapq", C.c_uint8), ("rev", C.c_uint8), ("mapped", C.c_uint8), ("_pad", C.c_uint8),
        ("ci", C.c_uint32), ("sl", C.c_uint32), ("cm", C.c_int32), ("nc", C.c_int32), ("s1", C.c_int32),
        ("mt_ms", C.c_float),
    ]


class TapC(C.Structure):
    _fields_ = [
        ("n_chunks", C.c_uint32),
        ("cnt", C.POINTER(C.c_int32)), ("cap_chunks", C.c_uint64),
        ("events", C.POINTER(C.c_float)), ("cap_events", C.c_uint64),
        ("seeds", C.POINTER(C.c_uint64)), ("cap_seeds", C.c_uint64),
        ("anchors", C.POINTER(C.c_uint64)), ("cap_anchors", C.c_uint64),
        ("u", C.POINTER(C.c_uint64)), ("cap_u", C.c_uint64),
        ("chain_a", C.POINTER(C.c_uint64)), ("cap_chain_a", C.c_uint64),
        ("prev_a", C.POINTER(C.c_uint64)), ("cap_prev_a", C.c_uint64),
        ("regs", C.POINTER(C.c_int32)), ("cap_regs", C.c_uint64),
    ]


class Tap:
    """Owns the numpy buffers behind an rh_tap_t and slices them per chunk."""

    def __init__(self, cap_chunks=64, cap_events=200_000, cap_seeds=200_000, cap_anchors=4_000_000,
                 cap_u=100_000, cap_chain=400_000, cap_regs=100_000):
        self.cnt = np.zeros(cap_chunks * TAP_NCNT, dtype=np.int32)
        self.events = np.zeros(cap_events, dtype=np.float32)
        self.seeds = np.zeros(cap_seeds * 2, dtype=np.uint64)
        self.anchors = np.zeros(cap_anchors * 2, dtype=np.uint64)
        self.u = np.zeros(cap_u, dtype=np.uint64)
        self.chain_a = np.zeros(cap_chain * 2, dtype=np.uint64)
        self.prev_a = np.zeros(cap_chain * 2, dtype=np.uint64)
        self.regs = np.zeros(cap_regs * TAP_REG_NF, dtype=np.int32)
        self.c = TapC()
        self.c.cnt = self.cnt.ctypes.data_as(C.POINTER(C.c_int32)); self.c.cap_chunks = cap_chunks
        self.c.events = self.events.ctypes.data_as(C.POINTER(C.c_float)); self.c.cap_events = cap_events
        self.c.seeds = self.seeds.ctypes.data_as(C.POINTER(C.c_uint64)); self.c.cap_seeds = cap_seeds
        self.c.anchors = self.anchors.ctypes.data_as(C.POINTER(C.c_uint64)); self.c.cap_anchors = cap_anchors
        self.c.u = self.u.ctypes.data_as(C.POINTER(C.c_uint64)); self.c.cap_u = cap_u
        self.c.chain_a = self.chain_a.ctypes.data_as(C.POINTER(C.c_uint64)); self.c.cap_chain_a = cap_chain
        self.c.prev_a = self.prev_a.ctypes.data_as(C.POINTER(C.c_uint64)); self.c.cap_prev_a = cap_chain
        self.c.regs = self.regs.ctypes.data_as(C.POINTER(C.c_int32)); self.c.cap_regs = cap_regs

    def chunks(self):
        """List of dicts, one per chunk, with the stage arrays of that chunk."""
        out = []
        o = dict(ev=0, seed=0, anc=0, u=0, ca=0, reg=0)
        for c in range(self.c.n_chunks):
            cnt = self.cnt[c * TAP_NCNT:(c + 1) * TAP_NCNT]
            ne, ns, na, nu, nv, nr = (int(cnt[TAP_NEVENTS]), int(cnt[TAP_NSEEDS]), int(cnt[TAP_NANCHORS]),
                                      int(cnt[TAP_NU]), int(cnt[TAP_NV]), int(cnt[TAP_NREGS]))
            d = {
                "cnt": cnt.copy(),
                "events": self.events[o["ev"]:o["ev"] + ne].copy(),
                "seeds": self.seeds[2 * o["seed"]:2 * (o["seed"] + ns)].reshape(-1, 2).copy(),
                "anchors": self.anchors[2 * o["anc"]:2 * (o["anc"] + na)].reshape(-1, 2).copy(),
                "u": self.u[o["u"]:o["u"] + nu].copy(),
                "chain_a": self.chain_a[2 * o["ca"]:2 * (o["ca"] + nv)].reshape(-1, 2).copy(),
                "prev_a": self.prev_a[2 * o["ca"]:2 * (o["ca"] + nv)].reshape(-1, 2).copy(),
                "regs": self.regs[o["reg"] * TAP_REG_NF:(o["reg"] + nr) * TAP_REG_NF].reshape(-1, TAP_REG_NF).copy(),
            }
            o["ev"] += ne; o["seed"] += ns; o["anc"] += na; o["u"] += nu; o["ca"] += nv; o["reg"] += nr
            out.append(d)
        return out


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _sig_arrays(sigs):
    n = len(sigs)
    sigs = [np.ascontiguousarray(s, dtype=np.float32) for s in sigs]
    ptrs = (C.POINTER(C.c_float) * n)(*[_fp(s) for s in sigs])
    lens = np.array([len(s) for s in sigs], dtype=np.uint32)
    return sigs, ptrs, lens


def _names(names):
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    return arr


class _CpuChecker:
    """Common surface of the reference harness and the oracle (prefix differs)."""
    prefix = ""

    def __init__(self, so_path):
        self.lib = C.CDLL(so_path)
        L, p = self.lib, self.prefix
        g = lambda name: getattr(L, p + name)
        g("open").restype = C.c_void_p
        g("open").argtypes = [C.c_char_p, C.c_int, C.c_char_p]
        g("get_params").argtypes = [C.c_void_p, C.POINTER(Params)]
        g("build_index").argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        g("build_index_sig").argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        g("mapopt_update").argtypes = [C.c_void_p]
        g("set_mid_occ").argtypes = [C.c_void_p, C.c_int]
        g("set_best_n").argtypes = [C.c_void_p, C.c_int]
        g("set_sampling").argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        g("set_chunks").argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        g("idx_get").restype = C.POINTER(C.c_uint64)
        g("idx_get").argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
        g("raw_to_pa").restype = C.c_uint32
        g("raw_to_pa").argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_void_p]
        g("detect_events").restype = C.c_uint32
        g("detect_events").argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32]
        g("sketch").restype = C.c_uint32
        g("sketch").argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32]
        g("map_paf").restype = C.c_void_p
        g("map_paf").argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        g("free").argtypes = [C.c_void_p]
        g("tap_read").argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_char_p, C.POINTER(TapC)]
        g("radix_sort_128x").argtypes = [C.c_void_p, C.c_uint64]
        g("radix_sort_64").argtypes = [C.c_void_p, C.c_uint64]
        g("pore_vals").restype = C.c_uint32
        g("pore_vals").argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        g("dynamic_quantize").restype = C.c_uint32
        g("dynamic_quantize").argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32]
        self._g = g
        self.h = None

    def open(self, preset="sensitive", r10=False, pore_path=""):
        self.h = self._g("open")(preset.encode(), int(r10), (pore_path or "").encode())
        assert self.h, "open failed"
        return self

    def params(self) -> Params:
        p = Params()
        self._g("get_params")(self.h, C.byref(p))
        return p

    def pore_vals(self):
        n = self._g("pore_vals")(self.h, None, 0)
        out = np.zeros(n, dtype=np.float32)
        self._g("pore_vals")(self.h, out.ctypes.data, n)
        return out

    def build_index(self, fasta, dump_path="", n_threads=4):
        rc = self._g("build_index")(self.h, fasta.encode(), (dump_path or "").encode(), n_threads)
        assert rc == 0, rc

    def build_index_sig(self, sigs, names, n_threads=4):
        sigs, ptrs, lens = _sig_arrays(sigs)
        rc = self._g("build_index_sig")(self.h, len(sigs), ptrs, lens.ctypes.data, _names(names), n_threads)
        assert rc == 0, rc

    def mapopt_update(self) -> int:
        return self._g("mapopt_update")(self.h)

    def set_mid_occ(self, v):
        self._g("set_mid_occ")(self.h, int(v))

    def set_best_n(self, v):
        self._g("set_best_n")(self.h, int(v))

    def set_sampling(self, sample_rate, bp_per_sec):
        self._g("set_sampling")(self.h, sample_rate, bp_per_sec)

    def set_chunks(self, chunk_size=0, max_num_chunk=0):
        self._g("set_chunks")(self.h, chunk_size, max_num_chunk)

    def idx_get(self, h):
        n = C.c_int(0)
        ptr = self._g("idx_get")(self.h, int(h), C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=np.uint64)
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()

    def raw_to_pa(self, raw, offset, rng, digitisation):
        raw = np.ascontiguousarray(raw, dtype=np.int16)
        out = np.zeros(len(raw), dtype=np.float32)
        n = self._g("raw_to_pa")(raw.ctypes.data, len(raw), offset, rng, digitisation, out.ctypes.data)
        return out[:n].copy()

    def detect_events(self, sig, state=None):
        """state = [mean_sum, std_dev_sum, n_events_sum] carried across chunks (updated in place)."""
        sig = np.ascontiguousarray(sig, dtype=np.float32)
        st = state if state is not None else [0.0, 0.0, 0]
        ms, ss, ns = C.c_double(st[0]), C.c_double(st[1]), C.c_uint32(st[2])
        out = np.zeros(len(sig) + 1, dtype=np.float32)
        n = self._g("detect_events")(self.h, sig.ctypes.data, len(sig), C.byref(ms), C.byref(ss), C.byref(ns),
                                     out.ctypes.data, len(out))
        st[0], st[1], st[2] = ms.value, ss.value, ns.value
        return out[:n].copy()

    def sketch(self, events, rid=0, strand=0):
        events = np.ascontiguousarray(events, dtype=np.float32)
        out = np.zeros((len(events) + 1) * 2 * 4, dtype=np.uint64)
        n = self._g("sketch")(self.h, events.ctypes.data, len(events), rid, strand, out.ctypes.data, len(out) // 2)
        assert n <= len(out) // 2
        return out[:2 * n].reshape(-1, 2).copy()

    def map_paf(self, sigs, names, n_threads=1):
        sigs, ptrs, lens = _sig_arrays(sigs)
        secs = C.c_double(0)
        p = self._g("map_paf")(self.h, len(sigs), ptrs, lens.ctypes.data, _names(names), n_threads, C.byref(secs))
        assert p
        s = C.string_at(p).decode()
        self._g("free")(p)
        return s, secs.value

    def tap_read(self, sig, name="q", tap: Tap | None = None):
        sig = np.ascontiguousarray(sig, dtype=np.float32)
        tap = tap or Tap()
        rc = self._g("tap_read")(self.h, sig.ctypes.data, len(sig), name.encode(), C.byref(tap.c))
        assert rc == 0, rc
        return tap.chunks()

    def radix_sort_128x(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.uint64).copy()
        self._g("radix_sort_128x")(xy.ctypes.data, xy.shape[0])
        return xy

    def radix_sort_64(self, x):
        x = np.ascontiguousarray(x, dtype=np.uint64).copy()
        self._g("radix_sort_64")(x.ctypes.data, len(x))
        return x

    def dynamic_quantize(self, v, fine_min=-2.0, fine_max=2.0, fine_range=0.4, n_buckets=16):
        return self._g("dynamic_quantize")(v, fine_min, fine_max, fine_range, n_buckets)


class RefLib(_CpuChecker):
    prefix = "ref_"

    def __init__(self):
        super().__init__(REF_SO)
        self.lib.ref_index_from_flat.restype = C.c_int
        self.lib.ref_index_from_flat.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]

    def index_from_flat(self, names, lens, keys, off, pos, n_threads=4):
        """Fill the reference's ri_idx_t (khash buckets + position arrays) from a flattened key -> positions table."""
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        assert len(off) == len(keys) + 1
        rc = self.lib.ref_index_from_flat(self.h, len(lens), _names(names), lens.ctypes.data, len(keys), keys.ctypes.data,
                                          off.ctypes.data, pos.ctypes.data, n_threads)
        assert rc == 0, rc


class OracleLib(_CpuChecker):
    prefix = "orc_"

    def __init__(self):
        super().__init__(ORACLE_SO)


def have_ref():
    return os.path.isfile(REF_SO)


def have_oracle():
    return os.path.isfile(ORACLE_SO)


def strip_mt(paf: str) -> str:
    """PAF text with the wall-time tag mt:f: removed (not comparable)."""
    out = []
    for line in paf.splitlines():
        out.append("\t".join(t for t in line.split("\t") if not t.startswith("mt:f:")))
    return "\n".join(out)
