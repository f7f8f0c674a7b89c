"""ctypes binding of the C-ABI (include/rawhash_b200.h) — the host-side mirror of the
reference's mapping interface.

Names follow the reference: ``Params`` mirrors ``ri_idxopt_t``/``ri_mapopt_t`` fields
(src/roptions.h:50-143), ``Index`` stands for ``ri_idx_t`` (src/rindex.h:29-60),
``Mapper.map_batch`` is the ``kt_for(map_worker_for)`` step of ``map_worker_pipeline``
(src/rmap.cpp:700) and returns ``MapRec`` records = ``ri_map_t`` (src/rmap.h:12-22);
``Index.format_paf`` prints them like step 2 of that pipeline (src/rmap.cpp:751-772).

The shared library is mandatory: importing this module without
``rawhash_b200/librawhash_b200.so`` raises, and every GPU entry point raises ``RawHashError``
when no CUDA device is usable.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librawhash_b200.so")

RH_I_SIG_TARGET = 0x20
RH_M_NO_ADAPTIVE = 0x20
RH_M_ALL_CHAINS = 0x2000

TAP_NCNT = 8
TAP_REG_NF = 14


class RawHashError(RuntimeError):
    pass


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


MAPREC_DTYPE = np.dtype([
    ("read_idx", "<u4"), ("c_id", "<u4"), ("read_length", "<u4"), ("ref_id", "<u4"),
    ("read_start_position", "<u4"), ("read_end_position", "<u4"),
    ("fragment_start_position", "<u4"), ("fragment_length", "<u4"),
    ("mapq", "u1"), ("rev", "u1"), ("mapped", "u1"), ("_pad", "u1"),
    ("ci", "<u4"), ("sl", "<u4"), ("cm", "<i4"), ("nc", "<i4"), ("s1", "<i4"), ("mt_ms", "<f4"),
])
assert MAPREC_DTYPE.itemsize == C.sizeof(MapRec)


class GpuStats(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint64), ("n_chunks", C.c_uint64), ("n_rounds", C.c_uint64),
        ("raw_samples_consumed", C.c_uint64),
        ("n_events", C.c_uint64), ("n_seeds", C.c_uint64), ("n_anchors", C.c_uint64), ("n_chains", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("ms_total", C.c_double), ("ms_event_kernel", C.c_double),
        ("ms_seed", C.c_double), ("ms_sort", C.c_double), ("ms_chain", C.c_double), ("ms_post", C.c_double),
        ("event_kernel_launches", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("ms_sort_ties", C.c_double),
        ("event_stage_samples", C.c_uint64), ("event_stage_seeds", C.c_uint64),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class _TapC(C.Structure):
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


class SigBatchC(C.Structure):
    """rh_sigbatch_t: one batch of raw reads out of a SLOW5/BLOW5 file."""
    _fields_ = [
        ("n", C.c_uint32), ("n_samples", C.c_uint64),
        ("raw", C.POINTER(C.POINTER(C.c_int16))), ("raw_len", C.POINTER(C.c_uint64)),
        ("offset", C.POINTER(C.c_double)), ("range", C.POINTER(C.c_double)),
        ("digitisation", C.POINTER(C.c_double)), ("sampling_rate", C.POINTER(C.c_double)),
        ("names", C.POINTER(C.c_char_p)), ("arena_pinned", C.c_int32), ("priv", C.c_void_p),
    ]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise RawHashError(
            f"{LIB_PATH} is missing: build it with `python -m rawhash_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, cp, u32, u64, i32, dbl = C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint64, C.c_int, C.c_double
    PP = C.POINTER(Params)
    sig = {
        "rh_params_init": (None, [PP]),
        "rh_params_preset": (i32, [PP, cp]),
        "rh_params_r10": (None, [PP]),
        "rh_pore_load": (i32, [cp, i32, i32, C.POINTER(C.POINTER(C.c_float)), C.POINTER(u32)]),
        "rh_free": (None, [vp]),
        "rh_index_build": (vp, [PP, vp, u32, u32, vp, vp, vp, i32]),
        "rh_index_build_gpu": (vp, [PP, vp, u32, u32, vp, vp, vp, i32]),
        "rh_index_build_sig": (vp, [PP, u32, vp, vp, vp, vp, vp, vp]),
        "rh_index_build_dev": (vp, [PP, vp, u32, u32, vp, vp, vp, i32]),
        "rh_index_on_device": (i32, [vp]),
        "rh_index_flat": (i32, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "rh_index_load": (vp, [cp, PP]),
        "rh_index_destroy": (None, [vp]),
        "rh_index_n_seq": (u32, [vp]),
        "rh_index_seq_name": (cp, [vp, u32]),
        "rh_index_seq_len": (u32, [vp, u32]),
        "rh_index_n_keys": (u64, [vp]),
        "rh_index_n_pos": (u64, [vp]),
        "rh_index_key": (u32, [vp, u64]),
        "rh_index_update_mapopt": (None, [vp, PP]),
        "rh_index_get": (C.POINTER(u64), [vp, u32, C.POINTER(i32)]),
        "rh_gpu_init": (vp, [vp, PP, i32, C.c_size_t]),
        "rh_gpu_destroy": (None, [vp]),
        "rh_gpu_last_error": (cp, []),
        "rh_gpu_set_stream": (None, [vp, vp]),
        "rh_gpu_set_workers": (i32, [vp, i32]),
        "rh_gpu_map_batch_raw": (i32, [vp, u32, vp, vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(u64)]),
        "rh_gpu_map_batch_dev": (i32, [vp, u32, vp, vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(u64)]),
        "rh_gpu_get_stats": (None, [vp, C.POINTER(GpuStats)]),
        "rh_gpu_tap_read": (i32, [vp, vp, u64, dbl, dbl, dbl, cp, C.POINTER(_TapC)]),
        "rh_format_paf": (vp, [vp, vp, u64, vp]),
        "rh_index_dump": (i32, [vp, cp, vp, u32]),
        "rh_fasta_load": (i32, [cp, C.POINTER(u32), C.POINTER(C.POINTER(cp)), C.POINTER(C.POINTER(cp)), C.POINTER(C.POINTER(u32))]),
        "rh_fasta_free": (None, [u32, vp, vp, vp]),
        "rh_sigfile_open": (vp, [cp, i32]),
        "rh_sigfile_close": (None, [vp]),
        "rh_sigfile_next_batch": (i32, [vp, u64, u32, C.POINTER(C.POINTER(SigBatchC))]),
        "rh_sigbatch_free": (None, [C.POINTER(SigBatchC)]),
        "rh_split_by_samples": (None, [u32, vp, u32, vp]),
        "rh_zlib_inflate": (i32, [vp, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "rh_find_sigfiles": (i32, [cp, C.POINTER(C.POINTER(vp)), C.POINTER(u32)]),
        "rh_slow5_write": (i32, [cp, u32, vp, vp, vp, vp, vp, vp, dbl, i32, i32]),
        "rh_plan_round": (i32, [vp, u32, u32, u32, u64, i32, vp, vp, vp, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return L, list(sig)


_lib, EXPORTED_SYMBOLS = _load()


def _err(prefix: str) -> RawHashError:
    return RawHashError(f"{prefix}: {_lib.rh_gpu_last_error().decode(errors='replace')}")


def make_params(preset: str = "sensitive", r10: bool = False, **overrides) -> Params:
    """Defaults + preset + --r10, in the order src/main.cpp applies them."""
    p = Params()
    _lib.rh_params_init(C.byref(p))
    if _lib.rh_params_preset(C.byref(p), (preset or "").encode()) != 0:
        raise RawHashError(f"unknown preset {preset!r}")
    if r10:
        _lib.rh_params_r10(C.byref(p))
    for k, v in overrides.items():
        setattr(p, k, v)
    if "sample_rate" in overrides or "bp_per_sec" in overrides:
        p.sample_per_base = np.float32(p.sample_rate) / np.float32(p.bp_per_sec)
    return p


def load_pore(path: str, k: int, lev_col: int = 1) -> np.ndarray:
    ptr = C.POINTER(C.c_float)()
    n = C.c_uint32(0)
    if _lib.rh_pore_load(path.encode(), k, lev_col, C.byref(ptr), C.byref(n)) != 0:
        raise _err("rh_pore_load")
    out = np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()
    _lib.rh_free(ptr)
    return out


def _cstr_array(strs: Sequence[str]):
    return (C.c_char_p * len(strs))(*[s.encode() for s in strs])


class Index:
    """Flattened hash index (the data ri_idx_t serves through ri_idx_get)."""

    def __init__(self, handle):
        if not handle:
            raise _err("index")
        self.h = handle

    @classmethod
    def build(cls, params: Params, pore_vals: np.ndarray, names: Sequence[str], seqs: Sequence[str | bytes], n_threads: int = 8):
        pore_vals = np.ascontiguousarray(pore_vals, dtype=np.float32)
        bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
        lens = np.array([len(b) for b in bs], dtype=np.uint32)
        h = _lib.rh_index_build(C.byref(params), pore_vals.ctypes.data, len(pore_vals), len(bs), _cstr_array(names),
                                (C.c_char_p * len(bs))(*bs), lens.ctypes.data, n_threads)
        return cls(h)

    @classmethod
    def build_gpu(cls, params: Params, pore_vals: np.ndarray, names: Sequence[str], seqs: Sequence[str | bytes], device: int = 0):
        """Same index as build(), constructed on the GPU (rh_index_build_gpu)."""
        pore_vals = np.ascontiguousarray(pore_vals, dtype=np.float32)
        bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
        lens = np.array([len(b) for b in bs], dtype=np.uint32)
        h = _lib.rh_index_build_gpu(C.byref(params), pore_vals.ctypes.data, len(pore_vals), len(bs), _cstr_array(names),
                                    (C.c_char_p * len(bs))(*bs), lens.ctypes.data, device)
        return cls(h)

    @classmethod
    def build_dev(cls, params: Params, pore_vals: np.ndarray, names: Sequence[str], d_codes_ptr: int, lens, device: int = 0):
        """Index from 2-bit base codes (uint8 0..3, sequences back to back) already in the memory of `device`
        (rh_index_build_dev); the result stays on the device."""
        pore_vals = np.ascontiguousarray(pore_vals, dtype=np.float32)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        h = _lib.rh_index_build_dev(C.byref(params), pore_vals.ctypes.data, len(pore_vals), len(lens), _cstr_array(names),
                                    C.c_void_p(int(d_codes_ptr)), lens.ctypes.data, device)
        return cls(h)

    @property
    def on_device(self) -> int:
        return _lib.rh_index_on_device(self.h)

    def flat(self):
        """(keys u32[n_keys], off u64[n_keys+1], pos u64[n_pos]) numpy views of the host mirror (no copy)."""
        k, o, p = C.c_void_p(), C.c_void_p(), C.c_void_p()
        if _lib.rh_index_flat(self.h, C.byref(k), C.byref(o), C.byref(p)) != 0:
            raise _err("rh_index_flat")
        nk, npos = self.n_keys, self.n_pos
        mk = lambda ptr, n, ct: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(max(n, 1),))[:n]
        return mk(k, nk, C.c_uint32), mk(o, nk + 1, C.c_uint64), mk(p, npos, C.c_uint64)

    def key(self, i: int) -> int:
        return _lib.rh_index_key(self.h, int(i))

    @classmethod
    def build_from_signals(cls, params: Params, names, raws, offset, rng, digitisation):
        raws = [np.ascontiguousarray(r, dtype=np.int16) for r in raws]
        n = len(raws)
        ptrs = (C.c_void_p * n)(*[r.ctypes.data for r in raws])
        lens = np.array([len(r) for r in raws], dtype=np.uint64)
        off = np.ascontiguousarray(offset, dtype=np.float64); rg = np.ascontiguousarray(rng, dtype=np.float64)
        dg = np.ascontiguousarray(digitisation, dtype=np.float64)
        h = _lib.rh_index_build_sig(C.byref(params), n, _cstr_array(names), ptrs, lens.ctypes.data, off.ctypes.data, rg.ctypes.data, dg.ctypes.data)
        return cls(h)

    @classmethod
    def load(cls, path: str, params: Params):
        return cls(_lib.rh_index_load(path.encode(), C.byref(params)))

    def dump(self, path: str, pore_vals: np.ndarray | None = None):
        """Write the reference's `.ind` layout (ri_idx_dump, src/rindex.c:545-648)."""
        pv = None if pore_vals is None else np.ascontiguousarray(pore_vals, dtype=np.float32)
        rc = _lib.rh_index_dump(self.h, path.encode(), None if pv is None else pv.ctypes.data, 0 if pv is None else len(pv))
        if rc != 0:
            raise _err("rh_index_dump")

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:  # _lib is None while the interpreter shuts down
            _lib.rh_index_destroy(self.h)
            self.h = None

    @property
    def n_seq(self):
        return _lib.rh_index_n_seq(self.h)

    def seq_name(self, i):
        return _lib.rh_index_seq_name(self.h, i).decode()

    def seq_len(self, i):
        return _lib.rh_index_seq_len(self.h, i)

    @property
    def n_keys(self):
        return _lib.rh_index_n_keys(self.h)

    @property
    def n_pos(self):
        return _lib.rh_index_n_pos(self.h)

    def update_mapopt(self, params: Params) -> int:
        _lib.rh_index_update_mapopt(self.h, C.byref(params))
        return params.mid_occ

    def get(self, h: int) -> np.ndarray:
        n = C.c_int(0)
        ptr = _lib.rh_index_get(self.h, int(h) & 0xFFFFFFFF, C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=np.uint64)
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()

    def format_paf(self, recs: np.ndarray, names: Sequence[str]) -> str:
        recs = np.ascontiguousarray(recs)
        p = _lib.rh_format_paf(self.h, recs.ctypes.data, len(recs), _cstr_array(names))
        s = C.string_at(p).decode()
        _lib.rh_free(p)
        return s


class Mapper:
    """GPU mapping context: index resident in HBM + work arenas (rh_gpu_ctx)."""

    def __init__(self, index: Index, params: Params, device: int = 0, arena_bytes: int = 0):
        self.index = index
        self.params = params
        self.h = _lib.rh_gpu_init(index.h, C.byref(params), device, arena_bytes)
        if not self.h:
            raise _err("rh_gpu_init")

    def close(self):
        if getattr(self, "h", None):
            _lib.rh_gpu_destroy(self.h)
            self.h = None

    __del__ = close

    def _take(self, rc, recs_p, n_recs):
        if rc != 0:
            raise _err(f"map_batch (rc={rc})")
        n = n_recs.value
        out = np.frombuffer(C.string_at(recs_p.value, n * MAPREC_DTYPE.itemsize), dtype=MAPREC_DTYPE).copy() if n else np.zeros(0, MAPREC_DTYPE)
        _lib.rh_free(recs_p)
        return out

    def map_batch(self, raws, offset, rng, digitisation, names=None) -> np.ndarray:
        """Host buffers in (list of int16 arrays), records out — the kt_for(map_worker_for) replacement."""
        raws = [np.ascontiguousarray(r, dtype=np.int16) for r in raws]
        n = len(raws)
        ptrs = (C.c_void_p * n)(*[r.ctypes.data for r in raws])
        lens = np.array([len(r) for r in raws], dtype=np.uint64)
        off = np.ascontiguousarray(offset, dtype=np.float64); rg = np.ascontiguousarray(rng, dtype=np.float64)
        dg = np.ascontiguousarray(digitisation, dtype=np.float64)
        recs_p, n_recs = C.c_void_p(), C.c_uint64(0)
        nm = _cstr_array(names) if names is not None else None
        rc = _lib.rh_gpu_map_batch_raw(self.h, n, ptrs, lens.ctypes.data, off.ctypes.data, rg.ctypes.data, dg.ctypes.data, nm,
                                       C.byref(recs_p), C.byref(n_recs))
        return self._take(rc, recs_p, n_recs)

    def map_batch_ptrs(self, ptrs, lens: np.ndarray, offset, rng, digitisation, names_c=None) -> np.ndarray:
        """Same call with a prebuilt pointer array (avoids per-call Python marshalling in benchmarks)."""
        recs_p, n_recs = C.c_void_p(), C.c_uint64(0)
        rc = _lib.rh_gpu_map_batch_raw(self.h, len(lens), ptrs, lens.ctypes.data, offset.ctypes.data, rng.ctypes.data,
                                       digitisation.ctypes.data, names_c, C.byref(recs_p), C.byref(n_recs))
        return self._take(rc, recs_p, n_recs)

    def map_batch_device(self, d_raw_ptr: int, raw_off: np.ndarray, offset, rng, digitisation, names=None) -> np.ndarray:
        """Raw samples already resident in HBM (one concatenated int16 buffer + offsets)."""
        raw_off = np.ascontiguousarray(raw_off, dtype=np.uint64)
        n = len(raw_off) - 1
        off = np.ascontiguousarray(offset, dtype=np.float64); rg = np.ascontiguousarray(rng, dtype=np.float64)
        dg = np.ascontiguousarray(digitisation, dtype=np.float64)
        recs_p, n_recs = C.c_void_p(), C.c_uint64(0)
        nm = _cstr_array(names) if names is not None else None
        rc = _lib.rh_gpu_map_batch_dev(self.h, n, C.c_void_p(d_raw_ptr), raw_off.ctypes.data, off.ctypes.data, rg.ctypes.data, dg.ctypes.data, nm,
                                       C.byref(recs_p), C.byref(n_recs))
        return self._take(rc, recs_p, n_recs)

    def set_stream(self, cuda_stream: int | None):
        _lib.rh_gpu_set_stream(self.h, C.c_void_p(cuda_stream or 0))

    def set_workers(self, n: int) -> int:
        """Concurrent read ranges per batch (own CUDA stream each); returns the count in effect."""
        return _lib.rh_gpu_set_workers(self.h, int(n))

    def stats(self) -> dict:
        st = GpuStats()
        _lib.rh_gpu_get_stats(self.h, C.byref(st))
        return st.as_dict()

    def tap_read(self, raw, offset, rng, digitisation, name="q", caps=None):
        """Per-chunk stage outputs of one read (parity tests)."""
        caps = dict(chunks=64, events=200_000, seeds=200_000, anchors=4_000_000, u=100_000, chain=400_000, regs=100_000) | (caps or {})
        raw = np.ascontiguousarray(raw, dtype=np.int16)
        cnt = np.zeros(caps["chunks"] * TAP_NCNT, dtype=np.int32)
        events = np.zeros(caps["events"], dtype=np.float32)
        seeds = np.zeros(caps["seeds"] * 2, dtype=np.uint64)
        anchors = np.zeros(caps["anchors"] * 2, dtype=np.uint64)
        u = np.zeros(caps["u"], dtype=np.uint64)
        chain_a = np.zeros(caps["chain"] * 2, dtype=np.uint64)
        prev_a = np.zeros(caps["chain"] * 2, dtype=np.uint64)
        regs = np.zeros(caps["regs"] * TAP_REG_NF, dtype=np.int32)
        t = _TapC()
        t.cnt = cnt.ctypes.data_as(C.POINTER(C.c_int32)); t.cap_chunks = caps["chunks"]
        t.events = events.ctypes.data_as(C.POINTER(C.c_float)); t.cap_events = caps["events"]
        t.seeds = seeds.ctypes.data_as(C.POINTER(C.c_uint64)); t.cap_seeds = caps["seeds"]
        t.anchors = anchors.ctypes.data_as(C.POINTER(C.c_uint64)); t.cap_anchors = caps["anchors"]
        t.u = u.ctypes.data_as(C.POINTER(C.c_uint64)); t.cap_u = caps["u"]
        t.chain_a = chain_a.ctypes.data_as(C.POINTER(C.c_uint64)); t.cap_chain_a = caps["chain"]
        t.prev_a = prev_a.ctypes.data_as(C.POINTER(C.c_uint64)); t.cap_prev_a = caps["chain"]
        t.regs = regs.ctypes.data_as(C.POINTER(C.c_int32)); t.cap_regs = caps["regs"]
        rc = _lib.rh_gpu_tap_read(self.h, raw.ctypes.data, len(raw), offset, rng, digitisation, name.encode(), C.byref(t))
        if rc != 0:
            raise _err(f"rh_gpu_tap_read (rc={rc})")
        out = []
        o = dict(ev=0, seed=0, anc=0, u=0, ca=0, reg=0)
        for c in range(t.n_chunks):
            k = cnt[c * TAP_NCNT:(c + 1) * TAP_NCNT]
            ne, nsd, na, nu, nv, nr = int(k[1]), int(k[2]), int(k[3]), int(k[4]), int(k[5]), int(k[6])
            out.append({
                "cnt": k.copy(),
                "events": events[o["ev"]:o["ev"] + ne].copy(),
                "seeds": seeds[2 * o["seed"]:2 * (o["seed"] + nsd)].reshape(-1, 2).copy(),
                "anchors": anchors[2 * o["anc"]:2 * (o["anc"] + na)].reshape(-1, 2).copy(),
                "u": u[o["u"]:o["u"] + nu].copy(),
                "chain_a": chain_a[2 * o["ca"]:2 * (o["ca"] + nv)].reshape(-1, 2).copy(),
                "prev_a": prev_a[2 * o["ca"]:2 * (o["ca"] + nv)].reshape(-1, 2).copy(),
                "regs": regs[o["reg"] * TAP_REG_NF:(o["reg"] + nr) * TAP_REG_NF].reshape(-1, TAP_REG_NF).copy(),
            })
            o["ev"] += ne; o["seed"] += nsd; o["anc"] += na; o["u"] += nu; o["ca"] += nv; o["reg"] += nr
        return out


# ---- files either side of the path (csrc/rh_io.cpp) -------------------------------------------------------------
def read_fasta(path: str):
    """(names, sequences) of a FASTA / FASTA.gz file, read like mm_bseq_read (src/bseq.c)."""
    n = C.c_uint32(0)
    names, seqs, lens = C.POINTER(C.c_char_p)(), C.POINTER(C.c_char_p)(), C.POINTER(C.c_uint32)()
    if _lib.rh_fasta_load(path.encode(), C.byref(n), C.byref(names), C.byref(seqs), C.byref(lens)) != 0:
        raise _err("rh_fasta_load")
    out_n = [names[i].decode() for i in range(n.value)]
    out_s = [C.string_at(seqs[i], lens[i]) for i in range(n.value)]
    _lib.rh_fasta_free(n, names, seqs, lens)
    return out_n, out_s


def split_by_samples(lens, parts: int) -> np.ndarray:
    """rh_split_by_samples: boundaries of `parts` contiguous read ranges balanced by sample count."""
    lens = np.ascontiguousarray(lens, dtype=np.uint64)
    cut = np.zeros(parts + 1, dtype=np.uint32)
    _lib.rh_split_by_samples(len(lens), lens.ctypes.data, parts, cut.ctypes.data)
    return cut


def zlib_inflate(data: bytes) -> bytes:
    """One zlib stream through the library's own inflate (csrc/rh_inflate.h)."""
    out, n = C.c_void_p(), C.c_size_t(0)
    buf = (C.c_char * max(len(data), 1)).from_buffer_copy(data or b"\0")
    if _lib.rh_zlib_inflate(C.cast(buf, C.c_void_p), len(data), C.byref(out), C.byref(n)) != 0:
        raise _err("rh_zlib_inflate")
    res = C.string_at(out, n.value)
    _lib.rh_free(out)
    return res


def find_signal_files(path: str):
    files, n = C.POINTER(C.c_void_p)(), C.c_uint32(0)
    if _lib.rh_find_sigfiles(path.encode(), C.byref(files), C.byref(n)) != 0:
        raise _err("rh_find_sigfiles")
    out = [C.string_at(files[i]).decode() for i in range(n.value)]
    for i in range(n.value):
        _lib.rh_free(files[i])
    _lib.rh_free(C.cast(files, C.c_void_p))
    return out


class SignalFile:
    """SLOW5/BLOW5 reader (stands where ri_sig_open_slow5/ri_read_sig_slow5 do, src/rsig.c:170-207,478-533)."""

    def __init__(self, path: str, n_threads: int = 4):
        self._h = _lib.rh_sigfile_open(path.encode(), n_threads)
        if not self._h:
            raise _err("rh_sigfile_open")

    def close(self):
        if self._h:
            _lib.rh_sigfile_close(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def next_batch(self, max_samples: int = 0, max_reads: int = 0):
        """dict(names, raw=[int16 arrays], offset, range, digitisation, sampling_rate) or None at end of file."""
        b = C.POINTER(SigBatchC)()
        if _lib.rh_sigfile_next_batch(self._h, max_samples, max_reads, C.byref(b)) != 0:
            raise _err("rh_sigfile_next_batch")
        if not b:
            return None
        c = b.contents
        n = c.n
        out = {
            "names": [c.names[i].decode() for i in range(n)],
            "raw": [np.ctypeslib.as_array(c.raw[i], shape=(c.raw_len[i],)).copy() if c.raw_len[i] else np.zeros(0, np.int16) for i in range(n)],
            "offset": np.array([c.offset[i] for i in range(n)]), "range": np.array([c.range[i] for i in range(n)]),
            "digitisation": np.array([c.digitisation[i] for i in range(n)]),
            "sampling_rate": np.array([c.sampling_rate[i] for i in range(n)]),
            "contiguous": all(C.addressof(c.raw[i].contents) + 2 * c.raw_len[i] == C.addressof(c.raw[i + 1].contents) for i in range(n - 1) if c.raw_len[i] and c.raw_len[i + 1]),
        }
        _lib.rh_sigbatch_free(b)
        return out

    def __iter__(self):
        while True:
            b = self.next_batch()
            if b is None:
                return
            yield b


def write_slow5(path: str, names, raws, offset, rng, digitisation, sampling_rate: float = 4000.0, record_press: int = 1, signal_press: int = 1):
    """Write `.slow5` (ASCII) or `.blow5` (record_press 0 none | 1 zlib, signal_press 0 none | 1 svb-zd)."""
    n = len(raws)
    raws = [np.ascontiguousarray(r, dtype=np.int16) for r in raws]
    ptrs = (C.c_void_p * max(n, 1))(*[r.ctypes.data for r in raws])
    lens = np.array([len(r) for r in raws], dtype=np.uint64)
    off = np.ascontiguousarray(np.broadcast_to(np.asarray(offset, dtype=np.float64), (n,)))
    rg = np.ascontiguousarray(np.broadcast_to(np.asarray(rng, dtype=np.float64), (n,)))
    dg = np.ascontiguousarray(np.broadcast_to(np.asarray(digitisation, dtype=np.float64), (n,)))
    nm = _cstr_array(list(names))
    rc = _lib.rh_slow5_write(path.encode(), n, nm, ptrs, lens.ctypes.data, off.ctypes.data, rg.ctypes.data, dg.ctypes.data, float(sampling_rate), record_press, signal_press)
    if rc != 0:
        raise _err("rh_slow5_write")


def plan_round(n_anchors, n_mandatory: int, max_optional: int, arena_bytes: int, heavy_lane: bool = True):
    """Layout of one chunk round in the anchor arena (rh_plan_round; pure host logic, the scheduler calls the same code).
    Returns dict(order, a_off, groups=[(first, count, heavy)], n_run, main_bytes)."""
    na = np.ascontiguousarray(n_anchors, dtype=np.uint32)
    ns = len(na)
    order = np.zeros(max(ns, 1), dtype=np.uint32)
    a_off = np.zeros(max(ns, 1), dtype=np.uint64)
    cap = ns + 2
    groups = np.zeros(3 * cap, dtype=np.uint32)
    ng, nr, mb = C.c_uint32(0), C.c_uint32(0), C.c_uint64(0)
    rc = _lib.rh_plan_round(na.ctypes.data, ns, n_mandatory, max_optional, arena_bytes, 1 if heavy_lane else 0,
                            order.ctypes.data, a_off.ctypes.data, groups.ctypes.data, cap, C.byref(ng), C.byref(nr), C.byref(mb))
    if rc != 0:
        raise _err(f"rh_plan_round (rc={rc})")
    g = groups[: 3 * ng.value].reshape(-1, 3)
    return dict(order=order[:ns], a_off=a_off[:ns], groups=[(int(a), int(b), int(c)) for a, b, c in g], n_run=int(nr.value), main_bytes=int(mb.value))
