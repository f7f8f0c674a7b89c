#!/usr/bin/env python
"""bench.py — reads/sec of the RawHash2 mapping hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of R synthetic reads
(rh_gpu_map_batch_*: the drop-in for the reference's kt_for(map_worker_for)).
Workload (config.workload): BASELINE.json configs[1] — a 12 Mb, 16-contig yeast-sized
synthetic genome indexed with `-x sensitive`, R9.4 4 kHz / 450 bp/s synthetic reads of 5 kb.

  value   reads/s with the raw int16 samples already resident in HBM
  e2e     reads/s through the same C-ABI call with HOST (pinned) buffers: H2D of the raw
          samples and D2H of the records inside the timed region
  roofline  the fused event->quantise->hash kernel (k_signal_to_seeds): algorithmic bytes
          (2 B per raw sample consumed + 16 B per seed emitted) / its CUDA-event
          time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU path (oracle/_ref, all host threads) on a bounded sample

Under torchrun (N>1) every rank maps its own R reads against its own replica of the index
(weak scaling; reads are independent), NCCL only all-reduces the per-rank counters.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GENOME_LEN = 12_000_000
N_CONTIGS = 16
READ_BP = 5000
PRESET = "sensitive"
WORKLOAD = "yeast-sized 12 Mb synthetic genome (16 contigs), -x sensitive, synthetic R9.4 4 kHz 450 bp/s reads of 5 kb (BASELINE configs[1])"


# dram__bytes_read.sum + dram__bytes_write.sum of the event-stage kernels in the `ncu --set full` capture
# profiles/r01_final_event_raw.csv (one launch group of 20 000 chunks), divided by that group's algorithmic bytes.
# The intermediates (z, prefix sums, two t-statistic arrays) round-trip through HBM between the stage's kernels.
TRAFFIC_PER_ALG_BYTE = 17.5
TRAFFIC_SOURCE = "profiles/r01_final_event_raw.csv: DRAM read+write bytes of the stage's kernels for a 20000-chunk launch group / its algorithmic bytes, scaled to this run's average launch"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.stop = False
        self.gpu = gpu_index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max((int(r[1]) for r in self.rows if r[1].isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons}


def build_world(seed=1):
    from rawhash_b200 import synth
    mp = synth.model_path("r9.4")
    data = "synthetic genome + synthetic reads from the ONT R9.4 6-mer model"
    if mp is None:
        mp = os.path.join("/tmp", "rh_bench_synth.model")
        synth.write_synthetic_model(mp, 6)
        data = "synthetic genome + synthetic reads from a SYNTHETIC 6-mer pore model (ONT table not staged)"
    means, stdv = synth.load_model_pa(mp, 6)
    genome = synth.make_genome(N_CONTIGS, GENOME_LEN, seed=seed)
    return mp, means, stdv, genome, data


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref when the
    reference compiled here, else the oracle port), all host threads, bounded sample per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _bind
    from rawhash_b200 import synth
    mp, means, stdv, genome, data = build_world()
    cores = os.cpu_count() or 1
    kind = "reference" if _bind.have_ref() else "port"
    lib = (_bind.RefLib() if kind == "reference" else _bind.OracleLib()).open(PRESET, False, mp)
    fa = "/tmp/rh_bench_genome.fa"
    synth.write_fasta(fa, genome)
    lib.build_index(fa, "", cores)
    lib.mapopt_update()
    n_sample = args.ref_reads
    rd = synth.make_reads(genome, n_sample, READ_BP, 6, means, stdv, seed=77)
    sigs = [synth.raw_to_pa(r, synth.OFFSET, synth.RANGE, synth.DIGITISATION) for r in rd["raw"]]
    for _ in range(args.warmup):
        lib.map_paf(sigs[: max(8, n_sample // 8)], rd["names"][: max(8, n_sample // 8)], cores)
    t_tot = 0.0
    for _ in range(args.steps):
        _, secs = lib.map_paf(sigs, rd["names"], cores)
        t_tot += secs
    v = n_sample * args.steps / t_tot
    line = {
        "impl": "reference", "metric": "reads/sec mapped", "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64+u64", "data": data, "config": {"workload": WORKLOAD, "reads_per_step": n_sample},
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": kind,
                         "sample": f"{n_sample} reads per step, kt_for(map_worker_for) wall time, index load and file parsing excluded"},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """Keep stdout for the one JSON line: libraries (NCCL prints its version banner with printf when
    NCCL_DEBUG=VERSION, the reference prints progress) write to fd 1, so fd 1 is pointed at stderr for the
    duration of the run and the result goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=int(os.environ.get("RH_BENCH_READS", "100000")), help="reads per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-reads", type=int, default=2000, help="bounded CPU sample (reads) for the reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="profiling run under ncu: no minimum warm-up, numbers are not bench values")
    args = ap.parse_args()
    if args.impl == "ours" and not args.profile:
        args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from rawhash_b200 import api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the mapping path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    mp, means, stdv, genome, data = build_world()
    P = api.make_params(PRESET)
    pore = api.load_pore(mp, 6)
    gs = synth.genome_to_strings(genome)
    t0 = time.time()
    idx = api.Index.build_gpu(P, pore, [n for n, _ in gs], [s for _, s in gs], local_rank)   # rh_index_build_gpu: not part of the timed region
    idx.update_mapopt(P)
    t_index = time.time() - t0
    R = args.reads
    t0 = time.time()
    raw_dev, raw_off, lens, truth = synth.make_reads_torch(genome, R, READ_BP, 6, means, stdv, device=dev, seed=1000 + rank)
    torch.cuda.synchronize()
    t_synth = time.time() - t0
    n_samples = int(raw_off[-1])
    cal = (np.full(R, synth.OFFSET), np.full(R, synth.RANGE), np.full(R, synth.DIGITISATION))
    names = [f"read_{rank}_{i:07d}" for i in range(R)]

    free_b, _ = torch.cuda.mem_get_info()
    arena = int(float(os.environ["RH_ARENA_GB"]) * (1 << 30)) if "RH_ARENA_GB" in os.environ else int(free_b * 0.55)
    mapper = api.Mapper(idx, P, local_rank, arena)
    n_workers = mapper.set_workers(int(os.environ.get("RH_WORKERS", "1")))   # concurrent read ranges (own CUDA stream each)

    def step_dev():
        return mapper.map_batch_device(raw_dev.data_ptr(), raw_off, *cal, names=None)

    # host copy for the end-to-end leg (pinned)
    raw_host = torch.empty(n_samples + 8, dtype=torch.int16, pin_memory=True)
    raw_host.copy_(raw_dev[: n_samples + 8])
    torch.cuda.synchronize()
    base = raw_host.data_ptr()
    import ctypes as C
    ptrs = (C.c_void_p * R)(*[base + 2 * int(o) for o in raw_off[:-1]])
    lens64 = np.ascontiguousarray(lens, dtype=np.uint64)

    def step_host():
        return mapper.map_batch_ptrs(ptrs, lens64, *cal, names_c=None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(args.warmup):
        recs = step_dev()
    step_host()

    # ---- timed: HBM-resident ----
    agg = {k: 0.0 for k in ("ms_event_kernel", "ms_seed", "ms_sort", "ms_sort_ties", "ms_chain", "ms_post", "ms_total")}
    cnt = {k: 0 for k in ("raw_samples_consumed", "n_seeds", "event_kernel_launches", "kernel_launches", "n_chunks", "n_anchors", "n_rounds")}
    with ClockSampler(local_rank) as clk:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            recs = step_dev()
            st = mapper.stats()
            for k in agg:
                agg[k] += st[k]
            for k in cnt:
                cnt[k] += st[k]
        e1.record()
        barrier()
        ms_dev = e0.elapsed_time(e1)
        # ---- timed: end to end from host buffers ----
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        h2d = d2h = 0
        for _ in range(args.steps):
            recs_h = step_host()
            st = mapper.stats()
            h2d += st["h2d_bytes"]; d2h += st["d2h_bytes"]
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
    clocks = clk.summary()
    # ---- roofline pass: one worker, so the event-stage launches are timed alone on their stream ----
    mapper.set_workers(1)
    step_dev()
    iso = mapper.stats()
    mapper.set_workers(n_workers)

    mapped = int((recs["mapped"] == 1).sum())
    same = bool(np.array_equal(recs, recs_h))
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    c = torch.tensor([R, mapped], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # slowest rank defines the step time
        dist.all_reduce(c, op=dist.ReduceOp.SUM)   # NCCL over NVLink: per-rank counters only
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    tot_reads, tot_mapped = int(c[0]), int(c[1])

    if rank == 0:
        peak, peak_src = _peaks()
        ev_launches = max(iso["event_kernel_launches"], 1)
        alg_bytes = 2.0 * iso["raw_samples_consumed"] + 16.0 * iso["n_seeds"]
        ev_ms = iso["ms_event_kernel"]
        achieved = alg_bytes / (ev_ms * 1e-3) / 1e9 if ev_ms > 0 else 0.0
        alg_bytes_timed = 2.0 * cnt["raw_samples_consumed"] + 16.0 * cnt["n_seeds"]
        achieved_timed = alg_bytes_timed / (agg["ms_event_kernel"] * 1e-3) / 1e9 if agg["ms_event_kernel"] > 0 else 0.0
        line = {
            "metric": "reads/sec mapped", "value": tot_reads * args.steps / (ms_dev * 1e-3), "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+f64 (events), u64 (hash/chain)", "data": data,
            "config": {"workload": WORKLOAD, "reads_per_step_per_gpu": R, "raw_bytes_per_step_per_gpu": 2 * n_samples,
                       "l2": "inputs larger than L2 (no flush needed)", "mid_occ": int(P.mid_occ), "index_keys": int(idx.n_keys),
                       "index_positions": int(idx.n_pos), "index_build_s": round(t_index, 2), "index_built_on": "gpu", "read_synthesis_s": round(t_synth, 2),
                       "parallelism": f"index replicated, reads sharded x{world}", "workers_per_gpu": n_workers},
            "e2e": {"value": tot_reads * args.steps / (ms_e2e * 1e-3), "unit": "reads/s",
                    "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps, "records_equal_to_resident_run": same},
            "gpu_launches": int(cnt["kernel_launches"]),
            "clocks": clocks,
            "roofline": {"kernel": "event stage: k_sig_norm+k_sig_prefix+k_sig_tstat+k_sig_peaks+k_sig_events_fast(+k_sig_events)+k_sig_sketch (back-to-back launches per chunk group, timed as one span)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None,
                         "traffic": TRAFFIC_PER_ALG_BYTE * alg_bytes / ev_launches, "traffic_source": TRAFFIC_SOURCE, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / ev_launches, "avg_launch_ms": ev_ms / ev_launches,
                         "timing": "CUDA events on the launching stream around each event-stage launch group, in a one-worker pass of the same step right after the timed region (kernel timed alone)",
                         "achieved_in_timed_region": achieved_timed,
                         "note": "bit-exact event detection is instruction-issue bound (~250 instr per 2-byte sample), see DESIGN.md"},
            "stage_ms_per_step": {k: v / args.steps for k, v in agg.items()},
            "stage_ms_note": "per-stage CUDA-event spans summed over the concurrent workers (they overlap in time; ms_total is the slowest worker)",
            "stage_ms_one_worker": {k: iso[k] for k in agg},
            "mapped_fraction": tot_mapped / max(tot_reads, 1),
            "chunks_per_read": cnt["n_chunks"] / max(R * args.steps, 1),
            "anchors_per_chunk": cnt["n_anchors"] / max(cnt["n_chunks"], 1),
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(mp, means, stdv, genome, args.ref_reads)
        _emit(line)
    mapper.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(mp, means, stdv, genome, n_sample):
    """Reference CPU path timed on this box's host cores on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _bind
    from rawhash_b200 import synth
    cores = os.cpu_count() or 1
    kind = "reference" if _bind.have_ref() else "port"
    try:
        lib = (_bind.RefLib() if kind == "reference" else _bind.OracleLib()).open(PRESET, False, mp)
    except OSError:
        kind = "port"
        lib = _bind.OracleLib().open(PRESET, False, mp)
    fa = "/tmp/rh_bench_genome.fa"
    synth.write_fasta(fa, genome)
    lib.build_index(fa, "", cores)
    lib.mapopt_update()
    rd = synth.make_reads(genome, n_sample, READ_BP, 6, means, stdv, seed=77)
    sigs = [synth.raw_to_pa(r, synth.OFFSET, synth.RANGE, synth.DIGITISATION) for r in rd["raw"]]
    lib.map_paf(sigs[:64], rd["names"][:64], cores)
    _, secs = lib.map_paf(sigs, rd["names"], cores)
    return {"value": n_sample / secs, "unit": "reads/s", "cores": cores, "kind": kind,
            "sample": f"{n_sample} reads of the same workload, kt_for(map_worker_for) wall time with {cores} threads, index load and file parsing excluded"}


if __name__ == "__main__":
    main()
