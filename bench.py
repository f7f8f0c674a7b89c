#!/usr/bin/env python
"""bench.py — reads/sec of the RawHash2 mapping hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--reads R] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of R synthetic reads
(rh_gpu_map_batch_*: the drop-in for the reference's kt_for(map_worker_for), src/rmap.cpp:700).

Workloads (--config, numbering of BASELINE.json `configs`):
  2 (default)  human-size: GRCh38-shaped synthetic genome (24 contigs, 3.09 Gb), `-x fast`, synthetic R9.4 4 kHz
               450 bp/s reads of 5 kb — the configuration the metric is quoted on
               (test/evaluation/read_mapping/d5_human_na12878_r94/run_rawhash2.sh:14-15); 24 000 reads per step, fewer
               when the steps asked for would not fit the run's time budget (_reads_per_step)
  3            the same genome with the R10.4.1 9-mer model, `-x fast --r10`, 5 kHz / 400 bp/s
  1            yeast-size: 12 Mb / 16 contigs, `-x sensitive` (the round-1 workload, a parity-test size)

  value     reads/s with the raw int16 samples already resident in HBM
  e2e       reads/s through the same C-ABI call with HOST (pinned) buffers: H2D of the raw samples and D2H of the
            records inside the timed region
  roofline  the event->quantise->hash stage: algorithmic bytes (2 B per raw sample consumed + 16 B per seed emitted,
            SURVEY.md §8d) / its CUDA-event time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the compiled reference's own kt_for(map_worker_for) on all host threads over a bounded sample of
            the same reads; the GPU maps that sample too and `parity` says whether the two PAF texts are equal

The genome is generated in device memory and indexed there (rh_index_build_dev, timed, outside the step); the
reference arm gets the same index by filling the reference's ri_idx_t from the flattened table (oracle/ref_tap.cpp
ref_index_from_flat — its own 3.1 Gb index build takes ≈13 min of host time).

Under torchrun (N>1) every rank maps its own R reads against its own replica of the index (weak scaling; reads
are independent), NCCL only all-reduces the per-rank counters.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_BP = 5000


def _configs():
    from rawhash_b200 import synth
    return {
        1: dict(workload="yeast-size 12 Mb synthetic genome (16 contigs), -x sensitive, synthetic R9.4 4 kHz 450 bp/s reads of 5 kb (BASELINE configs[1])",
                preset="sensitive", r10=False, names=[f"chr{i + 1}" for i in range(16)], lens=[750_000] * 16, kind="r9.4", k=6,
                sample_rate=4000, bp_per_sec=450, reads=100_000, ref_reads=2000),
        2: dict(workload="human-size: GRCh38-shaped synthetic genome (24 contigs, 3.09 Gb), -x fast, synthetic R9.4 4 kHz 450 bp/s reads of 5 kb (BASELINE configs[2])",
                preset="fast", r10=False, names=synth.GRCH38_NAMES, lens=synth.GRCH38_LENS, kind="r9.4", k=6,
                sample_rate=4000, bp_per_sec=450, reads=int(os.environ.get("RH_BENCH_READS_HUMAN", "24000")), ref_reads=400, rate_est=700.0),
        3: dict(workload="human-size: GRCh38-shaped synthetic genome (24 contigs, 3.09 Gb), -x fast --r10, R10.4.1 9-mer model, synthetic 5 kHz 400 bp/s reads of 5 kb (BASELINE configs[3])",
                preset="fast", r10=True, names=synth.GRCH38_NAMES, lens=synth.GRCH38_LENS, kind="r10.4.1", k=9,
                sample_rate=5000, bp_per_sec=400, reads=int(os.environ.get("RH_BENCH_READS_HUMAN", "24000")), ref_reads=400, rate_est=380.0),
    }


E2E_STEPS_MAX = 4   # the end-to-end leg repeats the same step from host buffers: a few steps time it as well as K do


def _reads_per_step(cfg, args):
    """Reads per step per GPU.  A step of the human-size workload lasts tens of seconds, and a batch ends with a tail of
    nearly empty iterations whatever its size, so large batches are the efficient ones (24 000 by default); but the whole
    run — W warm-up, K timed, the end-to-end steps and two more — has to fit a time budget (RH_BENCH_BUDGET_S, 540 s), so
    the batch shrinks when many steps are asked for.  `rate_est` = reads/s expected of this workload on one B200."""
    if args.reads:
        return args.reads
    R = cfg["reads"]
    rate = cfg.get("rate_est")
    if rate and "RH_BENCH_READS_HUMAN" not in os.environ:
        total_steps = args.warmup + args.steps + min(args.steps, E2E_STEPS_MAX) + 2
        budget = float(os.environ.get("RH_BENCH_BUDGET_S", "540"))
        R = min(R, max(4000, int(budget * rate / total_steps) // 1000 * 1000))
    return R


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic_ratio():
    """DRAM bytes per algorithmic byte of the event stage, from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "event_stage_traffic.json")
    try:
        j = json.load(open(p))
        return float(j["dram_bytes_per_algorithmic_byte"]), j.get("source", p)
    except Exception:
        return None, None


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


class World:
    """Genome + index on one GPU, plus the read generator of the chosen config."""

    def __init__(self, cfg, device_index: int):
        import torch
        from rawhash_b200 import api, synth
        self.cfg = cfg
        self.dev = torch.device("cuda", device_index)
        self.model = synth.model_path(cfg["kind"])
        self.data = f"synthetic genome (i.i.d. ACGT, generated on the device) + synthetic reads from the ONT {cfg['kind']} {cfg['k']}-mer model"
        if self.model is None:
            self.model = os.path.join("/tmp", f"rh_bench_synth_{cfg['k']}.model")
            synth.write_synthetic_model(self.model, cfg["k"])
            self.data = f"synthetic genome + synthetic reads from a SYNTHETIC {cfg['k']}-mer pore model (ONT table not staged)"
        self.means, self.stdv = synth.load_model_pa(self.model, cfg["k"])
        over = dict(sample_rate=cfg["sample_rate"], bp_per_sec=cfg["bp_per_sec"]) if cfg["r10"] else {}
        self.P = api.make_params(cfg["preset"], cfg["r10"], **over)
        t0 = time.time()
        self.G = synth.DeviceGenome(cfg["names"], cfg["lens"], device=self.dev, seed=1)
        torch.cuda.synchronize()
        self.t_genome = time.time() - t0
        pore = api.load_pore(self.model, cfg["k"])
        t0 = time.time()
        self.idx = api.Index.build_dev(self.P, pore, self.G.names, self.G.codes.data_ptr(), self.G.lens, device_index)
        self.idx.update_mapopt(self.P)   # mid_occ by a radix select over the CSR offsets in HBM
        torch.cuda.synchronize()
        self.t_index = time.time() - t0

    def reads(self, n, seed):
        from rawhash_b200 import synth
        return synth.make_reads_torch(self.G, n, READ_BP, self.cfg["k"], self.means, self.stdv, device=self.dev,
                                      sample_rate=float(self.cfg["sample_rate"]), bp_per_sec=float(self.cfg["bp_per_sec"]), seed=seed)


def reference_handle(world):
    """The compiled reference (oracle/_ref/libref_tap.so) with its ri_idx_t filled from this world's index."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _bind
    cfg = world.cfg
    if not _bind.have_ref():
        raise RuntimeError("oracle/_ref/libref_tap.so is missing (built by __graft_entry__.build() where /root/reference exists)")
    ref = _bind.RefLib().open(cfg["preset"], cfg["r10"], world.model)
    if cfg["r10"]:
        ref.set_sampling(cfg["sample_rate"], cfg["bp_per_sec"])
    t0 = time.time()
    keys, off, pos = world.idx.flat()
    t_dl = time.time() - t0
    t0 = time.time()
    ref.index_from_flat(world.G.names, world.G.lens, keys, off, pos, os.cpu_count() or 1)
    t_fill = time.time() - t0
    mo = ref.mapopt_update()
    assert mo == world.P.mid_occ, f"mid_occ differs: reference {mo}, device {world.P.mid_occ}"
    return ref, _bind, {"index_download_s": round(t_dl, 2), "reference_index_fill_s": round(t_fill, 2)}


def host_signals(raw_dev, raw_off, n):
    from rawhash_b200 import synth
    host = raw_dev[: int(raw_off[n])].cpu().numpy()
    return [synth.raw_to_pa(host[int(raw_off[i]):int(raw_off[i + 1])], synth.OFFSET, synth.RANGE, synth.DIGITISATION) for i in range(n)]


class HostWorld:
    """Small configs without a CUDA device (the CPU test suite): numpy genome, the reference builds its own index."""

    def __init__(self, cfg):
        from rawhash_b200 import synth
        self.cfg = cfg
        self.model = synth.model_path(cfg["kind"])
        self.data = f"synthetic genome (i.i.d. ACGT) + synthetic reads from the ONT {cfg['kind']} {cfg['k']}-mer model"
        if self.model is None:
            self.model = os.path.join("/tmp", f"rh_bench_synth_{cfg['k']}.model")
            synth.write_synthetic_model(self.model, cfg["k"])
            self.data = "synthetic genome + synthetic reads from a SYNTHETIC pore model (ONT table not staged)"
        self.means, self.stdv = synth.load_model_pa(self.model, cfg["k"])
        rng = np.random.Generator(np.random.PCG64(1))
        self.genome = [(n, rng.integers(0, 4, l, dtype=np.uint8)) for n, l in zip(cfg["names"], cfg["lens"])]


def run_reference(args, cfg, rank):
    """--impl reference: the reference's own kt_for(map_worker_for) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    import torch
    from rawhash_b200 import synth
    n_sample = args.ref_reads or cfg["ref_reads"]
    cores = os.cpu_count() or 1
    if not torch.cuda.is_available():
        if sum(cfg["lens"]) > 200_000_000:
            raise SystemExit("the reference arm of a human-size config takes its index table from the GPU builder (the reference's own 3.1 Gb build needs ~13 min of host time): no CUDA device here")
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _bind
        hw = HostWorld(cfg)
        ref = (_bind.RefLib() if _bind.have_ref() else _bind.OracleLib()).open(cfg["preset"], cfg["r10"], hw.model)
        fa = "/tmp/rh_bench_genome.fa"
        synth.write_fasta(fa, hw.genome)
        ref.build_index(fa, "", cores)
        ref.mapopt_update()
        rd = synth.make_reads(hw.genome, n_sample, READ_BP, cfg["k"], hw.means, hw.stdv, sample_rate=float(cfg["sample_rate"]), bp_per_sec=float(cfg["bp_per_sec"]), seed=77)
        sigs = [synth.raw_to_pa(r, synth.OFFSET, synth.RANGE, synth.DIGITISATION) for r in rd["raw"]]
        data, prep, mid_occ = hw.data, {"index": "built by the reference from the FASTA (ri_idx_gen)"}, None
        kind = "reference" if _bind.have_ref() else "port"
    else:
        torch.cuda.set_device(0)
        world = World(cfg, 0)
        ref, _bind, prep = reference_handle(world)
        prep["index"] = "ri_idx_t filled from the flattened index built by rh_index_build_dev (ref_index_from_flat; equal to ri_idx_gen's on small genomes: tests/test_oracle_vs_ref.py)"
        raw_dev, raw_off, lens, truth = world.reads(n_sample, 1000)
        sigs = host_signals(raw_dev, raw_off, n_sample)
        data, mid_occ, kind = world.data, int(world.P.mid_occ), "reference"
    names = [f"read_0_{i:07d}" for i in range(n_sample)]
    nw = max(8, n_sample // 8)
    for _ in range(args.warmup):
        ref.map_paf(sigs[:nw], names[:nw], cores)
    t_tot = 0.0
    for _ in range(args.steps):
        _, secs = ref.map_paf(sigs, names, cores)
        t_tot += secs
    v = n_sample * args.steps / t_tot
    line = {
        "impl": "reference", "metric": "reads/sec mapped", "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64 (events), u64 (hash/chain)", "data": data,
        "config": {"workload": cfg["workload"], "reads_per_step": n_sample, "mid_occ": mid_occ, **prep},
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": kind,
                         "sample": f"{n_sample} reads per step, kt_for(map_worker_for) wall time with {cores} threads, index load and file parsing excluded"},
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
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=int(os.environ.get("RH_BENCH_CONFIG", "2")), choices=[1, 2, 3])
    ap.add_argument("--reads", type=int, default=0, help="reads per step per GPU (0 = the config's default)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-reads", type=int, default=0, help="bounded CPU sample (reads) for the reference arm / cpu_baseline (0 = the config's default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="profiling run under ncu: no minimum warm-up, numbers are not bench values")
    args = ap.parse_args()
    if args.impl == "ours" and not args.profile:
        args.warmup = max(args.warmup, 3)
    cfg = _configs()[args.config]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    from rawhash_b200 import api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the mapping path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    world = World(cfg, local_rank)   # genome + rh_index_build_dev: not part of the timed region
    P, idx = world.P, world.idx
    R = _reads_per_step(cfg, args)
    e2e_steps = max(1, min(args.steps, E2E_STEPS_MAX))
    t0 = time.time()
    raw_dev, raw_off, lens, truth = world.reads(R, 1000 + rank)
    torch.cuda.synchronize()
    t_synth = time.time() - t0
    n_samples = int(raw_off[-1])
    cal = (np.full(R, synth.OFFSET), np.full(R, synth.RANGE), np.full(R, synth.DIGITISATION))

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
    ptrs = (C.c_void_p * R)(*[base + 2 * int(o) for o in raw_off[:-1]])
    lens64 = np.ascontiguousarray(lens, dtype=np.uint64)

    def step_host():
        return mapper.map_batch_ptrs(ptrs, lens64, *cal, names_c=None)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(args.warmup):
        recs = step_dev()
    step_host()

    # ---- timed: HBM-resident ----
    agg = {k: 0.0 for k in ("ms_event_kernel", "ms_seed", "ms_sort", "ms_sort_ties", "ms_chain", "ms_post", "ms_total")}
    cnt = {k: 0 for k in ("raw_samples_consumed", "n_seeds", "event_kernel_launches", "kernel_launches", "n_chunks", "n_anchors", "n_rounds", "event_stage_samples", "event_stage_seeds")}
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
        for _ in range(e2e_steps):
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
    first = {}
    for r in recs:
        first.setdefault(int(r["read_idx"]), r)
    true_locus = sum(1 for i, r in first.items() if r["mapped"] and int(r["ref_id"]) == truth[i][0] and abs(int(r["fragment_start_position"]) - truth[i][1]) < READ_BP + 1000)
    cmp_fields = [f for f in recs.dtype.names if f != "mt_ms"]   # mt:f: is a time, everything else must agree
    same = bool(np.array_equal(recs[cmp_fields], recs_h[cmp_fields]))
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    c = torch.tensor([R, mapped, true_locus], dtype=torch.int64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # slowest rank defines the step time
        dist.all_reduce(c, op=dist.ReduceOp.SUM)   # NCCL over NVLink: per-rank counters only
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    tot_reads, tot_mapped, tot_true = int(c[0]), int(c[1]), int(c[2])

    if rank == 0:
        peak, peak_src = _peaks()
        ev_launches = max(iso["event_kernel_launches"], 1)
        # what the event-stage launches processed: 2 B per raw sample read + 16 B per seed emitted; the stage runs
        # ahead of the stop decisions, so this includes chunks that were detected and sketched but never mapped
        alg_bytes = 2.0 * iso["event_stage_samples"] + 16.0 * iso["event_stage_seeds"]
        ev_ms = iso["ms_event_kernel"]
        achieved = alg_bytes / (ev_ms * 1e-3) / 1e9 if ev_ms > 0 else 0.0
        alg_bytes_timed = 2.0 * cnt["event_stage_samples"] + 16.0 * cnt["event_stage_seeds"]
        achieved_timed = alg_bytes_timed / (agg["ms_event_kernel"] * 1e-3) / 1e9 if agg["ms_event_kernel"] > 0 else 0.0
        ratio, ratio_src = _traffic_ratio()
        line = {
            "metric": "reads/sec mapped", "value": tot_reads * args.steps / (ms_dev * 1e-3), "unit": "reads/s",
            "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+f64 (events), u64 (hash/chain)", "data": world.data,
            "config": {"workload": cfg["workload"], "reads_per_step_per_gpu": R, "raw_bytes_per_step_per_gpu": 2 * n_samples,
                       "l2": "inputs larger than L2 (no flush needed)", "mid_occ": int(P.mid_occ), "index_keys": int(idx.n_keys),
                       "index_positions": int(idx.n_pos), "genome_bases": int(world.G.total), "genome_generation_s": round(world.t_genome, 2),
                       "index_build_s": round(world.t_index, 2), "index_built_on": "gpu (rh_index_build_dev, device-resident)", "read_synthesis_s": round(t_synth, 2),
                       "parallelism": f"index replicated, reads sharded x{world_size}", "workers_per_gpu": n_workers},
            "e2e": {"value": tot_reads * e2e_steps / (ms_e2e * 1e-3), "unit": "reads/s", "steps": e2e_steps,
                    "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps, "records_equal_to_resident_run": same},
            "gpu_launches": int(cnt["kernel_launches"]),
            "clocks": clocks,
            "roofline": {"kernel": "event stage (raw int16 -> events -> quantise -> hash), timed as one span per launch group", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None,
                         "traffic": (ratio * alg_bytes / ev_launches) if ratio else None, "traffic_source": ratio_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / ev_launches, "avg_launch_ms": ev_ms / ev_launches,
                         "timing": "CUDA events on the launching stream around each event-stage launch group, in a one-worker pass of the same step right after the timed region (kernel timed alone)",
                         "achieved_in_timed_region": achieved_timed,
                         "event_stage_chunks_per_mapped_chunk": (iso["event_stage_samples"] / max(iso["raw_samples_consumed"], 1)),
                         "kernels": "k_evt_sums (warp/chunk) + k_evt_stream (lane/chunk: z, prefix sums, t-statistics, peak detectors) + k_evt_finish (CTA/chunk: segment sort, events, quantise, hash)"},
            "stage_ms_per_step": {k: v / args.steps for k, v in agg.items()},
            "stage_ms_note": "per-stage CUDA-event spans summed over the concurrent workers (they overlap in time; ms_total is the slowest worker)",
            "stage_ms_one_worker": {k: iso[k] for k in agg},
            "mapped_fraction": tot_mapped / max(tot_reads, 1),
            "true_locus_fraction": tot_true / max(tot_reads, 1),
            "chunks_per_read": cnt["n_chunks"] / max(R * args.steps, 1),
            "anchors_per_chunk": cnt["n_anchors"] / max(cnt["n_chunks"], 1),
        }
        if not args.no_cpu_baseline and world_size == 1:
            try:
                cb, par = cpu_baseline_and_parity(world, mapper, raw_dev, raw_off, args.ref_reads or cfg["ref_reads"], cal, recs)
                line["cpu_baseline"] = cb
                line["parity"] = par
            except Exception as e:  # the number above stands; say why the baseline is missing
                line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}
        _emit(line)
    mapper.close()
    if world_size > 1:
        dist.destroy_process_group()


def cpu_baseline_and_parity(world, mapper, raw_dev, raw_off, n_sample, cal, timed_recs):
    """The compiled reference on this box's host cores over the first n_sample reads of the step's batch, and the
    GPU's records for the same reads compared with the reference's PAF text (mt:f: aside)."""
    ref, _bind, prep = reference_handle(world)
    n_sample = min(n_sample, len(raw_off) - 1)
    sigs = host_signals(raw_dev, raw_off, n_sample)
    names = [f"read_0_{i:07d}" for i in range(n_sample)]
    cores = os.cpu_count() or 1
    ref.map_paf(sigs[:32], names[:32], cores)
    exp, secs = ref.map_paf(sigs, names, cores)
    recs = mapper.map_batch_device(raw_dev.data_ptr(), raw_off[: n_sample + 1], cal[0][:n_sample], cal[1][:n_sample], cal[2][:n_sample], names=None)
    got = world.idx.format_paf(recs, names)
    g, e = _bind.strip_mt(got).splitlines(), _bind.strip_mt(exp).splitlines()
    n_eq = sum(1 for a, b in zip(g, e) if a == b)
    cb = {"value": n_sample / secs, "unit": "reads/s", "cores": cores, "kind": "reference",
          "sample": f"first {n_sample} reads of the step's batch, kt_for(map_worker_for) wall time with {cores} threads, index load and file parsing excluded", **prep}
    # the sample is mapped in a call of its own (a small batch takes the plain chunk-round loop); the timed batch's records
    # for the same reads (streaming scheduler, other reads in flight beside them) must be the same records
    cmp_fields = [f for f in recs.dtype.names if f != "mt_ms"]
    head = timed_recs[timed_recs["read_idx"] < n_sample]
    par = {"reads": n_sample, "paf_lines_reference": len(e), "paf_lines_gpu": len(g), "lines_equal": n_eq, "equal": g == e,
           "timed_batch_records_equal": bool(len(head) == len(recs) and np.array_equal(head[cmp_fields], recs[cmp_fields])),
           "against": "oracle/_ref/libref_tap.so (unmodified reference), same reads, PAF text without mt:f:"}
    return cb, par


if __name__ == "__main__":
    main()
