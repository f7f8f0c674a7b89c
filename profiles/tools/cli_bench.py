#!/usr/bin/env python
"""End-to-end run of the command line (not a bench.py line): synthetic 12 Mb genome -> FASTA, synthetic R9.4 reads ->
BLOW5 (svb-zd signals), then `rawhash2_b200 -d` (index build + dump) and `rawhash2_b200 idx reads.blow5 > paf`.
Reports what the CLI itself prints: pipeline reads/s (file decode + H2D + map + PAF) and the mapping step alone.

    python profiles/tools/cli_bench.py [n_reads] [record_press] [-K value[@threads] ...]
"""
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rawhash_b200 import api, synth  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rec_press = int(sys.argv[2]) if len(sys.argv) > 2 else 0
CLI = os.path.join(ROOT, "rawhash_b200", "rawhash2_b200")
tmp = "/tmp/rh_cli_bench"
os.makedirs(tmp, exist_ok=True)
mp = synth.model_path("r9.4")
means, stdv = synth.load_model_pa(mp, 6)
genome = synth.make_genome(16, 12_000_000, seed=1)
fa = os.path.join(tmp, "g.fa")
synth.write_fasta(fa, genome)
t0 = time.time()
raw, raw_off, lens, truth = synth.make_reads_torch(genome, n_reads, 5000, 6, means, stdv, device="cuda", seed=2)
host = raw.cpu().numpy()
raws = [host[int(raw_off[i]):int(raw_off[i + 1])] for i in range(n_reads)]
names = [f"read_{i:07d}" for i in range(n_reads)]
blow5 = os.path.join(tmp, "reads.blow5")
api.write_slow5(blow5, names, raws, synth.OFFSET, synth.RANGE, synth.DIGITISATION, 4000.0, rec_press, 1)
out = {"n_reads": n_reads, "raw_samples": int(lens.sum()), "blow5_mb": round(os.path.getsize(blow5) / 1e6, 1), "record_press": rec_press, "signal_press": 1,
       "make_inputs_s": round(time.time() - t0, 2), "host_threads": os.cpu_count()}
ind = os.path.join(tmp, "g.ind")
t0 = time.time()
r = subprocess.run([CLI, "-x", "sensitive", "-t", str(os.cpu_count()), "-p", mp, "-d", ind, fa], capture_output=True, text=True)
out["index_cmd_s"] = round(time.time() - t0, 2)
out["index_rc"] = r.returncode
out["ind_mb"] = round(os.path.getsize(ind) / 1e6, 1) if os.path.isfile(ind) else None
paf = os.path.join(tmp, "out.paf")
ks = sys.argv[3:] or ["500M", "500M"]
for attempt in ks:
    t0 = time.time()
    kval, _, nthr = attempt.partition("@")  # "500M@8" = -K 500M -t 8
    r = subprocess.run([CLI, "-x", "sensitive", "-t", nthr or str(os.cpu_count()), "-K", kval, "-o", paf, ind, blow5], capture_output=True, text=True, env=dict(os.environ, RH_CLI_VERBOSE="1"))
    attempt = "K" + attempt + ("_again" if "K" + attempt in out else "")
    wall = time.time() - t0
    log_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log_dir):
        open(os.path.join(log_dir, f"cli_bench_K{attempt}.stderr"), "w").write(r.stderr)
    m = re.search(r"mapped (\d+) of (\d+) reads .*pipeline: ([\d.]+) sec \((\d+) reads/s\); mapping step alone: ([\d.]+) sec \((\d+) reads/s\); file decode alone: ([\d.]+) sec; index \+ GPU ready after ([\d.]+) sec; real time: ([\d.]+)", r.stderr)
    out[attempt] = {"rc": r.returncode, "wall_s": round(wall, 2)}
    if m:
        out[attempt] |= {"mapped": int(m.group(1)), "pipeline_s": float(m.group(3)), "pipeline_reads_per_s": int(m.group(4)),
                         "map_step_s": float(m.group(5)), "map_step_reads_per_s": int(m.group(6)), "decode_s": float(m.group(7)), "ready_s": float(m.group(8)), "process_real_s": float(m.group(9)),
                         "batch_map_s": [float(x) for x in re.findall(r"map ([\d.]+) sec", r.stderr)]}
    else:
        out[attempt]["stderr"] = r.stderr[-500:]
# the same files through the reference's own binary (CPU, all host threads) and through the drop-in build of it
# (oracle/_ref/rawhash2_gpu: reference main/reader/printer, our library at the kt_for line); PAFs must be identical
strip = lambda txt: re.sub(r"mt:f:[^\t]*\t", "", txt)
if os.environ.get("RH_WITH_REFERENCE"):
    mine = strip(open(paf).read())
    for tag, exe in (("reference_cpu", "rawhash2"), ("reference_dropin_gpu", "rawhash2_gpu")):
        exe = os.path.join(ROOT, "oracle", "_ref", exe)
        if not os.path.isfile(exe):
            continue
        t0 = time.time()
        r = subprocess.run([exe, "-x", "sensitive", "-t", str(os.cpu_count()), ind, blow5], capture_output=True, text=True)
        wall = time.time() - t0
        out[tag] = {"rc": r.returncode, "wall_s": round(wall, 2), "reads_per_s_whole_program": round(n_reads / wall), "threads": os.cpu_count(),
                    "paf_identical_to_rawhash2_b200": strip(r.stdout) == mine, "paf_lines": len(r.stdout.splitlines()),
                    "stderr_tail": r.stderr[-300:]}
lines = open(paf).read().splitlines() if os.path.isfile(paf) else []
ok = 0
clen = [len(s) for _, s in genome]
for ln in lines:  # true-locus check: the mapped interval overlaps the interval the read was drawn from
    f = ln.split("\t")
    if f[4] == "*":
        continue
    i = int(f[0].split("_")[1])
    ci, st, sd = truth[i]
    if f[5] == genome[ci][0] and int(f[7]) < st + 5000 and int(f[8]) > st and (f[4] == "-") == bool(sd):
        ok += 1
out["paf_lines"] = len(lines)
out["mapped_to_true_locus"] = ok
print(json.dumps(out))
