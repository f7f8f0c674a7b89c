#!/usr/bin/env python
"""Builds oracle/_ref/rawhash2_gpu: the reference's own rawhash2 (main.cpp, slow5lib reader, kt_pipeline, PAF printer)
with step 1 of map_worker_pipeline replaced by librawhash_b200.so — the drop-in of INTEGRATION.md as a real binary.

The reference sources are copied to a scratch directory (never into the repo), edited there in four places, and
compiled with the reference's flags; only the binary lands in oracle/_ref/ (git-ignored).  Each edit is anchored on
the exact reference line it replaces and the script fails if an anchor is missing.

    python integration/build_dropin.py            # needs /root/reference and a built librawhash_b200.so
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RH_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref", "rawhash2_gpu")
LIBDIR = os.path.join(ROOT, "rawhash_b200")

EDITS = {
    # 1. keep what slow5lib returned: raw samples + calibration
    "rsig.h": [(
        "\tfloat* sig; //signal values of a read\n} ri_sig_t;",
        "\tfloat* sig; //signal values of a read\n"
        "\tint16_t *raw; uint64_t l_raw; double cal_offset, cal_range, cal_digitisation; /* rawhash_b200 */\n} ri_sig_t;")],
    "rsig.c": [(
        "\ts->name = strdup(rec->read_id);\n\tfloat *sigF = (float*)malloc(rec->len_raw_signal * sizeof(float));",
        "\ts->name = strdup(rec->read_id);\n"
        "\ts->l_raw = rec->len_raw_signal; s->raw = (int16_t*)malloc((s->l_raw + 1) * sizeof(int16_t)); /* rawhash_b200 */\n"
        "\tmemcpy(s->raw, rec->raw_signal, s->l_raw * sizeof(int16_t));\n"
        "\ts->cal_offset = rec->offset; s->cal_range = rec->range; s->cal_digitisation = rec->digitisation;\n"
        "\tfloat *sigF = (float*)malloc(rec->len_raw_signal * sizeof(float));")],
    # 2. the call site
    "rmap.cpp": [
        ('#include "dtw.h"\n', '#include "dtw.h"\n#include "rh_dropin.h" /* rawhash_b200 */\n'),
        ("\t\tif(!p->su_stop) kt_for(p->n_threads, map_worker_for, in, s->n_sig);",
         "\t\tif(!p->su_stop) rh_dropin_map_step(s); /* rawhash_b200: was kt_for(p->n_threads, map_worker_for, in, s->n_sig) */"),
        # the shim's init/destroy are static in this translation unit; main() reaches them through two exported hooks
        ("static void *map_worker_pipeline(void *shared,",
         "int rh_dropin_start(const ri_idx_t *ri, const ri_mapopt_t *opt, const char *ind_path) { return rh_dropin_init(ri, opt, ind_path); }\n"
         "void rh_dropin_stop(void) { rh_dropin_destroy(); }\n\n"
         "static void *map_worker_pipeline(void *shared,")],
    # 3. one context per run, created once mid_occ is known
    "main.cpp": [
        ("int main(int argc, char *argv[])\n{",
         "int rh_dropin_start(const ri_idx_t *ri, const ri_mapopt_t *opt, const char *ind_path); /* rawhash_b200 */\n"
         "void rh_dropin_stop(void);\n\nint main(int argc, char *argv[])\n{"),
        ("\t\tif (argc != o.ind + 1) ri_mapopt_update(&opt, ri);\n",
         "\t\tif (argc != o.ind + 1) ri_mapopt_update(&opt, ri);\n"
         "\t\tif (argc != o.ind + 1 && rh_dropin_start(ri, &opt, idx_rdr->is_idx ? argv[o.ind] : fnw) != 0) return 1; /* rawhash_b200 */\n"),
        ("\t\tri_idx_destroy(ri);\n\t\tif (ret < 0) {", "\t\tri_idx_destroy(ri);\n\t\trh_dropin_stop(); /* rawhash_b200 */\n\t\tif (ret < 0) {")],
}

C_SOURCES = ["kthread", "kalloc", "bseq", "roptions", "sequence_until", "rutils", "rsig", "revent", "rsketch", "rindex", "lchain", "rseed", "hit"]
CPP_SOURCES = ["dtw", "rmap", "main"]


def main():
    if not os.path.isdir(os.path.join(REF, "src")):
        print(f"[dropin] {REF}/src not present: keeping a prebuilt {OUT} if any")
        return 0
    if not os.path.isfile(os.path.join(LIBDIR, "librawhash_b200.so")):
        sys.exit("[dropin] build rawhash_b200/librawhash_b200.so first (python -m rawhash_b200.build)")
    s5_objs = [os.path.join(ROOT, "oracle", "_ref", "s5", f) for f in sorted(os.listdir(os.path.join(ROOT, "oracle", "_ref", "s5"))) if f.endswith(".o")]
    if not s5_objs:
        sys.exit("[dropin] build the slow5lib objects first (make -C oracle ref)")
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    march = os.environ.get("REF_MARCH", "x86-64-v3")
    with tempfile.TemporaryDirectory(prefix="rh_dropin_") as tmp:
        src = os.path.join(tmp, "src")
        shutil.copytree(os.path.join(REF, "src"), src)
        for fn, edits in EDITS.items():
            path = os.path.join(src, fn)
            text = open(path).read()
            for old, new in edits:
                if text.count(old) != 1:
                    sys.exit(f"[dropin] anchor not found exactly once in src/{fn}: {old[:60]!r}")
                text = text.replace(old, new)
            open(path, "w").write(text)
        flags = ["-std=c++11", "-O3", f"-march={march}", "-pthread", "-DHAVE_KALLOC", "-DNHDF5RH=1", "-DNPOD5RH=1", "-w",
                 "-I" + os.path.join(REF, "extern", "slow5lib", "include"), "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "integration")]
        objs = []
        for name, ext in [(n, ".c") for n in C_SOURCES] + [(n, ".cpp") for n in CPP_SOURCES]:
            obj = os.path.join(tmp, name + ".o")
            subprocess.run([cxx] + flags + ["-c", os.path.join(src, name + ext), "-o", obj], check=True)
            objs.append(obj)
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        zstd = [p for p in ("/lib/x86_64-linux-gnu/libzstd.so.1", "/usr/lib/x86_64-linux-gnu/libzstd.so.1", "/usr/lib64/libzstd.so.1") if os.path.isfile(p)][:1]  # as oracle/Makefile
        subprocess.run([cxx, "-pthread", "-o", OUT] + objs + s5_objs + zstd + ["-L" + LIBDIR, "-lrawhash_b200", "-Wl,-rpath,$ORIGIN/../../rawhash_b200", "-lz", "-lm", "-ldl"], check=True)
    print(OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
