#!/usr/bin/env python
"""Text summaries for profiles/: (1) `sass` — tensor-core / TMEM / bulk-copy / mbarrier mnemonics per kernel of the
built library (cuobjdump -sass); (2) `ncu REPORT` — the metrics DESIGN.md quotes from one `ncu --set full` capture."""
import collections
import csv
import io
import re
import subprocess
import sys

MNEMONICS = re.compile(r"\b(UTC[A-Z0-9]*MMA[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|LDTM[.\w]*|STTM[.\w]*|UBLKCP[.\w]*|UTMALDG[.\w]*|"
                       r"SYNCS\.[.\w]*|UCGABAR_\w+|F2FP\.SATFINITE\.E[45]M[23][.\w]*|MEMBAR[.\w]*)")
METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]
STALLS = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def sass(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    fn, per = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = MNEMONICS.search(line)
        if m and fn:
            per.setdefault(fn, collections.OrderedDict()).setdefault(m.group(1), [0, line.strip()[:110]])[0] += 1
    print(f"SASS of {lib} (cuobjdump -sass): tensor-core / TMEM / bulk-copy / mbarrier / 8-bit-float mnemonics per kernel")
    print("(B200_PROFILING.md: tcgen05.mma kind::f16 -> UTCHMMA, kind::f8f6f4 -> UTCQMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR,")
    print(" cp.async.bulk -> UBLKCP, expect_tx -> SYNCS.ARRIVE.TRANS64, cvt.e4m3x2 / e5m2x2 -> F2FP.SATFINITE.E4M3 / E5M2)\n")
    for fn, d in per.items():
        print(fn)
        for k, (n, ex) in d.items():
            print(f"  {n:4d} x {k:40s} e.g. {ex}")
        print()


def ncu(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        print("Kernel Name =", d.get("Kernel Name", ("", "?"))[1])
        for k in METRICS:
            if k in d:
                print(f"{k} [{d[k][0]}] = {d[k][1]}")
        st = sorted(((float(v[1]), STALLS.match(k).group(1)) for k, v in d.items() if STALLS.match(k) and v[1]), reverse=True)
        print("warp stalls per issue (top): " + ", ".join(f"{n} {x:.2f}" for x, n in st[:8]))
        print()


if __name__ == "__main__":
    if sys.argv[1] == "sass":
        sass(sys.argv[2] if len(sys.argv) > 2 else "gpuvmem_b200/libgvmb200.so")
    else:
        ncu(sys.argv[2])
