"""Digest of an ncu --set full report: per launch the headline metrics, and (with --source) the hottest
source lines by stall samples. usage: python tools/ncu_digest.py rep.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_sleeping",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_tex_throttle", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_sample_count"]
for w in want:
    if w in h:
        i = h.index(w)
        print(f"{w:70s} {rows[1][i]:>10s} " + " | ".join(r[i][:40] for r in rows[2:]))
if "--source" in sys.argv:
    n = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    print(src[:200])
    blocks = src.split("\n\n")
    for blk in blocks:
        rr = list(csv.reader(io.StringIO(blk)))
        if len(rr) < 3:
            continue
        hh = rr[0]
        cand = [c for c in hh if "Sampling" in c and "All" in c] or [c for c in hh if "Samples" in c]
        if not cand or "Source" not in hh:
            continue
        si, ci = hh.index("Source"), hh.index(cand[0])
        li = hh.index("#") if "#" in hh else 0
        data = []
        for r in rr[1:]:
            try:
                data.append((float(r[ci] or 0), r[li], r[si]))
            except (ValueError, IndexError):
                pass
        tot = sum(d[0] for d in data) or 1
        print(f"--- {cand[0]}: total {tot:.0f}")
        for v, ln, s in sorted(data, reverse=True)[:n]:
            print(f"{v/tot*100:5.1f}%  L{ln:>5s}  {s.strip()[:150]}")
