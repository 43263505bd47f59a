"""Development helper: turn ncu reports brought back in gpurun_out/ into the committed text summaries under profiles/.
   python tools/profile_summary.py <tag> <rep1> <title1> [<rep2> <title2> ...]"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed_op_tma_ld.sum']


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    tag = sys.argv[1]
    pairs = list(zip(sys.argv[2::2], sys.argv[3::2]))
    txt = [f"# {tag} -- ncu summary of the dominant kernel (ncu --set full --clock-control none; durations under ncu are cold-cache and serialised)\n"]
    traffic = {}
    for rep, title in pairs:
        hdr, units, rows = raw(rep)
        r = rows[-1]
        txt.append(f"## {title}\nsource: {rep}\n")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                txt.append(f"{w:84s} {r[i]:>18s} {units[i]}")
        det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        for line in det.splitlines():
            if "highest-utilized pipeline" in line:
                txt.append("pipe: " + line.strip())
        st = []
        for i, h in enumerate(hdr):
            if 'stalled' in h and 'ratio' in h and 'not_issued' not in h and 'per_issue_active' in h:
                try:
                    st.append((float(r[i]), h))
                except ValueError:
                    pass
        txt.append("top warp-stall reasons (warps per issue-active cycle):")
        for v, h in sorted(st, reverse=True)[:6]:
            txt.append(f"   {v:6.3f} {h.split('issue_stalled_')[1].split('_per_')[0]}")
        scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
        rd = float(r[hdr.index('dram__bytes_read.sum')]) * scale[units[hdr.index('dram__bytes_read.sum')]]
        wr = float(r[hdr.index('dram__bytes_write.sum')]) * scale[units[hdr.index('dram__bytes_write.sum')]]
        txt.append(f"DRAM traffic per launch = {(rd + wr) / 1e6:.1f} MB (read {rd / 1e6:.1f} + write {wr / 1e6:.1f})\n")
        traffic[title.split()[0]] = rd + wr
    open(os.path.join("profiles", f"{tag}_fused_ncu_summary.md"), "w").write("\n".join(txt) + "\n")
    print("\n".join(txt))
    return traffic


if __name__ == "__main__":
    main()
