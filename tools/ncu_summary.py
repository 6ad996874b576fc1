"""Summarise an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/r01_layer.ncu-rep [--stalls]  > profiles/r01_layer_kernel.md
"""

import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second", "launch__registers_per_thread",
    "launch__grid_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def ncu(path, page):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    path = sys.argv[1]
    rows = ncu(path, "raw")
    hdr, units = rows[0], rows[1]
    print("# ncu summary of `%s`\n" % path.split("/")[-1])
    print("Captured with `ncu --set full --clock-control none --import-source on` under gpurun (cold caches, "
          "serialised replays: compare shares, not absolutes).\n")
    for n, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        print("## launch %d: `%s`\n" % (n, d.get("Kernel Name", "?")[:90]))
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                print("| %s | %s | %s |" % (k, d[k], units[hdr.index(k)]))
        try:
            rd = float(d["dram__bytes_read.sum"].replace(",", ""))
            wr = float(d["dram__bytes_write.sum"].replace(",", ""))
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            print("| **traffic = dram read + write** | %.1f | Mbyte per launch |" % ((rd * scale[ur] + wr * scale[uw]) / 1e6))
        except Exception:
            pass
        print()
    if "--stalls" in sys.argv:
        rows = ncu(path, "source")
        heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
        start = heads[0]
        end = heads[1] - 1 if len(heads) > 1 else len(rows)
        hdr, body = rows[start], rows[start + 1:end]
        i_s, i_n = hdr.index("Source"), hdr.index("# Samples")
        cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
        total = sum(int(r[i_n] or 0) for r in body if len(r) > i_n)
        print("## hottest SASS instructions of launch 0 (warp-state samples, %d total)\n" % total)
        print("| # | samples | share | instruction | dominant stall |\n|---|---|---|---|---|")
        top = sorted(((int(r[i_n] or 0), i) for i, r in enumerate(body) if len(r) > i_n), reverse=True)[:25]
        for n, i in top:
            r = body[i]
            st = {c: int(r[hdr.index(c)] or 0) for c in cols}
            dom = max(st, key=st.get)
            ins = re.sub(r"\s+", " ", r[i_s]).strip()[:70]
            print("| %d | %d | %.1f%% | `%s` | %s |" % (i, n, 100.0 * n / total, ins, dom))


if __name__ == "__main__":
    main()
