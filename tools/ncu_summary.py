#!/usr/bin/env python
"""Summarise ncu captures (gpurun_out/*.ncu-rep, launches_*.csv) into small text files for profiles/.

    python tools/ncu_summary.py TAG      # reads gpurun_out/*_TAG.*, writes profiles/TAG_*.txt
"""
import collections
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'lts__t_bytes.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'l1tex__t_bytes.sum', 'sm__inst_executed_pipe_lsu.sum']


def raw(rep):
    if rep.endswith(".csv"):          # raw page already exported on the GPU box (tools/profile_gpu.sh)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(out.splitlines()) if r]
    st = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[st], rows[st + 1], rows[st + 2:]


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(tag):
    import json
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    traffic = {}
    reps = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_*_{tag}.ncu-rep"))) + sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_*_{tag}.csv")))
    for rep in reps:
        try:
            hdr, units, vals = raw(rep)
        except StopIteration:
            print("skip (no data)", rep); continue
        name = os.path.splitext(os.path.basename(rep))[0][5:][:-(len(tag) + 1)]
        with open(os.path.join(ROOT, "profiles", f"{tag}_{name}.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on (one launch)  source: gpurun_out/{os.path.basename(rep)}\n")
            for v in vals:
                f.write(f"kernel: {v[hdr.index('Kernel Name')]}\n")
                try:     # DRAM bytes per launch (read + write) -> profiles/traffic.json, the `traffic` key of bench.py's roofline
                    ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
                    traffic[name] = float(v[ir].replace(',', '')) * UNIT[units[ir]] + float(v[iw].replace(',', '')) * UNIT[units[iw]]
                except (ValueError, KeyError):
                    pass
                try:     # L2 utilisation of the launch as ncu reports it -> `l2_frac` of bench.py's roofline entries
                    il = hdr.index('lts__throughput.avg.pct_of_peak_sustained_elapsed')
                    traffic[name + "__l2_pct"] = float(v[il].replace(',', ''))
                except (ValueError, KeyError):
                    pass
                for w in WANT:
                    if w in hdr:
                        f.write(f"  {w:72s} {v[hdr.index(w)]:>18s} {units[hdr.index(w)]}\n")
                f.write("  warp stall reasons (smsp__average_warp*_issue_stalled_*_per_issue_active, > 0.2):\n")
                for i, h in enumerate(hdr):
                    if 'issue_stalled' in h and 'not_issued' not in h and h.endswith('.ratio'):
                        try:
                            if float(v[i]) > 0.2:
                                f.write(f"    {h.split('issue_stalled_')[1]:50s} {float(v[i]):8.2f}\n")
                        except ValueError:
                            pass
        print("wrote", f"profiles/{tag}_{name}.txt")
    if traffic:
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump({"tag": tag, **traffic}, f, indent=1)
        print("wrote profiles/traffic.json", traffic)
    for lc in glob.glob(os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")):
        rows = list(csv.reader(open(lc)))
        st = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
        hdr = rows[st]
        ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
        L = [(r[ik], float(r[iv].replace(',', ''))) for r in rows[st + 1:] if len(r) > iv]
        idx = [i for i, (k, _) in enumerate(L) if 'wb_march_count' in k]
        with open(os.path.join(ROOT, "profiles", f"{tag}_launches.txt"), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: per-step kernel time by kernel (ns are cold-cache, serialised)\n")
            f.write("# command: python bench.py --steps 1 --warmup 1 --no-cpu-baseline   (steps: warm-up, timed, e2e)\n")
            for a, b in zip(idx, idx[1:] + [len(L)]):
                seg = L[a:b]
                tot = sum(v for _, v in seg)
                c = collections.Counter(); n = collections.Counter()
                for k, v in seg:
                    c[k.split('(')[0][:70]] += v; n[k.split('(')[0][:70]] += 1
                f.write(f"step: {len(seg)} launches, {tot / 1e6:.3f} ms of kernels\n")
                for k, v in c.most_common(16):
                    f.write(f"   {v / 1e6:9.3f} ms  {100 * v / tot:5.1f}%  x{n[k]:<3d} {k}\n")
        print("wrote", f"profiles/{tag}_launches.txt")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
