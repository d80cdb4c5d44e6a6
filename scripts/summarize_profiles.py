"""Turn the ncu outputs that scripts/profile_round.sh left in gpurun_out/ into the committed summaries under profiles/.
usage: python scripts/summarize_profiles.py <tag>        (e.g. r01)
  gpurun_out/launches_<tag>.csv      -> profiles/launches_<tag>.csv (copy) + profiles/launches_<tag>.md (per-kernel shares)
  gpurun_out/prof_fwd_<tag>.ncu-rep  -> profiles/raster_fwd_<tag>.md  + traffic.json entry
  gpurun_out/prof_fill_<tag>.ncu-rep -> profiles/raster_fill_<tag>.md + traffic.json entry
  gpurun_out/prof_bwd_<tag>.ncu-rep  -> profiles/raster_bwd_<tag>.md  + traffic.json entry
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(PR, exist_ok=True)

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def launches():
    src = os.path.join(GO, f"launches_{tag}.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PR, f"launches_{tag}.csv"))
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[iV].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iU], 1.0)
        agg[r[iK]][0] += 1
        agg[r[iK]][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PR, f"launches_{tag}.md"), "w") as f:
        f.write(f"# Launch list, `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` under ncu ({tag})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised:\n"
                "compare SHARES, not absolutes. Covers target set-up, 3 warm-up + 2 timed steps and the e2e loop.\n"
                f"Raw list: `launches_{tag}.csv` ({len(rows) - 1} launches, {tot / 1e3:.1f} ms of device time).\n\n"
                "| share | total us | launches | kernel |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"| {100 * v[1] / tot:.1f}% | {v[1]:.0f} | {v[0]} | `{k[:110]}` |\n")
    print("wrote launches summary")


def full(kind, kname):
    rep = os.path.join(GO, f"prof_{kind}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    r = rows[2]
    name = r[hdr.index("Kernel Name")]
    vals = {k: (r[hdr.index(k)], units[hdr.index(k)]) for k in KEYS if k in hdr}
    rd = to_bytes(*vals['dram__bytes_read.sum'])
    wr = to_bytes(*vals['dram__bytes_write.sum'])
    st = sorted(((float(r[i].replace(",", "")), h.split("issue_stalled_")[1].split("_per")[0]) for i, h in enumerate(hdr)
                 if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith(".ratio") and r[i]), reverse=True)[:8]
    with open(os.path.join(PR, f"raster_{kind}_{tag}.md"), "w") as f:
        f.write(f"# ncu --set full, `{name[:100]}` ({tag})\n\n"
                f"One launch of the bench step (C2: 512 renders, 256x256, K=20), `--clock-control none`, report `gpurun_out/prof_{kind}_{tag}.ncu-rep` "
                "(scratch, not committed).\n\n| metric | value |\n|---|---|\n")
        for k in KEYS:
            if k in vals:
                f.write(f"| `{k}` | {vals[k][0]} {vals[k][1]} |\n")
        f.write(f"| DRAM traffic (read + write) | {(rd + wr) / 1e9:.3f} GB |\n")
        f.write("\nWarp stall reasons (cycles per issued instruction): " + ", ".join(f"{n}={v:.2f}" for v, n in st) + "\n")
    tj = os.path.join(PR, "traffic.json")
    d = json.load(open(tj)) if os.path.exists(tj) else {}
    d.setdefault("C2", {})[kname] = rd + wr
    d["C2"][kname + "_sm_throughput_pct"] = float(vals['sm__throughput.avg.pct_of_peak_sustained_elapsed'][0].replace(",", ""))
    d["C2"][kname + "_source"] = f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch, {tag}"
    json.dump(d, open(tj, "w"), indent=1)
    print("wrote", kind)


launches()
full("fwd", "raster_fwd_kernel")
full("fill", "raster_fill_kernel")
full("bwd", "raster_soft_bwd_kernel")
# the forward op of the split path = raster_fwd_kernel (live regions) + raster_fill_kernel (padding of the empty regions), concurrent
tj = os.path.join(PR, "traffic.json")
if os.path.exists(tj):
    d = json.load(open(tj))
    c2 = d.get("C2", {})
    if "raster_fwd_kernel" in c2 and "raster_fill_kernel" in c2:
        c2["acfm_raster_fwd"] = c2["raster_fwd_kernel"] + c2["raster_fill_kernel"]
        c2["acfm_raster_fwd_source"] = "sum of the raster_fwd_kernel and raster_fill_kernel captures (raster_prep_kernel: < 5 MB, not captured)"
        json.dump(d, open(tj, "w"), indent=1)
