"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line using nvdisasm -g line info.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> <function substring> [top]"""
import csv, re, subprocess, sys, collections
rep, kre, cubin, fsub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iA, iI, iS, iT = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
for i, r in enumerate(data):
    if r and r[0] == "Kernel Name":
        data = data[:i]
        break
base = int(data[0][iA], 16)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur, infn = {}, None, False
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l) or re.match(r"\s*//-+ \.text\.(\S+)", l)
    if m:
        infn = fsub in m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and infn:
        line_of[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot_i = tot_s = 0
for r in data:
    off = int(r[iA], 16) - base
    key = line_of.get(off, ("?", 0))
    a = agg[key]
    ii, ss, tt = int(r[iI] or 0), int(r[iS] or 0), int(r[iT] or 0)
    a[0] += ii; a[1] += ss; a[2] += tt
    for c in stall_cols:
        if r[c] and r[c] != "0":
            a[3][hdr[c]] += int(r[c])
    tot_i += ii; tot_s += ss
print(f"total warp-instr {tot_i}, samples {tot_s}, mapped lines {len(line_of)}")
src_cache = {}
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ", ".join(f"{k[6:]}:{v}" for k, v in a[3].most_common(3))
    print(f"{key[0]}:{key[1]:4d}  instr {100*a[0]/max(tot_i,1):5.1f}%  samples {100*a[1]/max(tot_s,1):5.1f}%  thr/inst {a[2]/max(a[0],1):4.1f}  [{st}]")
