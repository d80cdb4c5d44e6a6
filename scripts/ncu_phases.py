"""Group the per-line output of ncu_lines.py into the phases of raster_fwd_kernel (line anchors read from the source)."""
import re, sys
# second argument: the raster_fwd.cu the profiled library was built from (default: the working tree) — the line anchors must
# come from the same source as the report
src = open(sys.argv[2] if len(sys.argv) > 2 else "acfm_video_3d_reconstruction_b200/csrc/raster_fwd.cu").read().splitlines()
def find(pat):
    for i, l in enumerate(src, 1):
        if pat in l: return i
    raise KeyError(pat)
anchors = [("fill helpers", find("void warp_fill_frag(")), ("eval", find("the exact per-(pixel, face) test")), ("setup_face", find("// Full per-face set-up")),
           ("insert", find("// ---- per-pixel K-nearest SET")), ("prologue+bbox", find("template <int NWARPS, typename IdxT, int KT>\n__global__".split("\n")[0])),
           ("cull+bucket", find("// ---- 2. cull all faces")), ("records", find("// ---- 3. per-region face records")),
           ("tile head", find("// ---- 4. warps pull")), ("scan", find("// (a) scan")), ("filter", find("// (b) filter")),
           ("evalloop", find("// (c) evaluate")), ("overflow", find("// overflow: region faces")), ("blend+out", find("// ---- (d) depth order, blend")),
           ("end", find("// choose the CTA size"))]
acc = {}
for l in open(sys.argv[1]):
    m = re.match(r"(\S+):\s*(\d+)\s+instr\s+([\d.]+)%\s+samples\s+([\d.]+)%\s+thr/inst\s+([\d.]+)", l)
    if not m: continue
    f, ln, i, s, t = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))
    g = f
    if f == "raster_fwd.cu":
        g = "?"
        for (name, a), (_, b) in zip(anchors, anchors[1:]):
            if a <= ln < b: g = name
    elif f == "common.cuh": g = "arith wrappers (common.cuh)" if ln < 70 else "tma/other common"
    elif f == "raster_common.cuh": g = "fdiv (raster_common)"
    a = acc.setdefault(g, [0, 0, 0]); a[0] += i; a[1] += s; a[2] += i * t
for g, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{g:35s} instr {a[0]:5.1f}%  samples {a[1]:5.1f}%  lanes {a[2] / max(a[0], 1e-9):4.1f}")
