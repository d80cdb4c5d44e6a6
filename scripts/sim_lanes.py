"""CPU simulation of the forward rasterizer's lane balance (no GPU): how many warp rounds the evaluation and the sorted
insertion take per render under different pixel -> lane assignments and face orders.  Scratch tool behind the round-2
redesign of raster_fwd_kernel; it does not touch the product or the oracle.
usage: sim_lanes.py [renders] [template]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from acfm_video_3d_reconstruction_b200 import synthetic
from oracle import pt3d_oracle as orc

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
name = sys.argv[2] if len(sys.argv) > 2 else "bird"
S, K = 256, 20
blur = orc.BLUR_SOFT
sqb = np.sqrt(blur)
wl = synthetic.Workload(name, 64, 8, 32, S, seed=0)
sel = np.linspace(0, 511, R).astype(int)
X = wl.mean_v[None].repeat(len(sel), 1, 1).numpy()
cams = wl.cams.numpy()[sel]
ndc = orc.view(orc.project(X, cams, 5.0), yflip=True).astype(np.float64)
faces = wl.faces.numpy()
pix = 1.0 - (2 * np.arange(S) + 1) / S  # PixToNdc(S-1-i)

def region_data(v, order_mode):
    """yields per live region: cand[nf,1024] (filter pass), ok[nf,1024] (exact pass), pz[nf,1024], in processing order"""
    fv = v[faces]  # F,3,3
    x, y, z = fv[..., 0], fv[..., 1], fv[..., 2]
    bx0, bx1, by0, by1 = x.min(1) - sqb, x.max(1) + sqb, y.min(1) - sqb, y.max(1) + sqb
    zc = z.sum(1)
    zlo, zhi = v[:, 2].min(), v[:, 2].max()
    bucket = np.clip(((zc - 3 * zlo) * (64 / (3 * (zhi - zlo)))).astype(int), 0, 63)
    for ry in range(S // 32):
        for rx in range(S // 32):
            px, py = pix[rx * 32:(rx + 1) * 32], pix[ry * 32:(ry + 1) * 32]
            keep = ~((px[-1] > bx1) | (px[0] < bx0) | (py[-1] > by1) | (py[0] < by0))
            idx = np.nonzero(keep)[0]
            if len(idx) == 0:
                continue
            if order_mode == "bucket":
                idx = idx[np.argsort(bucket[idx], kind="stable")]
            elif order_mode == "centroid":
                idx = idx[np.argsort(zc[idx], kind="stable")]
            elif order_mode == "zmin":
                idx = idx[np.argsort(z[idx].min(1), kind="stable")]
            PX, PY = np.meshgrid(px, py)  # [32 rows, 32 cols]
            P = np.stack([PX.ravel(), PY.ravel()], -1)  # 1024,2 ; pixel id = row*32+col
            a, b, c = fv[idx, 0, :2], fv[idx, 1, :2], fv[idx, 2, :2]
            za, zb, zc_ = fv[idx, 0, 2], fv[idx, 1, 2], fv[idx, 2, 2]
            def edge(p, a, b):
                return (p[..., 0] - a[..., 0]) * (b[..., 1] - a[..., 1]) - (p[..., 1] - a[..., 1]) * (b[..., 0] - a[..., 0])
            den = edge(c, a, b)[:, None] + 1e-8
            Pn = P[None]
            w0 = edge(Pn, b[:, None], c[:, None]) / den
            w1 = edge(Pn, c[:, None], a[:, None]) / den
            w2 = edge(Pn, a[:, None], b[:, None]) / den
            pz = w0 * za[:, None] + w1 * zb[:, None] + w2 * zc_[:, None]
            def segd(p, a, b):
                ba = b - a
                l2 = (ba ** 2).sum(-1)
                t = np.clip(((p - a) * ba).sum(-1) / l2, 0, 1)
                q = a + t[..., None] * ba
                return ((p - q) ** 2).sum(-1)
            def lined(p, a, b, opp):  # signed distance to the line ab, positive on the side of opp
                ba = b - a
                nrm = np.stack([-ba[..., 1], ba[..., 0]], -1) / np.linalg.norm(ba, axis=-1, keepdims=True)
                s = ((p - a) * nrm).sum(-1)
                so = ((opp - a) * nrm).sum(-1)
                return s * np.sign(so)
            d = np.minimum(np.minimum(segd(Pn, a[:, None], b[:, None]), segd(Pn, a[:, None], c[:, None])), segd(Pn, b[:, None], c[:, None]))
            inside = (w0 > 0) & (w1 > 0) & (w2 > 0)
            inb = ~((Pn[..., 0] > bx1[idx, None]) | (Pn[..., 0] < bx0[idx, None]) | (Pn[..., 1] > by1[idx, None]) | (Pn[..., 1] < by0[idx, None]))
            ok = inb & (pz >= 0) & (inside | (d < blur))
            T = 1.001 * sqb + 2e-5
            s0 = lined(Pn, b[:, None], c[:, None], a[:, None]); s1 = lined(Pn, c[:, None], a[:, None], b[:, None]); s2 = lined(Pn, a[:, None], b[:, None], c[:, None])
            cand = inb & (s0 >= -T) & (s1 >= -T) & (s2 >= -T)
            yield cand, ok, pz, idx

def shifts_for_pixel(zs, fids):
    """sorted insertion with K truncation: returns list of shift counts (one per accepted insert; -1 = rejected when full)"""
    lst = []
    out = []
    for zf in zip(zs, fids):
        if len(lst) == K:
            if zf > lst[-1]:
                out.append(-1); continue
            lst.pop()
        pos = len(lst)
        while pos > 0 and lst[pos - 1] > zf:
            pos -= 1
        out.append(len(lst) - pos)
        lst.insert(pos, zf)
    return out

def simulate(order_mode):
    tot = dict(pairs=0, okpairs=0, ev_rounds_tile=0, ev_rounds_sorted=0, ev_ideal=0,
               ins_rounds_tile=0, ins_cost_tile=0, ins_work=0, ins_cost_decoupled=0, ins_cost_sorted=0, ins_cost_sorted_dec=0, appends=0, rejects=0, inserts=0,
               ins_rounds_dec=0)
    for r in range(len(sel)):
        for cand, ok, pz, idx in region_data(ndc[r], order_mode):
            nf = cand.shape[0]
            ncand = cand.sum(0)  # per pixel
            tot["pairs"] += int(ncand.sum()); tot["okpairs"] += int(ok.sum())
            # per-pixel sequences: for every candidate, (passes, shift)
            seqs = []
            tile_order = {}
            if order_mode == "tile":
                for ty in range(8):
                    for tx in range(4):
                        pxs = [(ty * 4 + yy) * 32 + tx * 8 + xx for yy in range(4) for xx in range(8)]
                        tile_order[(ty, tx)] = np.argsort(pz[:, pxs].mean(1), kind="stable")
            for p in range(1024):
                ci = np.nonzero(cand[:, p])[0]
                if order_mode == "tile":
                    o = tile_order[((p // 32) // 4, (p % 32) // 8)]
                    ci = o[cand[o, p]]
                okp = ok[ci, p]
                sh = shifts_for_pixel(pz[ci[okp], p], idx[ci[okp]])
                full = np.full(len(ci), -2)  # -2: fails the exact test
                full[okp] = sh
                seqs.append(full)
            def cost_group(pxs):
                """pxs: list of pixel ids forming one warp's lanes. returns eval rounds, insert cost current scheme, insert cost decoupled"""
                n = [len(seqs[p]) for p in pxs]
                rounds = max(n) if n else 0
                ins_cur = 0; ins_rounds = 0
                for i in range(rounds):
                    sh = [seqs[p][i] for p in pxs if i < len(seqs[p])]
                    act = [s for s in sh if s >= -1]
                    if act:
                        ins_rounds += 1
                        ins_cur += 1 + max((max(s, 0) + 1) // 2 for s in act)  # 1 base trip + shift trips (2 elements per trip)
                # decoupled: each lane's passing results compacted
                comp = [[s for s in seqs[p] if s >= -1] for p in pxs]
                m = max((len(c) for c in comp), default=0)
                ins_dec = 0
                for i in range(m):
                    act = [c[i] for c in comp if i < len(c)]
                    ins_dec += 1 + max((max(s, 0) + 1) // 2 for s in act)
                return rounds, ins_cur, ins_dec, ins_rounds, m
            def policy(pxs, T, cE=110, cS=20):
                """state machine: E phase when >= T lanes are ready to evaluate (or nobody shifts), else one S trip for the shifting lanes"""
                i = [0] * len(pxs); pend = [0] * len(pxs)
                nEp = nSp = 0
                while True:
                    ready = [k for k in range(len(pxs)) if pend[k] == 0 and i[k] < len(seqs[pxs[k]])]
                    shifting = [k for k in range(len(pxs)) if pend[k] > 0]
                    if not ready and not shifting: break
                    if not shifting or len(ready) >= T:
                        nEp += 1
                        for k in ready:
                            v = seqs[pxs[k]][i[k]]; i[k] += 1
                            if v > 0: pend[k] = (v + 1) // 2
                    else:
                        nSp += 1
                        for k in shifting: pend[k] -= 1
                return nEp, nSp
            for T in (1, 4, 8, 12, 16):
                for ty in range(8):
                    for tx in range(4):
                        pxs = [(ty * 4 + yy) * 32 + tx * 8 + xx for yy in range(4) for xx in range(8)]
                        e, sp = policy(pxs, T)
                        tot.setdefault(f"polE{T}", 0); tot.setdefault(f"polS{T}", 0)
                        tot[f"polE{T}"] += e; tot[f"polS{T}"] += sp
            # tile grouping
            for ty in range(8):
                for tx in range(4):
                    pxs = [(ty * 4 + yy) * 32 + tx * 8 + xx for yy in range(4) for xx in range(8)]
                    a, b, c, d, m = cost_group(pxs)
                    tot["ev_rounds_tile"] += a; tot["ins_cost_tile"] += b; tot["ins_cost_decoupled"] += c; tot["ins_rounds_tile"] += d; tot["ins_rounds_dec"] += m
            # sorted grouping (by candidate count, covered pixels only)
            order = np.argsort(-ncand, kind="stable")
            order = [p for p in order if ncand[p] > 0]
            for g0 in range(0, len(order), 32):
                a, b, c, d, m = cost_group(order[g0:g0 + 32])
                tot["ev_rounds_sorted"] += a; tot["ins_cost_sorted"] += b; tot["ins_cost_sorted_dec"] += c
            tot["ev_ideal"] += (int(ncand.sum()) + 31) // 32
            for s in seqs:
                for v in s:
                    if v == -1: tot["rejects"] += 1
                    elif v == 0: tot["appends"] += 1; tot["inserts"] += 1; tot["ins_work"] += 1
                    elif v > 0: tot["inserts"] += 1; tot["ins_work"] += 1 + (v + 1) // 2
    n = len(sel)
    print(f"--- order={order_mode}  ({n} renders of {name})")
    print(f"filter pairs/render {tot['pairs']/n:.0f}  exact-pass {tot['okpairs']/n:.0f}  rejects(full) {tot['rejects']/n:.0f} appends {tot['appends']/n:.0f} inserts {tot['inserts']/n:.0f}")
    print(f"eval rounds/render: tile {tot['ev_rounds_tile']/n:.0f} (lanes {tot['pairs']/tot['ev_rounds_tile']:.1f})  sorted {tot['ev_rounds_sorted']/n:.0f} (lanes {tot['pairs']/tot['ev_rounds_sorted']:.1f})  ideal {tot['ev_ideal']/n:.0f}")
    w = tot["ins_work"]
    print(f"insert trips/render (warp level): tile-coupled {tot['ins_cost_tile']/n:.0f} (lanes {w/tot['ins_cost_tile']:.1f})  tile-decoupled {tot['ins_cost_decoupled']/n:.0f} ({w/tot['ins_cost_decoupled']:.1f})"
          f"  sorted-coupled {tot['ins_cost_sorted']/n:.0f} ({w/tot['ins_cost_sorted']:.1f})  sorted-decoupled {tot['ins_cost_sorted_dec']/n:.0f} ({w/tot['ins_cost_sorted_dec']:.1f})  ideal {w/32/n:.0f}")
    for T in (1, 4, 8, 12, 16):
        e, sp = tot[f"polE{T}"] / n, tot[f"polS{T}"] / n
        print(f"policy T={T}: E phases {e:.0f}  S trips {sp:.0f}  cost {(e * 110 + sp * 20) / 1e3:.0f}k  (current {(tot['ev_rounds_tile'] * 110 + tot['ins_cost_tile'] * 20) / n / 1e3:.0f}k)")
    print(f"insert rounds: coupled {tot['ins_rounds_tile']/n:.0f} decoupled {tot['ins_rounds_dec']/n:.0f}")

for mode in (sys.argv[3:] or ["bucket", "centroid", "zmin"]):
    simulate(mode)
