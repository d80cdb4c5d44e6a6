"""CPU estimate (no GPU) of what a tile-level depth cull would remove from raster_fwd_kernel: a face whose nearest vertex is
farther than the K-th nearest fragment of EVERY pixel of an 8x4 tile (all 32 sets full) cannot contribute to the tile.
Counts, over the live tiles of a few C2 renders, the (tile, face) filter iterations and the per-pixel exact evaluations that
such a test would skip when the faces come front to back in 64 depth buckets, as the kernel orders them.
usage: sim_hiz.py [renders] [template]"""
import os, sys
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sim_lanes.py")).read()
exec(src[:src.index("def shifts_for_pixel")])

Q = int(os.environ.get("Q", 32))
tot = dict(pairs=0, pairs_skip=0, evals=0, evals_skip=0, evals_far=0, frags=0, tiles=0, drains=0, rounds=0)
for r in range(len(sel)):
    v = ndc[r]
    zmin_all = v[faces][..., 2].min(1)
    for cand, ok, pz, idx in region_data(v, "bucket"):
        nf = len(idx)
        zmin = zmin_all[idx]
        cand = cand.reshape(nf, 32, 32); ok = ok.reshape(nf, 32, 32); pz = pz.reshape(nf, 32, 32)
        for ty in range(8):
            for tx in range(4):
                c = cand[:, ty * 4:ty * 4 + 4, tx * 8:tx * 8 + 8].reshape(nf, 32)
                o = ok[:, ty * 4:ty * 4 + 4, tx * 8:tx * 8 + 8].reshape(nf, 32)
                z = pz[:, ty * 4:ty * 4 + 4, tx * 8:tx * 8 + 8].reshape(nf, 32)
                hit = c.any(1)
                if not hit.any():
                    continue
                tot["tiles"] += 1
                kth = np.full(32, np.inf)
                lists = [[] for _ in range(32)]
                qn = np.zeros(32, int)
                tile_far = np.inf
                for j in np.nonzero(hit)[0]:
                    tot["pairs"] += 1
                    ne = int(c[j].sum())
                    tot["evals"] += ne
                    if zmin[j] > tile_far:
                        tot["pairs_skip"] += 1
                        tot["evals_skip"] += ne
                        continue
                    for l in np.nonzero(c[j])[0]:
                        if z[j, l] > kth[l]:
                            tot["evals_far"] += 1
                        if o[j, l] and z[j, l] <= kth[l]:
                            lists[l].append(z[j, l])
                            if len(lists[l]) > K:
                                lists[l].remove(max(lists[l]))
                            if len(lists[l]) == K:
                                kth[l] = max(lists[l])
                    qn += c[j]
                    if qn.max() >= Q:          # the kernel refreshes the bound after a drain only
                        tot["rounds"] += qn.max()
                        qn[:] = 0
                        tile_far = kth.max()
                        tot["drains"] += 1
                tot["rounds"] += qn.max()
                tot["frags"] += sum(len(x) for x in lists)
print(tot)
print("filter iterations skipped: %.1f %%   exact evaluations skipped: %.1f %% (already leaving after the depth: %.1f %%)" % (
    100 * tot["pairs_skip"] / tot["pairs"], 100 * tot["evals_skip"] / tot["evals"], 100 * tot["evals_far"] / tot["evals"]))
