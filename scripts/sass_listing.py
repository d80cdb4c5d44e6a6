"""SASS evidence for profiles/: per kernel of libacfm_b200.so, the instruction count and the mnemonics that show which hardware
paths it uses (TMA bulk copies, mbarriers, programmatic dependent launch, native shared atomics, fp64 tensor cores ...).
usage: python scripts/sass_listing.py > profiles/sass_r05.txt      (cuobjdump -sass on the in-tree library; no GPU needed)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "acfm_video_3d_reconstruction_b200", "libacfm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = [("UBLKCP", "cp.async.bulk (TMA bulk copy)"), ("SYNCS", "mbarrier ops"), ("UTMA|UBLKPF", "other TMA"),
        ("PREEXIT", "griddepcontrol.launch_dependents"), ("ACQBULK", "griddepcontrol.wait"), ("ATOMS\\.ADD|ATOMS\\.POPC\\.INC|ATOMS\\.MAX|ATOMS\\.MIN", "native shared atomics"),
        ("ATOMS\\.CAST|ATOMS\\.CAS", "shared CAS loops"), ("RED\\.|ATOMG|ATOM\\.", "global atomics / reductions"), ("DMMA", "fp64 tensor core"),
        ("DFMA|DADD|DMUL", "scalar fp64"), ("MUFU", "special function unit"), ("F2I|I2F", "XU conversions"), ("FMNMX3|VIMNMX3", "3-input min/max"),
        ("SHFL", "shuffles"), ("MATCH|REDUX", "warp match / reduce"), ("LDS", "shared loads"), ("STS", "shared stores"), ("LDG|LD\\.E", "global loads"),
        ("STG|ST\\.E", "global stores"), ("LDL|STL", "local memory (spills)"), ("BAR\\.", "CTA barriers")]
cur, counts, total = None, None, 0
res = []
for l in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        if cur:
            res.append((cur, total, counts))
        cur, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and cur:
        total += 1
        op = m.group(1)
        for k, _ in KEYS:
            if re.match("(?:" + k + ")", op):
                counts[k] += 1
if cur:
    res.append((cur, total, counts))
demangle = subprocess.run(["c++filt"], input="\n".join(r[0] for r in res), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass of libacfm_b200.so (sm_100a): static instruction counts per kernel and per mnemonic family")
print("# " + "; ".join(f"{k} = {d}" for k, d in KEYS))
for (name, tot, c), dn in sorted(zip(res, demangle), key=lambda x: -x[0][1]):
    dn = re.sub(r"\(anonymous namespace\)::", "", dn)
    dn = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", dn)
    print(f"{tot:6d}  {dn[:70]:70s} " + " ".join(f"{k.split('|')[0].replace(chr(92), '')}:{v}" for k, v in c.items() if v))
