#!/usr/bin/env python3
"""profiles/sass_excerpt.py [obj] -- mnemonic counts of every eri_tile_kernel instance in the sm_100a cubin (cuobjdump -sass) and
excerpts of the (ps|ss) RHF kernel: the TMA bulk copies of a tile, the mbarrier wait, the speculative primitive screen issued
inside the evaluation, the far / near split of the root evaluation, the reds that remain per quartet."""
import re, subprocess, sys, collections
obj = sys.argv[1] if len(sys.argv) > 1 else "unomol_b200/csrc/build/eri_tile_classes.o"
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = collections.OrderedDict(); cur = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m and cur: funcs[cur].append((m.group(1), m.group(2).strip()))
print("profiles/r2_sass_tile_excerpt.txt -- cuobjdump -sass of %s (sm_100a cubin), round 2 final state (profiles/sass_excerpt.py)" % obj)
print("mnemonic counts per kernel: UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, REDG = red.global.add.f64,")
print("DFMA/DMUL/DADD = FP64 pipe, MUFU.RSQ64H / RCP64H = FP64 rsqrt / reciprocal seeds, LDS/STS = shared memory, BSSY = divergence points\n")
keys = ["UBLKCP", "SYNCS", "REDG", "DFMA", "DMUL", "DADD", "DSETP", "MUFU.RSQ64H", "MUFU.RCP64H", "LDS", "STS", "LDG", "LD.E", "BRA", "BSSY"]
for f, ins in funcs.items():
    m = re.search(r"eri_tile_kernelILi(\d)ELi(\d)ELi(\d)ELi(\d)ELi(\d)", f)
    if not m: continue
    c = collections.Counter()
    for _, t in ins:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        for k in keys:
            if op.startswith(k): c[k] += 1
    print("eri_tile_kernel<%s%s|%s%s, nspin %s>: %d instructions | %s" % (*m.groups(), len(ins), "  ".join("%s %d" % (k, c[k]) for k in keys)))
def show(f, title, pred, before, after, nmax=1):
    ins = funcs[f]; shown = 0
    for i, (a, t) in enumerate(ins):
        if pred(t):
            print("\n--- " + title)
            for a2, t2 in ins[max(0, i - before): i + after]: print("        /*%s*/   %s ;" % (a2, t2))
            shown += 1
            if shown >= nmax: break
ps = [f for f in funcs if "eri_tile_kernelILi1ELi0ELi0ELi0ELi1" in f][0]
show(ps, "(ps|ss) RHF kernel: TMA bulk copies of one tile (pair records, ket counts) by the elected thread", lambda t: t.startswith("UBLKCP"), 8, 12)
show(ps, "mbarrier wait on the stage that the TMA copies fill", lambda t: "SYNCS.PHASECHK" in t, 3, 4)
show(ps, "primitive quartet: bra record from the shared-memory stage, SPECULATIVE screen of the next bra primitive (@!P LDS.64 / DFMA / DSETP) interleaved with |PQ|^2, the far/near decision without a division (DMUL by 35, DSETP.GT), far branch = one MUFU.RSQ64H + cubic step, near branch = rsqrt(p+q) + Taylor row of F_0/F_1 from the shared-memory table", lambda t: "MUFU.RSQ64H" in t, 45, 75)
show(ps, "reds that remain per quartet (K[b_j,c], K[b_j,d]) after the per-ket register accumulation", lambda t: t.startswith("REDG") or " REDG" in t, 6, 6)
