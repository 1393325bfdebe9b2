"""Extract one kernel's SASS from a .so and print instruction statistics.  usage: sass_fn.py lib.so 'k_sweep_viewILi512ELi16ELb0' [--dump]"""
import re, subprocess, sys, collections
so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", b)
    ops = collections.Counter()
    for a, t in ins:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", t.strip())
        ops[m.group(2).split(".")[0] if m else t] += 1
    # hot loop = from the first SYNCS.PHASECHK back-edge region: find the backward branch with the largest span that contains PHASECHK
    addr = [int(a, 16) for a, _ in ins]
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr[i]:
                body = [x for x in ins if tgt <= int(x[0], 16) <= addr[i]]
                if any("PHASECHK" in x[1] for x in body) and any("UBLKCP" in x[1] for x in body):
                    if best is None or len(body) < len(best):
                        best = body
    print(name, "total", len(ins), "S2R", ops["S2R"] + ops["S2UR"], "hot-loop", len(best) if best else None)
    if best:
        hops = collections.Counter()
        for a, t in best:
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", t.strip())
            hops[m.group(2).split(".")[0]] += 1
        print("  ", dict(hops.most_common(14)))
    if "--dump" in sys.argv and best:
        for a, t in best:
            print(a, t)
