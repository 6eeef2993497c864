"""Top stall instructions of each kernel in an `ncu --page source --csv` export (SASS view)."""
import csv, io, sys
txt = open(sys.argv[1]).read()
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
secs = txt.split('"Kernel Name"')
for sec in secs[1:]:
    lines = sec.split('\n')
    print("=== kernel", lines[0][:90])
    rd = list(csv.DictReader(io.StringIO('\n'.join(lines[1:]))))
    tot = sum(int(r["# Samples"] or 0) for r in rd)
    print("instructions", len(rd), "total samples", tot)
    # by opcode class
    cls = {}
    for r in rd:
        op = r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1]
        op = op.split(".")[0]
        cls[op] = cls.get(op, 0) + int(r["# Samples"] or 0)
    print("by opcode:", sorted(cls.items(), key=lambda kv: -kv[1])[:14])
    order = sorted(range(len(rd)), key=lambda i: -int(rd[i]["# Samples"] or 0))
    for i in order[:topn]:
        r = rd[i]
        n = int(r["# Samples"] or 0)
        stalls = {k[6:]: int(v) for k, v in r.items() if k.startswith("stall_") and "Not Issued" not in k and v and int(v) > 0}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        print(f"{n:6d} {100*n/tot:5.1f}% [{i:5d}] {r['Source'].strip()[:90]:90s} exec {r['Instructions Executed']:>9s} {top}")
