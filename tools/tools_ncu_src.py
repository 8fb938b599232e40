"""Scratch: aggregate an `ncu --page source --csv` dump by opcode and stall reason."""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
npts_per_thread = float(sys.argv[2]) if len(sys.argv) > 2 else 4
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); samples = collections.Counter(); total_inst = 0; nthreads = 0
def num(x):
    try: return int(float(x))
    except Exception: return 0
body = [r for r in rows[2:] if len(r) >= len(hdr) and r[ix["Address"]] != "Address"]
for r in body:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    if not m: continue
    op = m.group(2)
    n = num(r[ix["Instructions Executed"]]); s = num(r[ix["# Samples"]])
    tot[op] += n; samples[op] += s; total_inst += n
first = max(num(r[ix["Instructions Executed"]]) for r in body)
print("warps launched ~", first, " total warp instr", total_inst, " instr per thread", total_inst / first, " per point", total_inst / first / npts_per_thread)
for op, n in tot.most_common(28):
    print(f"{op:10s} {n/first:8.2f} per thread   samples {samples[op]}")
st = collections.Counter()
for r in body:
    for h in hdr:
        if h.startswith("stall_"): st[h] += num(r[ix[h]])
print(st.most_common(8))
