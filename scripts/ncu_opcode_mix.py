import csv, sys, collections, re
f = sys.argv[1]
rows = list(csv.reader(open(f)))
hdr = rows[1]
iA = hdr.index("Address"); iS = hdr.index("Source"); iI = hdr.index("Instructions Executed"); iSamp = hdr.index("# Samples")
iT = hdr.index("Thread Instructions Executed")
mix = collections.Counter(); samp = collections.Counter(); tot = 0; thr = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    s = r[iS].strip()
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    op = s.split()[0].split(".")[0] if s else "?"
    n = int(r[iI] or 0); mix[op] += n; tot += n; samp[op] += int(r[iSamp] or 0); thr += int(r[iT] or 0)
print("total warp instr", tot, "avg threads", thr / tot)
ts = sum(samp.values())
for op, n in mix.most_common(40):
    print(f"{op:10s} {n:14d} {100*n/tot:6.2f}%  samples {100*samp[op]/ts:6.2f}%")
