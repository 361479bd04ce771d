"""Per-CUDA-source-line aggregates (instructions executed, average active lanes, stall samples) from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::N > file.csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, lines = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "" and r[2] == "-": lines.append((cur_file, r))
ie, it, ismp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ilsb = hdr.index("stall_long_sb")
I = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
tot = sum(I(r[ie]) for _, r in lines); tots = sum(I(r[ismp]) for _, r in lines); tott = sum(I(r[it]) for _, r in lines)
print(f"warp-inst {tot:.3e}  thread-inst {tott:.3e}  avg lanes {tott/tot:.2f}  samples {tots}")
for f, r in sorted(lines, key=lambda x: -I(x[1][ie]))[:top]:
    e, t = I(r[ie]), I(r[it])
    if e == 0: continue
    print(f"{e/tot*100:5.1f}% inst  lanes {t/e:5.1f}  smp {I(r[ismp])/tots*100:5.1f}% (long_sb {I(r[ilsb])/tots*100:4.1f}%)  {f}:{r[0]:>4s}: {r[1].strip()[:100]}")
