"""Dev tool: per-range opcode/instruction breakdown of an ncu source-page CSV.  usage: ncu_ranges.py file.csv kernel_idx a:b[:label] ..."""
import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]
k=int(sys.argv[2]); s=starts[k]; e=starts[k+1] if k+1<len(starts) else len(rows)
hdr=rows[s+1]; col={n:i for i,n in enumerate(hdr)}; body=rows[s+2:e]
tot=sum(int(r[col["Instructions Executed"]]) for r in body); tots=sum(int(r[col["# Samples"]]) for r in body)
if len(sys.argv)==3:
    # dump line index, instr executed, samples, source
    for i,r in enumerate(body):
        print(i, r[col["Instructions Executed"]], r[col["# Samples"]], r[col["Source"]].strip())
    sys.exit()
for spec in sys.argv[3:]:
    p=spec.split(':'); a,b=int(p[0]),int(p[1]); label=p[2] if len(p)>2 else spec
    ii=sum(int(r[col["Instructions Executed"]]) for r in body[a:b]); ss=sum(int(r[col["# Samples"]]) for r in body[a:b])
    ops=collections.Counter()
    for r in body[a:b]:
        src=r[col["Source"]].split()
        op=src[1] if src[0].startswith('@') else src[0]
        ops[op.split('.')[0]]+=int(r[col["Instructions Executed"]])
    print(f"{label}: instr {ii/1e6:.1f}M ({100*ii/tot:.1f}%) samples {100*ss/tots:.1f}%", [(k,round(100*v/ii,1)) for k,v in ops.most_common(14)])
