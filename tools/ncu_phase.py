"""Split an `ncu --page source --csv --print-source sass` export into phases (between BAR.SYNCs) and print,
per phase, the stall samples, executed warp instructions and the top stall reasons.  Dev tool."""
import csv, sys, collections

def main(path, which=0, top=12):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    s = starts[which]
    e = starts[which + 1] if which + 1 < len(starts) else len(rows)
    print(rows[s][1][:100])
    hdr = rows[s + 1]
    col = {n: i for i, n in enumerate(hdr)}
    body = rows[s + 2:e]
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    phases, cur = [], []
    for r in body:
        cur.append(r)
        if "BAR.SYNC" in r[col["Source"]]:
            phases.append(cur); cur = []
    phases.append(cur)
    tot_s = sum(int(r[col["# Samples"]]) for r in body)
    tot_i = sum(int(r[col["Instructions Executed"]]) for r in body)
    print(f"total samples {tot_s}, warp instr {tot_i/1e6:.1f} M")
    for k, ph in enumerate(phases):
        ss = sum(int(r[col["# Samples"]]) for r in ph)
        ii = sum(int(r[col["Instructions Executed"]]) for r in ph)
        st = collections.Counter()
        for r in ph:
            for n in stall_cols:
                st[n] += int(r[col[n]])
        tops = ", ".join(f"{n[6:]} {v*100/max(ss,1):.0f}%" for n, v in st.most_common(5))
        print(f"phase {k}: {len(ph)} sass lines, samples {ss*100/tot_s:.1f}%, instr {ii*100/tot_i:.1f}% ({ii/1e6:.1f} M) | {tops}")
    hot = sorted(body, key=lambda r: -int(r[col["# Samples"]]))[:top]
    for r in hot:
        print(f'{int(r[col["# Samples"]])*100/tot_s:5.1f}%  {r[col["Source"]].strip()[:70]}')

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
