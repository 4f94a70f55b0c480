"""Top stall locations from `ncu -i X.ncu-rep --page source --csv`. Usage: python profiles/top_stalls.py file.csv [N]"""
import csv, sys
path = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ci['# Samples']] or 0) for r in body)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print(f"total samples {tot}")
for r in sorted(body, key=lambda r: -int(r[ci['# Samples']] or 0))[:n]:
    s = int(r[ci['# Samples']] or 0)
    top = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{100*s/tot:5.1f}%  {r[ci['Source']].strip()[:70]:70s} exec={r[ci['Instructions Executed']]:>8s} {top}")
