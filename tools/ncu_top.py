"""Summarise an ncu --page source --csv dump: top stall instructions, shared-memory conflicts."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
print("total samples", tot)
top = sorted(range(len(data)), key=lambda i: -f(data[i], '# Samples'))[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for i in sorted(top):
    r = data[i]
    st = sorted(((f(r, s), s) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {r[ix['Source']].strip()[:60]:60s} samp={f(r,'# Samples'):8.0f} exec={f(r,'Instructions Executed'):10.0f} "
          f"wf={f(r,'L1 Wavefronts Shared'):10.0f} ideal={f(r,'L1 Wavefronts Shared Ideal'):10.0f} " + " ".join(f"{s}={v:.0f}" for v, s in st))
print("shared wavefronts", sum(f(r,'L1 Wavefronts Shared') for r in data), "ideal", sum(f(r,'L1 Wavefronts Shared Ideal') for r in data))
