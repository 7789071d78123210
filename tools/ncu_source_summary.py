"""Summarise an `ncu --page source --csv` dump: stall totals, shared-memory wavefronts and the
hottest instructions (tools only).  usage: ncu_source_summary.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
def f(r, h):
    try: return float(r[ix[h]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in body)
print("instructions:", len(body), "samples:", tot, "warp insts:", sum(f(r, "Instructions Executed") for r in body))
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("stall totals:", sorted(((round(sum(f(r, h) for r in body) / tot * 100, 1), h) for h in st), reverse=True)[:8])
print("smem wavefronts:", sum(f(r, "L1 Wavefronts Shared") for r in body), "ideal:", sum(f(r, "L1 Wavefronts Shared Ideal") for r in body))
print("global sectors:", sum(f(r, "L2 Theoretical Sectors Global") for r in body), "ideal:", sum(f(r, "L2 Theoretical Sectors Global Ideal") for r in body))
# per-opcode totals
ops = {}
for r in body:
    op = r[ix["Source"]].split()[0] if not r[ix["Source"]].strip().startswith("@") else r[ix["Source"]].split()[1]
    o = ops.setdefault(op, [0, 0, 0, 0])
    o[0] += f(r, "Instructions Executed"); o[1] += f(r, "# Samples"); o[2] += f(r, "L1 Wavefronts Shared"); o[3] += f(r, "L1 Wavefronts Shared Excessive")
print("opcode: warp-insts, samples%, smem wavefronts, excessive")
for op, o in sorted(ops.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"  {op:22s} {o[0]:14.0f} {o[1] / tot * 100:6.1f}% {o[2]:14.0f} {o[3]:12.0f}")
print("hottest instructions:")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:top]:
    stalls = sorted(((f(r, h), h[6:]) for h in st), reverse=True)[:2]
    print(f"  {r[ix['Address']][-5:]} {f(r, '# Samples') / tot * 100:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {stalls[0][1]}:{stalls[0][0]:.0f} {stalls[1][1]}:{stalls[1][0]:.0f}")
