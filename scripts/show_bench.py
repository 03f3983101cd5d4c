"""Print the interesting fields of a bench.py JSON line (last line of the file)."""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
r = d.get("roofline") or {}
print(f"{d['config']['workload']} N={d['n_gpus']}: value {d['value']:.1f} q/s  {d['ms_per_step']:.4f} ms/step  e2e {d['e2e']['value']:.1f} q/s "
      f"({d['e2e'].get('ms_per_step', 0):.4f} ms)  launches/step {d['gpu_launches'] / d['steps']:.2f}")
if r:
    print(f"  roofline: {r['achieved']:.0f} GB/s = {r['frac']:.3f}x measured peak, scan {r['scan_ms_per_search']:.4f} ms, share {r['scan_share_of_step']:.3f}")
print("  clocks:", d.get("clocks"))
print("  parity:", d.get("parity"))
if d.get("cpu_baseline"):
    print("  cpu_baseline:", round(d["cpu_baseline"]["value"], 2), "q/s on", d["cpu_baseline"]["cores"], "cores")
for c in d.get("configs") or []:
    rf = c.get("roofline") or {}
    extra = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in c.items() if k not in ("config", "roofline", "note", "scan_arithmetic")}
    print("  ", c["config"], extra, ("frac %.3f (whole %.3f)" % (rf.get("frac", 0), rf.get("frac_whole_search", 0))) if rf else "")
