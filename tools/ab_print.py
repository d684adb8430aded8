"""Pretty-print tools/ab.py JSON lines from stdin (one row per variant)."""
import json
import sys

for l in sys.stdin:
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    k = d.get("kernel_us", {})
    print(f"{d['variant']:16s} {d['workload']:4s} par={tuple((d.get('parity') or {}).values())} " + " ".join(f"{n[:6]}={v:6.1f}" for n, v in k.items())
          + f" draw={d.get('draw_us')} wall={d.get('wall_ms_per_step')} {d.get('e2e_ms', '')} {d.get('error', '')}")
