"""Condense bench.py JSON lines from stdin into one short line each."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for l in sys.stdin:
    try:
        d = json.loads(l)
        r = d["roofline"]
        print("%s %.1f Gpt/s  launch %.4f ms  %.0f GB/s  frac %.3f" % (tag, d["value"] / 1e9, r["avg_launch_ms"], r["achieved"], r["frac"]))
    except Exception:
        if l.strip():
            print(tag, l.strip()[-300:])
