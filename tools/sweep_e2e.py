#!/usr/bin/env python
"""measurement aid: end-to-end (host buffers) throughput vs the host-path chunk size"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for wl in sys.argv[2].split(","):
    for mb in sys.argv[3:]:
        env = dict(os.environ); env["FDC_HOST_CHUNK_MB"] = mb
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--no-cpu", "--steps", "5", "--warmup", "3"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            line = "%s host chunk %3s MiB: e2e %.0f Ms/s  (h2d %.0f MB + d2h %.0f MB per step)  value %.0f" % (wl, mb, d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"] / 1e6, d["e2e"]["d2h_bytes_per_step"] / 1e6, d["value"])
        except Exception:
            line = "%s %s FAILED %s" % (wl, mb, (r.stderr or r.stdout)[-300:])
        print(line)
        open(sys.argv[1], "a").write(line + "\n")
