#!/usr/bin/env python
"""measurement aid: run bench.py over combinations of the FDC_* run-time switches and print one compact line per run.

    python tools/sweep.py OUT.txt WORKLOADS "K=V K=V" "K=V" ...      (each quoted argument = one environment)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out, wls, envs = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
    with open(out, "w") as fh:
        for wl in wls:
            for e in envs:
                env = dict(os.environ)
                extra = []
                for kv in e.split():
                    k, v = kv.split("=")
                    if k.startswith("--"):
                        extra += [k, v]
                    else:
                        env[k] = v
                r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--no-cpu", "--no-e2e", "--steps", "20",
                                    "--warmup", "3"] + extra, env=env, capture_output=True, text=True)
                try:
                    d = json.loads(r.stdout.strip().splitlines()[-1]); k = d["roofline"]["kernels"]
                    line = "%-5s %-44s value %6.0f Ms/s  ms/step %.4f  fwd %.3f  ext %.3f  frac %.3f" % (
                        wl, e, d["value"], d["ms_per_step"], k["forward_fft"]["ms"], k["channel_extract"]["ms"], d["roofline"]["path"]["frac"])
                except Exception:
                    line = "%-5s %-44s FAILED %s" % (wl, e, (r.stderr or r.stdout)[-300:])
                print(line); fh.write(line + "\n"); fh.flush()


if __name__ == "__main__":
    main()
