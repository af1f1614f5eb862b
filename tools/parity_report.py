"""Measured parity distances of every golden case on the GPU (the numbers the test tolerances are set from).

    python tools/parity_report.py [case ...]  > gpurun_out/parity_report.json
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    names = sys.argv[1:] or list(parity_cases.RUNNERS)
    rep = {}
    for n in names:
        t0 = time.time()
        try:
            rep[n] = parity_cases.RUNNERS[n](n, dev)
        except Exception as e:      # keep going: the report is diagnostic
            rep[n] = {"error": repr(e)}
        rep[n]["seconds"] = round(time.time() - t0, 1)
        print(n, json.dumps(rep[n]), file=sys.stderr, flush=True)
        torch.cuda.empty_cache()
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
