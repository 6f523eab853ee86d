"""gpurun_out/r02_parity_errors.jsonl (written by the -m gpu tests through oracle/parity_log.py) -> profiles/r02_parity_errors.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import parity_log  # noqa: E402

if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else parity_log.log_path()
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_parity_errors.json")
    table = parity_log.fold(src)
    with open(dst, "w") as fh:
        json.dump(table, fh, indent=1, sort_keys=True)
    print(f"{sum(len(v) for v in table.values())} tensors of {len(table)} cases -> {dst}")
