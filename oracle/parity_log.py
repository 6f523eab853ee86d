"""ORACLE helper (test infrastructure): measured parity errors -> a JSON-lines log.

Every `-m gpu` parity test records what it MEASURED (max-abs, rel-L2, outlier fraction ...) per case and tensor, not
only whether it passed.  On the GPU box the log lands in gpurun_out/ (the only directory that travels back); the fold
script dev/fold_parity_log.py turns it into profiles/r02_parity_errors.json, from which the per-case thresholds of the
tests are derived (3x the measured value, <= 1e-4 wherever no argmin / floor flip is involved).
"""
import json
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def log_path():
    return os.environ.get("DD_PARITY_LOG", os.path.join(ROOT, "gpurun_out", "r02_parity_errors.jsonl"))


def record(case, tensor, **metrics):
    """Append one measurement; never raises (a read-only tree must not fail a parity test)."""
    try:
        path = log_path()
        os.makedirs(os.path.dirname(path), exist_ok=True)
        row = {"case": str(case), "tensor": str(tensor), "t": round(time.time(), 1)}
        row.update({k: (float(v) if isinstance(v, (int, float)) else v) for k, v in metrics.items()})
        with open(path, "a") as fh:
            fh.write(json.dumps(row) + "\n")
    except Exception:
        pass
    return metrics


def fold(path=None):
    """{case: {tensor: {metric: worst value}}} over every row of the log."""
    table = {}
    with open(path or log_path()) as fh:
        for line in fh:
            row = json.loads(line)
            slot = table.setdefault(row.pop("case"), {}).setdefault(row.pop("tensor"), {})
            row.pop("t", None)
            for k, v in row.items():
                slot[k] = max(slot[k], v) if isinstance(v, (int, float)) and k in slot else v
    return table
