"""Measured parity errors of the -m gpu tests, as a JSON artifact.

pytest swallows prints, so every GPU parity test records what it measured (and the tolerance it asserted) here:
`record("test name", "quantity", measured, tolerance)` merges into gpurun_out/r02_parity_errors.json on the GPU box
(gpurun_out/ is what travels back to the authoring container); the file is then committed as
profiles/r02_parity_errors.json.
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "gpurun_out", "r02_parity_errors.json")


def record(test: str, quantity: str, measured, tolerance=None, **extra):
    try:
        os.makedirs(os.path.dirname(PATH), exist_ok=True)
        data = {}
        if os.path.exists(PATH):
            with open(PATH) as f:
                data = json.load(f)
        entry = {"measured": float(measured)}
        if tolerance is not None:
            entry["tolerance"] = float(tolerance)
        entry.update(extra)
        data.setdefault(test, {})[quantity] = entry
        with open(PATH, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass  # read-only checkout: the assertion in the test is what counts
