"""Writes tests/golden/lateral_golden.json from oracle/lateral.py (run in the build container)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_oracle_lateral import _run_sequence  # noqa: E402

out = {str(s): _run_sequence(s) for s in (100, 200, 300)}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "lateral_golden.json")
json.dump(out, open(p, "w"), indent=0)
print("wrote", p, os.path.getsize(p))
