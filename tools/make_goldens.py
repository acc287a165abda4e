"""Documents how tests/golden/reference_vectors.json was produced: the numbers are the hard-coded expectations of the
reference's own unit tests (math_test.py:27-131, io_test.py:499-505, benchmarks/humanoid/README.md), transcribed by hand
because those tests cannot run here (they import warp and mujoco).  This script re-checks that every transcribed literal
still appears verbatim in the reference sources when /root/reference is mounted."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/mujoco_warp/_src"

if __name__ == "__main__":
  g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
  src = open(os.path.join(REF, "math_test.py")).read()
  missing = []
  for case in g["closest_segment_to_segment_points"]:
    for key in ("best_a", "best_b"):
      for v in case[key]:
        lit = repr(float(v)).rstrip("0").rstrip(".") if float(v) != int(v) else str(float(v))
        if lit not in src and str(v) not in src:
          missing.append((case["ref"], key, v))
  print("literals not found verbatim:", missing)
  sys.exit(1 if missing else 0)
