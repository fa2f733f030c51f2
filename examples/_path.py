"""Make the in-tree package importable when the examples are run from a source checkout."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-plus_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
