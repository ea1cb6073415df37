"""Regenerates tests/golden/oracle_checksums.json from the CPU oracle.  The reference has no golden
vectors for this path (parity unpinned); these frozen values only guard the oracle against drift."""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from test_oracle import oracle_checksums
with open(os.path.join(HERE, "oracle_checksums.json"), "w") as f:
    json.dump(oracle_checksums(), f, indent=1)
print("written")
