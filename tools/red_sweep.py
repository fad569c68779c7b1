#!/usr/bin/env python
"""RED.ADD.U32 rate to uniformly random slots of the pool for several pool sizes (the L2-atomic
limit of step 3, "must be measured"): nk_calibrate(h, 2) on handles of growing pool size.  Run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurokmer_b200 import SpikingKmerCounter  # noqa: E402

for pool in (1, 64, 4096, 65_536, 1_000_000, 2_000_000, 4_000_000, 16_000_000, 32_000_000, 64_000_000, 256_000_000):
    c = SpikingKmerCounter(31, 1.0, 0.95, 2, 1.0, pool, True)
    best = max(c.calibrate(2) for _ in range(2))
    print(json.dumps(dict(pool=pool, acc_mbytes=round(pool * 4 / 1e6, 2), red_gps=round(best / 1e9, 1))), flush=True)
    c.close()
