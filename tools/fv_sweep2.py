"""FV apply sweep over block width / layers per block (GDTB_FV_BLOCK, GDTB_FV_ROWS); one process per setting because the
knobs are read once.  python tools/fv_sweep2.py"""
import json
import os
import subprocess
import sys

for bx in (64, 128, 256):
    for rows in (0, 8, 16, 32, 64, 128):
        env = dict(os.environ, GDTB_FV_BLOCK=str(bx), GDTB_FV_ROWS=str(rows))
        out = subprocess.run([sys.executable, "tools/fv_ab.py"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
        if out:
            d = json.loads(out[-1])
            print(f"block {bx:4d} rows {rows:4d}: linear {d['linear_events_us']:6.1f} us  burgers {d['burgers_events_us']:6.1f} us  copy {d['copy_us']:5.1f} us", flush=True)
