#!/usr/bin/env python
"""Per-phase clock stamps of ONE tile of the column-sweep digit-pass kernel (build with B200_COL_DEBUG=1).
    B200_COL_DEBUG=1 python -m cuda.radixsort_b200.build --force ; python tools/col_timeline.py [--variant 36]
Prints, per warp, the SM-clock offsets of the phase boundaries (see COL_STAMP in csrc/colsweep.cuh)."""
import argparse, ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cuda.radixsort_b200 as rs
from cuda.radixsort_b200 import _lib

NAMES = ["start", "sync0", "tile_in", "regs", "count", "sync1", "scanA", "sync2", "bins", "sync3", "scanB", "sync4",
         "lookback", "turn", "arrive", "scatter", "sync5", "end", "lb_rounds", "lb_walk"]

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", type=int, default=36)
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--reps", type=int, default=1)
    args = ap.parse_args()
    rs.load()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    n = 1 << args.log2n
    keys = rs.generate("uniform", n)
    out = torch.empty_like(keys)
    ws = rs.Workspace("cuda")
    rs.set_param("variant", args.variant)
    for rep in range(args.reps):
      for _ in range(3):
        rs.sort_keys(keys, 8, out=out, workspace=ws)
      torch.cuda.synchronize()
      buf = (ctypes.c_longlong * 320)()
      rc = lib.b200sort_debug_read(buf)
      assert rc == 0, rc
      if rep + 1 < args.reps:
        t0 = min(buf[w * 20] for w in range(16) if buf[w * 20])
        print(json.dumps({"rep": rep, "chain_end": max(buf[w * 20 + 14] for w in range(16)) - t0,
                          "lookback_end": max(buf[w * 20 + 12] for w in range(16) if buf[w * 20]) - t0,
                          "end": max(buf[w * 20 + 17] for w in range(16)) - t0,
                          "lb_rounds": max(buf[w * 20 + 18] for w in range(16)), "lb_walk": max(buf[w * 20 + 19] for w in range(16))}))
    t0 = min(buf[w * 20] for w in range(16) if buf[w * 20])
    for w in range(16):
        if not buf[w * 20]:
            continue
        row = {NAMES[k]: buf[w * 20 + k] - t0 for k in range(18) if buf[w * 20 + k]}
        row.update({NAMES[k]: buf[w * 20 + k] for k in (18, 19)})
        print(json.dumps({"warp": w, **row}))

if __name__ == "__main__":
    main()
