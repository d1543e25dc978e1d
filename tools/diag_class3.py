"""Class-3 families whose sub-ranges can follow the first parent's states without an extra pass
(range_plan, csrc/common.cuh): count-kernel time with the top split and with the generic cut
(BIC_TOPSPLIT=0).  Run on a B200:  python tools/diag_class3.py"""
import os, sys
import numpy as np
sys.path.insert(0, '.')
import torch
import dags_vae_search_b200 as pkg

N = 8_000_000
gen = torch.Generator(device="cuda"); gen.manual_seed(5)
card = np.array([21, 20, 19, 7, 6, 5, 21, 20, 19, 7, 6, 5, 21, 20, 19, 7, 6, 5], dtype=np.int32)
codes = torch.stack([torch.randint(0, int(c), (N,), device="cuda", dtype=torch.uint8, generator=gen) for c in card])
fams = []
for b in (0, 6, 12):      # (child, parents): 21*20*19*7 = 55 860 cells (2 passes), 21*20*6*5*7 = 88 200 (2), 21*19*7*6*5 = 83 790 (2)
    fams += [(b + 3, [b, b + 1, b + 2]), (b + 3, [b, b + 1, b + 4, b + 5]), (b + 5, [b, b + 2, b + 3, b + 4])]
for top in ("1", "0", "1", "0"):
    os.environ["BIC_TOPSPLIT"] = top
    with pkg.BicScorer(codes, card, device=0) as s:
        s.score_families([f[0] for f in fams], [f[1] for f in fams], no_cache=True)
        s.profile_enable(True); s.profile_reset()
        for _ in range(5):
            out = s.score_families([f[0] for f in fams], [f[1] for f in fams], no_cache=True)
        p = s.profile()
        print("BIC_TOPSPLIT=%s  class-3 launches %d  ms per launch %.4f  checksum %.6f" % (
            top, p["class_launches"][3], p["class_ms"][3] / max(1, p["class_launches"][3]), float(out.sum())))
