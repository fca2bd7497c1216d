import sys, torch
sys.path.insert(0, "/root/repo")
import cuda.radixsort_b200 as rs
n = 1 << 28
for kind in ("sorted", "uniform", "all_equal", "iota"):
    keys = rs.generate(kind, n)
    out = torch.empty_like(keys)
    for _ in range(2): rs.sort_keys(keys, 8, out=out)
    torch.cuda.synchronize()
    rs.profile_enable(True); rs.profile_read()
    for _ in range(5): rs.sort_keys(keys, 8, out=out)
    torch.cuda.synchronize()
    prof = rs.profile_read(); rs.profile_enable(False)
    per = {}
    for t, ms in prof: per.setdefault(t, []).append(ms)
    print(kind, {t: round(sum(v)/len(v), 3) for t, v in sorted(per.items())})
