"""Measured HBM stream ceilings of this GPU beside the copy figure of MEASURED_PEAKS.json: pure write (fill), pure read (sum) and
copy over 4 GiB of fp64, best of 10, CUDA events.  The assembly kernels are write-dominated (3.1 GB of matrix values written
against 0.7 GB read at n=119), so the pure-write ceiling is the relevant one.  Writes profiles/hbm_stream_peaks.json."""
import json
import os

import torch

n = 1 << 29                       # 4 GiB of float64
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n, dtype=torch.float64, device="cuda")
a.fill_(1.0); b.fill_(2.0)


def best(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return min(t)


out = {"gpu": torch.cuda.get_device_name(0), "bytes": 8 * n,
       "write_gbs": 8 * n / best(lambda: a.fill_(3.0)) / 1e6,
       "read_gbs": 8 * n / best(lambda: a.sum()) / 1e6,
       "copy_gbs": 16 * n / best(lambda: b.copy_(a)) / 1e6,
       "how": "torch fill_ / sum / copy_ over 4 GiB of float64, best of 10, CUDA events"}
os.makedirs("profiles", exist_ok=True)
json.dump(out, open("profiles/hbm_stream_peaks.json", "w"), indent=1)
print(json.dumps(out))
