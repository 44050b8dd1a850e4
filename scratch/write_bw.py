"""HBM write-only and read-only rates next to the copy rate (the roofline denominators of write-dominated kernels)."""
import json, torch
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda")      # 4 GB
b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms_fill = t(lambda: a.zero_())
ms_copy = t(lambda: b.copy_(a))
ms_sum = t(lambda: a.sum())
print(json.dumps({"bytes": 4 * n, "write_only_gbs": 4 * n / ms_fill / 1e6, "copy_gbs_read_plus_write": 8 * n / ms_copy / 1e6, "read_only_gbs": 4 * n / ms_sum / 1e6}))
