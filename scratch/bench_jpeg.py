import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch, cv2
import cic_b200 as cic
from test_oracle_extras import _jpeg_test_image
for (b, h, w) in [(64, 512, 512), (8, 1080, 1920), (4, 2160, 3840)]:
    base = np.stack([_jpeg_test_image(h, w, "smooth", i) for i in range(min(b, 4))])
    x = torch.from_numpy(np.concatenate([base] * (b // len(base)))).cuda()
    cap = 1024 + 2 * h * w
    out, sizes = cic.ops.jpeg_encode_device(x, capacity=cap)
    torch.cuda.synchronize()
    ok = cic.ops.jpeg_encode(x[:1])[0] == bytes(cv2.imencode(".jpg", base[0])[1])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ws = torch.empty(int(cic._lib.lib.cic_jpeg_workspace_bytes(b, h, w)), dtype=torch.uint8, device="cuda")
    from cic_b200.runtime import ptr
    def run():
        cic._lib.check(cic._lib.lib.cic_jpeg_encode_u8(ptr(x), b, h, w, 0, 95, ptr(out), cap, ptr(sizes), ptr(ws), ws.numel(), cic.runtime.stream_ptr()))
    for _ in range(3): run()
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    t = time.time(); [cv2.imencode(".jpg", base[0]) for _ in range(5)]; cpu_ms = (time.time() - t) / 5 * 1e3
    print(f"{b}x{h}x{w}: {ms:.3f} ms  {b*h*w/ms/1e3:.0f} MPix/s  bytes/img {int(sizes[0])}  identical {ok}  cv2 one image {cpu_ms:.2f} ms ({h*w/cpu_ms/1e3:.0f} MPix/s/core)  ws {ws.numel()/1e6:.0f} MB")
