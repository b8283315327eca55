import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200 import image_prep
dev = torch.device("cuda")
u8 = torch.randint(0, 256, (2, 375, 1242, 3), dtype=torch.uint8, device=dev)
f32 = torch.randn(2, 3, 384, 1280, device=dev)
out = torch.empty(2, 384, 1280, 4, device=dev)
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 50
print("u8 -> nhwc4: %.1f us   nchw fp32 -> nhwc4: %.1f us" % (t(lambda: image_prep.normalise_pad(u8, nhwc4=out)), t(lambda: image_prep.nchw_to_nhwc4(f32, out))))
