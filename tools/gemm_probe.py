"""Probe: strict-fp32 GEMM / conv rates on this GPU (cuBLAS sgemm vs cuDNN conv) at the backbone's shapes."""
import torch, time
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = "cuda"
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it * 1e-3
print("== sgemm W[Cout,K] x X[K,N]")
for (co, k, n) in [(64, 27, 491520), (64, 576, 122880), (128, 576, 122880), (128, 1152, 30720), (256, 1152, 30720), (256, 2304, 7680), (512, 2304, 7680), (512, 4608, 1920),
                   (64, 99, 32768), (128, 64, 32768), (32, 3, 131072), (512, 1536, 512), (128, 256, 32768)]:
    w = torch.randn(co, k, device=dev); x = torch.randn(k, n, device=dev)
    dt = t(lambda: torch.matmul(w, x))
    print(f"Cout={co:4d} K={k:5d} N={n:7d}: {dt*1e6:8.1f} us  {2*co*k*n/dt/1e12:6.1f} TFLOP/s")
print("== cudnn conv3x3 B=2")
for (ci, co, h, w_, s) in [(3, 64, 384, 1280, 1), (64, 64, 384, 1280, 2), (64, 128, 192, 640, 1), (128, 128, 192, 640, 2), (128, 256, 96, 320, 1), (256, 256, 96, 320, 2), (256, 512, 48, 160, 1), (512, 512, 48, 160, 2)]:
    x = torch.randn(2, ci, h, w_, device=dev); conv = torch.nn.Conv2d(ci, co, 3, s, 1, bias=False).to(dev)
    with torch.no_grad():
        dt = t(lambda: conv(x))
        ho, wo = h // s, w_ // s
        fl = 2 * 2 * ho * wo * ci * 9 * co
        xc = x.to(memory_format=torch.channels_last); convc = conv.to(memory_format=torch.channels_last)
        dtc = t(lambda: convc(xc))
        dtu = t(lambda: torch.nn.functional.unfold(x, 3, padding=1, stride=s))
    print(f"conv {ci}->{co} {h}x{w_} s{s}: nchw {dt*1e6:8.1f} us {fl/dt/1e12:5.1f} TF | nhwc {dtc*1e6:8.1f} us | unfold {dtu*1e6:8.1f} us")
print("== tf32 allowed")
torch.backends.cudnn.allow_tf32 = True
for (ci, co, h, w_, s) in [(64, 128, 192, 640, 1), (256, 512, 48, 160, 1)]:
    x = torch.randn(2, ci, h, w_, device=dev); conv = torch.nn.Conv2d(ci, co, 3, s, 1, bias=False).to(dev)
    with torch.no_grad():
        dt = t(lambda: conv(x))
    print(f"conv tf32 {ci}->{co}: {dt*1e6:8.1f} us")
