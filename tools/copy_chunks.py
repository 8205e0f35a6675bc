"""Host-link ceiling of the host-buffer spectrum path against copy granularity: 32 MiB in + 64 MiB out per step as
plain pinned copies on two streams, split into k pieces each (no kernels, no dependencies between the directions)."""
import time, torch
IN, OUT = 32 << 20, 64 << 20
hin = torch.empty(IN, dtype=torch.uint8).pin_memory(); din = torch.empty(IN, dtype=torch.uint8, device="cuda")
hout = torch.empty(OUT, dtype=torch.uint8).pin_memory(); dout = torch.empty(OUT, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for k in (1, 2, 4, 8, 16, 32, 8, 1):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            for i in range(k):
                a, b = IN // k * i, IN // k * (i + 1)
                with torch.cuda.stream(s1): din[a:b].copy_(hin[a:b], non_blocking=True)
                a, b = OUT // k * i, OUT // k * (i + 1)
                with torch.cuda.stream(s2): hout[a:b].copy_(dout[a:b], non_blocking=True)
            s1.synchronize(); s2.synchronize()
        dt = (time.perf_counter() - t0) / 10
    print("pieces %3d: %.3f ms per step = %.2f Gsamples/s equivalent, %.1f GB/s both ways" % (k, dt * 1e3, (1 << 24) / dt / 1e9, (IN + OUT) / dt / 1e9), flush=True)
# one direction at a time
for name, fn in (("H2D 32 MiB", lambda: din.copy_(hin, non_blocking=True)), ("D2H 64 MiB", lambda: hout.copy_(dout, non_blocking=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("%s alone: %.3f ms" % (name, dt * 1e3), flush=True)
