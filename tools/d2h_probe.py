"""PCIe probe: device->pinned-host copy rate of 8 GiB with 1-8 concurrent streams and with chunked
copies in one stream (round 1: 52-56 GB/s in every arrangement: nothing to gain from splitting D2H)."""
import torch, time
n = 8 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
def run(k):
    ss = [torch.cuda.Stream() for _ in range(k)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    part = n // k
    for i, s in enumerate(ss):
        with torch.cuda.stream(s):
            h[i*part:(i+1)*part].copy_(d[i*part:(i+1)*part], non_blocking=True)
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t0) / 1e9
for k in (1, 1, 2, 4, 8, 1):
    print(k, "streams:", round(run(k), 2), "GB/s")
# many small copies in one stream
def chunks(mb):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = mb << 20
    for o in range(0, n, c):
        h[o:o+c].copy_(d[o:o+c], non_blocking=True)
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t0) / 1e9
for mb in (1024, 128, 32, 8):
    print("chunks of", mb, "MB:", round(chunks(mb), 2), "GB/s")
