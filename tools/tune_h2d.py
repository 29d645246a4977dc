import torch, time
dev = "cuda:0"
n = 16 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device=dev)
s = [torch.cuda.Stream() for _ in range(4)]
def run(k):
    per = n // k
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        for i in range(k):
            with torch.cuda.stream(s[i]):
                d[i*per:(i+1)*per].copy_(h[i*per:(i+1)*per], non_blocking=True)
        torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / 20
    print(f"{k} concurrent H2D streams: {t*1e6:7.1f} us per 16 MiB  ({n/t/1e9:.1f} GB/s)")
for k in (1, 2, 4, 1, 2): run(k)
