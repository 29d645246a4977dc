"""H2D copy time of a pinned buffer vs the state of the CPU caches (profiles/r1_h2d_cpu_cache_effect.txt)."""
import torch, time, numpy as np
dev = "cuda:0"
n = 1 << 20
torch.cuda.init()
def h2d_us(buf, dst, reps=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): dst.copy_(buf, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
src = torch.randint(0, 65536, (n, 4), dtype=torch.int32).to(torch.uint16)
hs = [src.clone().pin_memory() for _ in range(2)]
dst = torch.empty((n, 4), dtype=torch.uint16, device=dev)
warm = torch.empty(1 << 20, dtype=torch.uint8, pin_memory=True); h2d_us(warm, torch.empty(1 << 20, dtype=torch.uint8, device=dev))
print("fresh pinned buffers, 8 MiB each (us per H2D):")
for r in range(6):
    print("  ", " ".join(f"buf{i}:{h2d_us(hs[i], dst):6.0f}" for i in range(2)), flush=True)
big = np.ones(1 << 28, dtype=np.uint8); big += 1; s = int(big[::4096].sum())
print("after evicting CPU caches with a 256 MiB array:")
for r in range(3):
    print("  ", " ".join(f"buf{i}:{h2d_us(hs[i], dst):6.0f}" for i in range(2)), flush=True)
for trial in range(3):
    hs[1].copy_(src)            # CPU rewrites buffer 1 (the producer writing a new step's sticks)
    print("after the CPU rewrote buf1:", " ".join(f"buf{i}:{h2d_us(hs[i], dst):6.0f}" for i in range(2)),
          " again:", " ".join(f"buf{i}:{h2d_us(hs[i], dst):6.0f}" for i in range(2)), flush=True)
# write-combined host memory via cudaHostAlloc
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
p = ctypes.c_void_p()
assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n * 8), ctypes.c_uint(4)) == 0   # cudaHostAllocWriteCombined
arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint16)), shape=(n, 4))
wc = torch.from_numpy(arr)
for trial in range(3):
    t0 = time.perf_counter(); wc.copy_(src); t1 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), p, ctypes.c_size_t(n * 8), 1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)); e1.record(); torch.cuda.synchronize()
    print(f"write-combined buffer: CPU fill {1e6*(t1-t0):.0f} us, H2D {e0.elapsed_time(e1)*1e3:.0f} us", flush=True)
t0 = time.perf_counter(); hs[0].copy_(src); t1 = time.perf_counter(); print(f"cached pinned buffer: CPU fill {1e6*(t1-t0):.0f} us")
