"""Host -> HBM upload of one 800 MB fp64 column: pageable .to(), pinned staging ring, in-place page-locking."""
import time, numpy as np, torch

n = 100_000_000
x = np.random.default_rng(0).random(n)
dev = torch.device("cuda", 0)
torch.cuda.init(); torch.empty(1, device=dev)
print("threads", torch.get_num_threads())

def t_plain():
    torch.cuda.synchronize(); t = time.perf_counter()
    d = torch.from_numpy(x).to(dev); torch.cuda.synchronize()
    return time.perf_counter() - t, d

def t_staged(chunk_mb=32, nbuf=3):
    ce = chunk_mb * (1 << 20) // 8
    torch.cuda.synchronize(); t = time.perf_counter()
    stage = [torch.empty(ce, dtype=torch.float64).pin_memory() for _ in range(nbuf)]
    t_alloc = time.perf_counter() - t
    ev = [None] * nbuf
    src = torch.from_numpy(x)
    d = torch.empty(n, dtype=torch.float64, device=dev)
    s = torch.cuda.current_stream()
    k = 0
    for off in range(0, n, ce):
        m = min(ce, n - off)
        b = k % nbuf
        if ev[b] is not None: ev[b].synchronize()
        stage[b][:m].copy_(src[off:off + m])
        d[off:off + m].copy_(stage[b][:m], non_blocking=True)
        ev[b] = torch.cuda.Event(); ev[b].record(s)
        k += 1
    torch.cuda.synchronize()
    return time.perf_counter() - t, d, t_alloc

def t_register():
    rt = torch.cuda.cudart()
    torch.cuda.synchronize(); t = time.perf_counter()
    r = rt.cudaHostRegister(x.ctypes.data, x.nbytes, 0)
    t_reg = time.perf_counter() - t
    d = torch.empty(n, dtype=torch.float64, device=dev)
    d.copy_(torch.from_numpy(x), non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t
    rt.cudaHostUnregister(x.ctypes.data)
    return dt, d, t_reg, r

for rep in range(3):
    a, d0 = t_plain(); print(f"plain .to(): {a*1e3:.1f} ms  {x.nbytes/a/1e9:.1f} GB/s")
for cm, nb in [(8, 3), (32, 3), (64, 2), (16, 4)]:
    for rep in range(2):
        a, d1, ta = t_staged(cm, nb); print(f"staged {cm} MB x{nb}: {a*1e3:.1f} ms  {x.nbytes/a/1e9:.1f} GB/s (pinned alloc {ta*1e3:.1f} ms) equal={bool(torch.equal(d0, d1))}")
for rep in range(2):
    a, d2, tr, r = t_register(); print(f"hostRegister+copy: {a*1e3:.1f} ms (register {tr*1e3:.1f} ms) rc={r} equal={bool(torch.equal(d0, d2))}")
