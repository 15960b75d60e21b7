import os, time, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "torch threads", torch.get_num_threads(), "interop", torch.get_num_interop_threads())
try:
    print("cgroup cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("no cgroup v2 cpu.max", e)
def bench(nt):
    torch.set_num_threads(nt)
    x = torch.randn(500, 21, 21, dtype=torch.float64)
    for name, fn in (("full", lambda: torch.full((500, 21, 21), -1e30, dtype=torch.float32)), ("max", lambda: torch.max(x, dim=1)[0]),
                     ("zeros_like", lambda: torch.zeros_like(x)), ("log", lambda: torch.log(x.abs() + 1))):
        fn()
        t0 = time.perf_counter()
        for _ in range(200): fn()
        print(f"threads {nt:3d} {name:10s} {(time.perf_counter() - t0) / 200 * 1e6:8.1f} us")
for nt in (torch.get_num_threads(), 16, 8, 4, 1):
    bench(nt)
