import os, subprocess, time, torch
print(subprocess.run(r"nvidia-smi topo -m; lscpu | grep -i 'numa\|^CPU(s)\|Model name'; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; nproc; cat /proc/self/status | grep -i 'Cpus_allowed_list\|Mems_allowed_list'", shell=True, capture_output=True, text=True).stdout)
dev = torch.device("cuda:0")
def bw(tag):
    x = torch.empty(64 * 1024 * 1024 // 4).pin_memory()
    x.fill_(1.0)
    y = torch.empty_like(x, device=dev)
    for _ in range(3): y.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): y.copy_(x, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    h2d = 10 * x.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    e0.record()
    for _ in range(10): x.copy_(y, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    d2h = 10 * x.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(f"{tag}: H2D {h2d:.1f} GB/s  D2H {d2h:.1f} GB/s", flush=True)
bw("default affinity")
cpus = sorted(os.sched_getaffinity(0))
print("cpus", cpus)
for c in (cpus[0], cpus[len(cpus)//2], cpus[-1]):
    os.sched_setaffinity(0, {c})
    bw(f"pinned from cpu {c}")
