import time, torch, pynvml as n
n.nvmlInit()
h = n.nvmlDeviceGetHandleByIndex(0)
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(50): a @ a
torch.cuda.synchronize()
for name, fn in [("clock", lambda: n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), ("max", lambda: n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)),
                 ("reasons", lambda: n.nvmlDeviceGetCurrentClocksEventReasons(h)), ("power", lambda: n.nvmlDeviceGetPowerUsage(h))]:
    for _ in range(200): a @ a
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < 0.2:
        v = fn(); k += 1
    torch.cuda.synchronize()
    print(name, k, "calls in 0.2 s", v)
