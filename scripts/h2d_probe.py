"""Probe: pinned host->device bandwidth for the step's input (134 MB of fp32 clip features), 1 copy vs split over 2 / 4 streams."""
import torch, time
n = 32 * 256 * 4096
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
def run(k, reps=20):
    streams = [torch.cuda.Stream() for _ in range(k)]
    chunk = n // k
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print("streams=%d  %.2f ms  %.1f GB/s" % (k, dt * 1e3, n * 4 / dt / 1e9))
for k in (1, 2, 4):
    run(k)
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
