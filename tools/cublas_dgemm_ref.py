import torch, time
a = torch.randn(512, 512, dtype=torch.float64, device="cuda"); b = torch.randn(512, 4096, dtype=torch.float64, device="cuda")
for _ in range(5): c = a @ b
torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(50): c = a @ b
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 50
print("cuBLAS dgemm 512x512x4096: %.4f ms  %.1f TFLOP/s" % (ms, 2 * 512 * 512 * 4096 / ms / 1e9))
a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda"); b = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
for _ in range(3): c = a @ b
torch.cuda.synchronize(); s.record()
for _ in range(10): c = a @ b
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
print("cuBLAS dgemm 4096^3: %.4f ms  %.1f TFLOP/s" % (ms, 2 * 4096**3 / ms / 1e9))
