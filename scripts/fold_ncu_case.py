import sys, os, torch
sys.path.insert(0, os.getcwd())
import pulsarbat_b200 as pb
dev = torch.device("cuda:0")
for e in (1, 4):
    x = pb.DeviceArray(torch.rand((2 ** 25, e), device=dev, dtype=torch.float32))
    for _ in range(2):
        pb.kernels.fold(x, [0.123, 29.7, 1e-6], 16e6, 1024)
torch.cuda.synchronize()
