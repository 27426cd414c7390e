import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import Net_Restormer as N
torch.manual_seed(0)
T = N.T_net(decoder=True).cuda()
x = torch.rand(8, 3, 256, 256, device="cuda")
with torch.no_grad():
    for _ in range(2): y = T(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): y = T(x)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"c5: T_net forward 256x256 B=8: {ms:.1f} ms -> {8/ms*1000:.1f} img/s; out mean {y.mean().item():.4f}; mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
# arbitrary H x W (multiples of 8), as tester.py feeds whole images
x2 = torch.rand(1, 3, 200, 312, device="cuda")
with torch.no_grad():
    y2 = T(x2)
print("non-square ok", tuple(y2.shape), torch.isfinite(y2).all().item())
