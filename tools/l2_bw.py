import torch, time
for mb in (16, 32, 64, 96, 4096):
    n = mb * 1024 * 1024 // 4
    x = torch.randn(n, device="cuda")
    for _ in range(3): x.sum()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50 if mb < 1000 else 5
    e0.record()
    for _ in range(reps): x.sum()
    e1.record(); torch.cuda.synchronize()
    print(f"sum over {mb} MB: {mb/1024*reps/(e0.elapsed_time(e1)/1e3):.0f} GiB/s")
