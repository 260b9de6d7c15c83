import torch, time
for mb in (1, 8, 64, 256):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device='cuda')
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for name, a, b in (('d2h', h, d), ('h2d', d, h)):
        for _ in range(3): a.copy_(b, non_blocking=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): a.copy_(b, non_blocking=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print(f'{name} {mb} MB: {n / dt / 1e9:.1f} GB/s')
