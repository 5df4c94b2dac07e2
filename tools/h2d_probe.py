import torch, time
dev = torch.device("cuda", 0)
for mb in (64, 512):
    n = mb * (1 << 20)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h2 = torch.empty(n // 4, dtype=torch.uint8, pin_memory=True)
    d2 = torch.ones(n // 4, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        gb = 10 * (n if mode != "d2h" else 0) / 1e9, 10 * ((n // 4) if mode != "h2d" else 0) / 1e9
        print(f"{mb} MB buffers, {mode}: H2D {gb[0] / dt:.1f} GB/s, D2H {gb[1] / dt:.1f} GB/s")
