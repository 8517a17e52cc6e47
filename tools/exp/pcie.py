import torch, time
for mb in (16, 100, 168):
    h = torch.empty(mb * 1024 * 1024 // 4).pin_memory(); d = torch.empty_like(h, device="cuda")
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); [fn() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
        print(name, mb, "MB: %.1f GB/s" % (5 * mb / 1024 / (e0.elapsed_time(e1) / 1e3)))
