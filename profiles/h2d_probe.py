import time, torch, numpy as np
x = [torch.from_numpy(np.random.rand(144, 200, 128)).pin_memory() for _ in range(8)]
s = torch.cuda.Stream()
for name, ctxm in (('default', torch.cuda.stream(torch.cuda.current_stream())), ('side', torch.cuda.stream(s))):
    with ctxm:
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ys = [t.to('cuda', non_blocking=True) for t in x]
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(name, 'issue %.2f ms, complete %.2f ms, %.1f GB/s' % (1e3 * (t1 - t0), 1e3 * (t2 - t0), 8 * x[0].numel() * 8 / (t2 - t0) / 1e9))
            del ys
