"""Does H2D overlap D2H on this box?  Pinned buffers, two streams, CUDA events."""
import torch

dev = torch.device("cuda", 0)
n_in, n_out = 2073600 * 24, 2073600 * 32
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        fn()
        a, b = torch.cuda.Event(), torch.cuda.Event()
        a.record(s1)
        b.record(s2)
        torch.cuda.current_stream().wait_event(a)
        torch.cuda.current_stream().wait_event(b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


def chunks(k):
    def f():
        ci, co = n_in // k, n_out // k
        for i in range(k):
            with torch.cuda.stream(s1):
                d_in[i * ci:(i + 1) * ci].copy_(h_in[i * ci:(i + 1) * ci], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[i * co:(i + 1) * co].copy_(d_out[i * co:(i + 1) * co], non_blocking=True)
    return f


print("h2d alone  %.3f ms" % timed(h2d))
print("d2h alone  %.3f ms" % timed(d2h))
print("both       %.3f ms" % timed(both))
for k in (4, 8, 16, 32):
    print("both in %2d chunks each  %.3f ms" % (k, timed(chunks(k))))
