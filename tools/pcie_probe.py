import torch, time
torch.cuda.init()
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t = time.time()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.time() - t
    return reps * n * (h2d + d2h) / dt / 1e9
for _ in range(2):
    print("H2D only %.1f GB/s  D2H only %.1f GB/s  both %.1f GB/s (sum)" % (run(1, 0), run(0, 1), run(1, 1)))
