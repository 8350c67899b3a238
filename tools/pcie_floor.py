"""PCIe floor of the e2e number: pinned H2D of the GAF bytes and D2H of the PAF bytes of the bench workload,
alone and concurrently on two streams (what g2p_convert_host can at best overlap)."""
import time, torch
IN, OUT = 1_373_582_239, 3_057_789_728
h_in = torch.empty(IN, dtype=torch.uint8).pin_memory(); h_out = torch.empty(OUT, dtype=torch.uint8).pin_memory()
d_in = torch.empty(IN, dtype=torch.uint8, device="cuda"); d_out = torch.empty(OUT, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, n=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D {a:.1f} ms ({IN/a/1e6:.1f} GB/s)  D2H {b:.1f} ms ({OUT/b/1e6:.1f} GB/s)  both {c:.1f} ms")
