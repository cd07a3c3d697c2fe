"""How much of the PCIe link does the copy PATTERN of cz_world_step_host leave unused?  Same bytes (118 MB in, 160 MB
out per frame at 65 536 worlds x 8 bodies, f64), no kernels, no dependencies, both directions at once:
  one   : one copy per direction
  field : one copy per field and direction (9 + 9)
  chunk : one copy per field, chunk and direction (6 x 9 + 6 x 9), the pipeline's pattern"""
import time
import torch
NB = 65536 * 8
IN = [3, 4, 3, 3, 3, 9, 1]      # reals per body: pos ori vel rot acc iitb motion (+ 2 flag bytes)
OUT = [3, 4, 3, 3, 1, 3, 12, 9]  # pos ori vel rot motion lacc transform iitw (+ 1 flag byte)
def bufs(widths, nflags):
    h = [torch.empty(NB * w, dtype=torch.float64).pin_memory() for w in widths] + [torch.empty(NB, dtype=torch.uint8).pin_memory() for _ in range(nflags)]
    d = [torch.empty(NB * w, dtype=torch.float64, device="cuda") for w in widths] + [torch.empty(NB, dtype=torch.uint8, device="cuda") for _ in range(nflags)]
    return h, d, widths + [1] * nflags
hi, di, wi = bufs(IN, 2)
ho, do, wo = bufs(OUT, 1)
bi = sum(t.numel() * t.element_size() for t in hi); bo = sum(t.numel() * t.element_size() for t in ho)
Hi = torch.empty(bi, dtype=torch.uint8).pin_memory(); Di = torch.empty(bi, dtype=torch.uint8, device="cuda")
Ho = torch.empty(bo, dtype=torch.uint8).pin_memory(); Do = torch.empty(bo, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(mode, up=True, down=True, chunks=6):
    if mode == "one":
        if up:
            with torch.cuda.stream(s1): Di.copy_(Hi, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): Ho.copy_(Do, non_blocking=True)
        return
    edges = [0, NB] if mode == "field" else [NB * c // chunks for c in range(chunks + 1)]
    for c in range(len(edges) - 1):
        a, b = edges[c], edges[c + 1]
        if up:
            with torch.cuda.stream(s1):
                for h, d, w in zip(hi, di, wi): d[a * w:b * w].copy_(h[a * w:b * w], non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                for h, d, w in zip(ho, do, wo): h[a * w:b * w].copy_(d[a * w:b * w], non_blocking=True)
print(f"bytes per frame: in {bi/1e6:.1f} MB, out {bo/1e6:.1f} MB")
for mode in ("one", "field", "chunk"):
    for up, down in ((True, False), (False, True), (True, True)):
        run(mode, up, down); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.perf_counter(); R = 10
        for _ in range(R): run(mode, up, down)
        th = time.perf_counter() - t
        torch.cuda.synchronize(); el = (time.perf_counter() - t) / R
        gb = ((bi if up else 0) + (bo if down else 0)) / 1e9
        print(f"{mode:6s} up={int(up)} down={int(down)}: {el*1e3:6.2f} ms per frame, {gb/el:6.1f} GB/s total, host enqueue {th/R*1e3:.2f} ms")
