#!/usr/bin/env python
"""Film output stage timing on one GPU: device-resident C5 film (3840x2160) -> body bytes.
Algorithmic bytes = 12 B read + 3 (gamma8, bmp) or 4 (rgbe) B written per pixel.  L2 is flushed between launches."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ky_b200 as ky

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
dev = ky.Device(0)
film = torch.rand((h, w, 3), device="cuda") * 1.2 - 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
side = torch.cuda.Stream()   # a real stream handle: NULL would mean "the context's own stream, synchronous"
torch.cuda.synchronize()
torch.cuda.set_stream(side)
for name, fmt in (("gamma8", ky.FILM_GAMMA8), ("bmp24", ky.FILM_BMP24), ("rgbe", ky.FILM_RGBE)):
    n = ky.kyd().kyd_film_body_bytes(fmt, w, h)
    body = torch.zeros(n, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    times = []
    for i in range(13):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dev.film_encode_device(film.data_ptr(), w, h, fmt, body.data_ptr(), stream)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(a.elapsed_time(b))
    ms = sorted(times)[len(times) // 2]
    alg = w * h * 12 + n
    print(f"{name:7s} {w}x{h}: {ms * 1e3:8.1f} us  {alg / ms / 1e6:8.1f} GB/s algorithmic ({alg / 1e6:.1f} MB)")
