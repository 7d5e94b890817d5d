"""Raw pinned-memory copy bandwidth of this box (the ceiling of the e2e leg: 16 KB per sample)."""
import json, torch
n = 2_003_328_000 // 8
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64).pin_memory()
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    best = 1e9
    for rep in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(json.dumps(dict(copy=name, bytes=n * 8, ms=round(best, 2), GB_per_s=round(n * 8 / best / 1e6, 1),
                          samples_per_s_at_16KB=round(n * 8 / 16000 / (best * 1e-3)))))
# the shape of one sampling round's copy: 296 chains x 64 samples x 16 KB, row pitch = 423 samples
import ctypes
rt = ctypes.CDLL("libcudart.so")
rows, width, pitch = 296, 64 * 16000, 423 * 16000
assert rows * pitch <= n * 8
st = torch.cuda.current_stream().cuda_stream
for name, call in (("2D (one call)", lambda: rt.cudaMemcpy2DAsync(ctypes.c_void_p(h.data_ptr()), ctypes.c_size_t(pitch), ctypes.c_void_p(d.data_ptr()), ctypes.c_size_t(pitch),
                                                            ctypes.c_size_t(width), ctypes.c_size_t(rows), 2, ctypes.c_void_p(st))),
                   ("296 x 1D", lambda: [rt.cudaMemcpyAsync(ctypes.c_void_p(h.data_ptr() + r * pitch), ctypes.c_void_p(d.data_ptr() + r * pitch), ctypes.c_size_t(width), 2, ctypes.c_void_p(st)) for r in range(rows)])):
    best = 1e9
    for rep in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(json.dumps(dict(copy="D2H " + name, bytes=rows * width, ms=round(best, 2), GB_per_s=round(rows * width / best / 1e6, 1))))
