import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib
lib = _lib.load()
out = torch.zeros(4, device="cuda")
for N in (32, 64, 128, 256):
    row = []
    for variant in (0, 1, 2, 3, 4, 6):
        _lib.check(lib.vt_selftest_tc_gemm(None, None, out.data_ptr(), 0, N, 2000, variant, None))
        torch.cuda.synchronize()
        row.append(f"v{variant}:{out[0].item():7.1f}")
    print(f"N={N:3d} ideal {N // 2:3d} cyc/MMA | " + "  ".join(row))
