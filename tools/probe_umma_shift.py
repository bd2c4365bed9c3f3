"""GPU probe for row-shifted / strided-group UMMA A descriptors (see csrc/debug_umma.cu)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib
_lib.init(0)
lib = _lib.load()
lib.eqxv_debug_umma_shift.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
lib.eqxv_debug_umma_shift.restype = C.c_int
rows = 256
g = torch.Generator().manual_seed(0)
a = torch.randn(rows, 64, generator=g).to(torch.bfloat16).cuda()
b = torch.randn(64, 64, generator=g).to(torch.bfloat16).cuda()
for shift, grp in [(0, 8), (1, 8), (3, 8), (8, 8), (9, 8), (0, 10), (11, 10), (21, 10), (2, 12), (5, 16), (0, 9)]:
    for mode in (0, 1):
        if shift + 15 * grp + 8 > rows:
            continue
        out = torch.zeros(128, 64, device="cuda")
        rc = lib.eqxv_debug_umma_shift(a.data_ptr(), rows, b.data_ptr(), out.data_ptr(), shift, grp, mode, None)
        torch.cuda.synchronize()
        idx = torch.tensor([shift + (m // 8) * grp + (m % 8) for m in range(128)], device="cuda")
        ref = a[idx].float() @ b.float().t()
        err = (out - ref).abs().max().item()
        # which rows are right?
        ok_rows = ((out - ref).abs().amax(1) < 1e-2).sum().item()
        print(f"shift={shift:3d} group_rows={grp:3d} base_offset_mode={mode} rc={rc} max_err={err:9.3e} rows_ok={ok_rows}/128", flush=True)
