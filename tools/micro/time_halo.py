"""Time dsk_conv3x3_padded on the four ResCNN 3x3 shapes for several builds of libdsk (epilogue variants)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from deepspeaker_pytorch_b200 import _lib as L

def bind(path):
    lib = ctypes.CDLL(path)
    for name, (res, args) in L.SIGNATURES.items():
        fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
    return lib

libs = {"base": L.LIB_PATH}
for v in sys.argv[1:]:
    libs[v] = os.path.join(ROOT, "tools", "micro", f"libdsk_{v}.so")
N = 64
for name, path in libs.items():
    lib = bind(path)
    h = ctypes.c_void_p()
    assert lib.dsk_create(ctypes.byref(h), 0, 0) == 0
    out = []
    for (H, W, C) in ((80, 32, 64), (40, 16, 128), (20, 8, 256), (10, 4, 512)):
        npos = lib.dsk_padded_positions(N, H, W)
        x = torch.zeros(npos, C, dtype=torch.float16, device="cuda"); x.normal_()
        r = torch.zeros_like(x); o = torch.zeros_like(x)
        wp = torch.randn(9 * C * C, device="cuda").half()
        sc = torch.ones(C, device="cuda"); bi = torch.zeros(C, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        for flags in (2, 3):
            for _ in range(3):
                assert lib.dsk_conv3x3_padded(h, x.data_ptr(), wp.data_ptr(), sc.data_ptr(), bi.data_ptr(), r.data_ptr(), o.data_ptr(), N, H, W, C, flags, 20.0, 0, s) == 0
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                lib.dsk_conv3x3_padded(h, x.data_ptr(), wp.data_ptr(), sc.data_ptr(), bi.data_ptr(), r.data_ptr(), o.data_ptr(), N, H, W, C, flags, 20.0, 0, s)
            e1.record(); torch.cuda.synchronize()
            out.append(f"{e0.elapsed_time(e1) / 20 * 1e3:6.1f}")
    print(f"{name:5s} us per launch [S1 nores,res | S2 | S3 | S4]: " + " ".join(out), flush=True)
    lib.dsk_destroy(h)
