"""Per-role clock64 trace of CTA 0 of the halo kernel in its 5x5 s2 planar form. Usage: trace_s2.py Hout Wout cin cout"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from deepspeaker_pytorch_b200 import _lib as L
lib = L.load()
Hout, Wout, cin, cout = (int(a) for a in sys.argv[1:5])
N = 64
h = ctypes.c_void_p(); L.check(lib.dsk_create(ctypes.byref(h), 0, 0))
npl = lib.dsk_padded_positions(N, Hout, Wout)
x = torch.zeros(4 * npl, cin, dtype=torch.float16, device="cuda"); x.normal_()
o = torch.zeros(npl, cout, dtype=torch.float16, device="cuda")
w = torch.randn(cout, cin, 5, 5, device="cuda") * 0.02
sc = torch.ones(cout, device="cuda"); bi = torch.zeros(cout, device="cuda")
s = torch.cuda.current_stream().cuda_stream
tr = torch.zeros(3 * 512, dtype=torch.int64, device="cuda")
for it in range(3):
    if it == 2: L.check(lib.dsk_debug_set_trace(h, tr.data_ptr()))
    L.check(lib.dsk_conv5x5s2_planar(h, x.data_ptr(), w.data_ptr(), sc.data_ptr(), bi.data_ptr(), o.data_ptr(), N, Hout, Wout, cin, cout, 2, 20.0, s))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t = tr.cpu().view(3, 512)
t0 = int(t[t > 0].min())
prod = [int(v) - t0 for v in t[0] if v > 0]
print("producer A-issue stamps:", prod[:20])
names = ["start", "tmem_full", "bar1", "tmem_ld", "res_ok", "math+sts", "fence+bar+store", "done"]
for k in range(6):
    m = [int(v) - t0 if v > 0 else None for v in t[1, 4 * k:4 * k + 3]]
    e = [int(v) - t0 if v > 0 else None for v in t[2, 8 * k:8 * k + 8]]
    if m[0] is None: break
    print(f"tile {k}: MMA tmem_empty_ok {m[0]} a_full_ok {m[1]} issued {m[2]} (issue span {m[2]-m[1]}) | EPI tmem_full_ok {e[1]} done {e[7]}")
