"""Per-role clock64 trace of CTA 0 of conv3x3_halo_kernel. Usage: python tools/micro/trace_halo.py H W C [flags]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from deepspeaker_pytorch_b200 import _lib as L
lib = L.load()
H, W, C = (int(a) for a in sys.argv[1:4])
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 3
N = 64
h = ctypes.c_void_p(); L.check(lib.dsk_create(ctypes.byref(h), 0, 0))
npos = lib.dsk_padded_positions(N, H, W)
x = torch.zeros(npos, C, dtype=torch.float16, device="cuda")
ZERO = os.environ.get("ZERO") == "1"
if not ZERO: x.normal_()
r = torch.zeros_like(x); o = torch.zeros_like(x)
wp = (torch.zeros(9 * C * C, device="cuda") if ZERO else torch.randn(9 * C * C, device="cuda")).half(); sc = torch.ones(C, device="cuda"); bi = torch.zeros(C, device="cuda")
s = torch.cuda.current_stream().cuda_stream
tr = torch.zeros(3 * 512, dtype=torch.int64, device="cuda")
for it in range(3):
    if it == 2: L.check(lib.dsk_debug_set_trace(h, tr.data_ptr()))
    L.check(lib.dsk_conv3x3_padded(h, x.data_ptr(), wp.data_ptr(), sc.data_ptr(), bi.data_ptr(), r.data_ptr(), o.data_ptr(), N, H, W, C, flags, 20.0, 0, s))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
L.check(lib.dsk_conv3x3_padded(h, x.data_ptr(), wp.data_ptr(), sc.data_ptr(), bi.data_ptr(), r.data_ptr(), o.data_ptr(), N, H, W, C, flags, 20.0, 0, s))
e1.record(); torch.cuda.synchronize()
print('single launch event time us:', e0.elapsed_time(e1) * 1e3)
t = tr.cpu().view(3, 512)
t0 = int(t[t > 0].min())
prod = [int(v) - t0 for v in t[0, :480] if v > 0]
print("CTA0 entry / after-setup / after-pdl_wait / exit (cycles):", [int(v) - t0 for v in t[0, 480:484]])
print("   producer: barriers initialised / weight boxes issued / after its pdl_wait:", [int(v) - t0 for v in t[0, 484:487]],
      "| TMEM alloc begin / end:", [int(v) - t0 for v in t[0, 487:489]])
print("producer A-issue stamps (cycles):", prod[:24])
print("   deltas:", [b - a for a, b in zip(prod, prod[1:])][:24])
names = ["start", "tmem_full", "bar1", "tmem_ld", "res_ok", "math+sts", "fence+bar+store", "done"]
for k in range(12):
    m = [int(v) - t0 if v > 0 else None for v in t[1, 4 * k:4 * k + 3]]
    e = [int(v) - t0 if v > 0 else None for v in t[2, 8 * k:8 * k + 8]]
    if m[0] is None: break
    d = [e[i + 1] - e[i] if (e[i] is not None and e[i + 1] is not None) else None for i in range(7)]
    if e[0] is None or m[2] is None: break
    print(f"tile {k:2d} MMA issue {m[2]-m[1]:5d} | EPI start {e[0]:7d} deltas " + " ".join(f"{names[i+1]}={d[i]}" for i in range(7)))
