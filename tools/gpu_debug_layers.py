"""Per-layer comparison of the train-mode forward (saved raw / y tensors) against the oracle (fp32 and storage-matched)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepspeaker_pytorch_b200 as dsk
from deepspeaker_pytorch_b200 import _lib as L
from oracle import rescnn_oracle as O
import torch.nn.functional as F

B, T = 4, 160
sd = O.make_state_dict(0, 16)
m = dsk.DeepSpeakerModel(512, 16).cuda().train()
m.load_state_dict(sd)
x = O.make_input(B, T, 10, 3.0)
out = m(x.cuda())
torch.cuda.synchronize()
eng = m._engine
tctx = out.grad_fn.guard.tctx
taps32, tapsq = {}, {}
e32 = O.forward(sd, x, True, {}, taps32)
eq = O.forward(sd, x, True, {}, tapsq, storage=torch.float16)
print("emb: gpu-vs-fp32 %.3e  gpu-vs-q %.3e  q-vs-fp32 %.3e" % tuple(
    ((a - b).norm(dim=1) / b.norm(dim=1)).max().item() for a, b in ((out.detach().cpu(), e32), (out.detach().cpu(), eq), (eq, e32))))
for i in range(12):
    H, W, C = T >> (i // 3 + 1), 64 >> (i // 3 + 1), 64 << (i // 3)
    buf = torch.empty(B, C, H, W, device="cuda")
    L.check(eng.lib.dsk_train_ctx_read(eng.handle, tctx, 1, i, buf.data_ptr(), L.cur_stream()))
    torch.cuda.synchronize()
    y = buf.cpu()
    r32 = ((y - taps32[i]).norm() / taps32[i].norm()).item()
    rq = ((y - tapsq[i].detach()).norm() / tapsq[i].norm()).item()
    nflip32 = ((y > 0) != (taps32[i] > 0)).float().mean().item()
    nflipq = ((y > 0) != (tapsq[i] > 0)).float().mean().item()
    print(f"layer {i:2d} y: rel vs fp32 {r32:.3e} vs storage-matched {rq:.3e}; zero-mask mismatch {nflip32:.2e} / {nflipq:.2e}; max|y| {y.abs().max():.2f}")
