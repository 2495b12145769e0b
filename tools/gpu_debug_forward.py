"""GPU bring-up check: dsk_rescnn_forward (eval) vs the CPU oracle, plus a first timing.
Usage: python tools/gpu_debug_forward.py [bf16]"""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepspeaker_pytorch_b200 import _lib as L  # noqa: E402
from oracle import rescnn_oracle as O  # noqa: E402


def weights_struct(sd_dev):
    w = L.DskWeights()
    for i, (ck, bp, _, _) in enumerate(O.conv_names()):
        w.conv_w[i] = sd_dev[ck].data_ptr()
        w.bn_gamma[i] = sd_dev[bp + ".weight"].data_ptr()
        w.bn_beta[i] = sd_dev[bp + ".bias"].data_ptr()
        w.bn_running_mean[i] = sd_dev[bp + ".running_mean"].data_ptr()
        w.bn_running_var[i] = sd_dev[bp + ".running_var"].data_ptr()
    w.fc_w = sd_dev["model.fc.weight"].data_ptr()
    w.fc_b = sd_dev["model.fc.bias"].data_ptr()
    w.embedding_size = 512
    return w


def main():
    bf16 = len(sys.argv) > 1 and sys.argv[1] == "bf16"
    lib = L.load()
    h = ctypes.c_void_p()
    L.check(lib.dsk_create(ctypes.byref(h), 0, L.DSK_BF16 if bf16 else L.DSK_F16), "create")
    sd = O.make_state_dict(0)
    sd_dev = {k: v.cuda().contiguous() for k, v in sd.items()}
    w = weights_struct(sd_dev)
    s = L.cur_stream()
    L.check(lib.dsk_load_weights(h, ctypes.byref(w), s), "load_weights")
    ok = True
    for (B, T, scale) in ((4, 160, 1.0), (3, 32, 10.0), (5, 160, 10.0)):
        x = O.make_input(B, T, seed=B, scale=scale)
        with torch.no_grad():
            ref = O.forward(sd, x)
        xd = x.cuda()
        emb = torch.empty(B, 512, device="cuda")
        L.check(lib.dsk_rescnn_forward(h, xd.data_ptr(), B, T, emb.data_ptr(), 0, s), "forward")
        torch.cuda.synchronize()
        got = emb.cpu()
        rel = ((got - ref).norm(dim=1) / ref.norm(dim=1))
        print(f"forward B={B} T={T} scale={scale}: rel-L2 max {rel.max():.3e} mean {rel.mean():.3e}; "
              f"norms {got.norm(dim=1)[:3].tolist()}", flush=True)
        lim = 4e-3 if bf16 else 1e-3
        ok &= bool(rel.max() < lim)
    # timing, batch 64
    B, T = 64, 160
    xd = O.make_input(B, T).cuda()
    emb = torch.empty(B, 512, device="cuda")
    for _ in range(5):
        L.check(lib.dsk_rescnn_forward(h, xd.data_ptr(), B, T, emb.data_ptr(), 0, s))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        L.check(lib.dsk_rescnn_forward(h, xd.data_ptr(), B, T, emb.data_ptr(), 0, s))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"batch-64 forward: {ms:.3f} ms/iter -> {B / ms * 1e3:.0f} emb/s, {B * 2.30667e9 / ms / 1e9:.1f} TFLOP/s")
    print("ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
