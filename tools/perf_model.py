"""Shared-memory-port model of conv3x3_halo_kernel tile periods and what it predicts for round 2.

Per 128-position tile every byte that crosses the SM's 128 B/clk shared-memory port is counted as one cycle per 128 B:
  MMA operand fetch   (4 KB of A + n_b bytes of B per M128 x K16 MMA; never faster than the tensor-bound N/2 cycles)
  TMA writes          (halo tile per 64-channel chunk and plane, residual tile)
  epilogue            (per 64 output channels: STS staging 128 + TMA-store read 128 [+ residual LDS 128])
The measured periods (clock64 traces on the B200, tools/micro/trace_halo.py / trace_s2.py) are printed next to the model.
"""
PORT = 128.0


def tile_period(n_tile, cin, taps, W, residual, planes=1, pair=False):
    chunks = cin // 64
    mmas = taps * chunks * 4
    b_bytes = n_tile * 32 * (0.5 if pair else 1.0)
    tensor_total = mmas * n_tile / 2.0                      # cycles if the tensor pipe were the only limit
    mma_port = mmas * (4096 + b_bytes) / PORT               # cycles of the shared-memory port taken by operand fetch
    halo_rows = 128 + 2 * W + 4
    a_writes = chunks * planes * halo_rows * 128 / PORT
    out_chunks = n_tile // 64
    epi = out_chunks * (128 + 128 + (256 if residual else 0))   # STS + TMA read (+ residual TMA write + LDS)
    port_total = mma_port + a_writes + epi
    return max(tensor_total, port_total), mma_port, tensor_total


LAYERS = [  # name, n_tile, cin, taps, W(out), residual, planes, tiles at batch 64, measured period (cycles) or None
    ("stage1 3x3 C=64", 64, 64, 9, 32, False, 1, 1337, 2860),
    ("stage2 3x3 C=128", 128, 128, 9, 16, True, 1, 349, 6183),
    ("conv2 5x5s2 64->128 (planar)", 128, 64, 25, 16, False, 4, 349, 7900),
    ("stage3 3x3 C=256", 128, 256, 9, 8, True, 1, 190, 11300),
    ("conv3 5x5s2 128->256", 128, 128, 25, 8, False, 4, 190, None),
    ("stage4 3x3 C=512", 128, 512, 9, 4, True, 1, 84, 22300),
    ("conv4 5x5s2 256->512", 128, 256, 25, 4, False, 4, 84, None),
]

if __name__ == "__main__":
    print(f"{'layer':34s} {'model':>8s} {'measured':>9s} {'MMA port':>9s} {'tensor':>8s} | {'pair mode':>9s} {'gain':>6s}")
    for name, nt, cin, taps, W, res, planes, tiles, meas in LAYERS:
        t, m, n = tile_period(nt, cin, taps, W, res, planes)
        tp, mp, _ = tile_period(nt, cin, taps, W, res, planes, pair=True)
        print(f"{name:34s} {t:8.0f} {meas if meas else '-':>9} {m:9.0f} {n:8.0f} | {tp:9.0f} {100 * (1 - tp / t):5.1f}%")
    print("\nstage 1 is additionally limited by its epilogue chain (~2700 cycles per tile): pair mode alone does not help it.")
