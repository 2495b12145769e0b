"""Log-fbank front-end (audio_processing.py:9-36): the numpy oracle's invariants on the CPU, and the GPU kernel against the
oracle.  python_speech_features itself is absent (not vendored in the reference): parity against the package is unpinned."""
import ctypes
import math

import numpy as np
import pytest
import torch

from deepspeaker_pytorch_b200 import _lib as L
from oracle import fbank_oracle as FO


def synth(n, seed, sr=16000):
    g = np.random.RandomState(seed)
    t = np.arange(n) / sr
    x = 0.3 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 2300 * t + 1.0) + 0.05 * g.randn(n)
    x *= np.linspace(0.2, 1.0, n)                      # non-stationary envelope
    return x.astype(np.float32)                        # librosa.load yields float32


def test_frame_count_matches_the_published_formula():
    lib = L.load()                                     # host-only entry point: no GPU needed
    for sr in (8000, 16000):
        flen, step = FO.round_half_up(0.025 * sr), FO.round_half_up(0.01 * sr)
        for n in (1, flen - 1, flen, flen + 1, flen + step, 3 * sr + 7, 16000, 48000, 100003):
            want = 1 if n <= flen else 1 + int(math.ceil((n - flen) / step))
            assert lib.dsk_fbank_num_frames(n, sr) == want, (sr, n)
            assert FO.fbank(np.zeros(n, np.float32) + 1e-3, samplerate=sr, nfilt=64)[0].shape == (want, 64)


def test_oracle_filterbank_and_tone_response():
    fb = FO.get_filterbanks(64, 512, 16000, 0, 8000)
    assert fb.shape == (64, 257) and fb.min() >= 0.0 and fb.max() <= 1.0
    peaks = fb.argmax(axis=1)
    assert np.all(np.diff(peaks) >= 0) and peaks[-1] < 257                      # triangles ordered along frequency
    sr, f0 = 16000, 1000.0
    x = np.sin(2 * np.pi * f0 * np.arange(sr) / sr).astype(np.float32)
    feat, energy = FO.fbank(x, samplerate=sr, nfilt=64)
    k = int(round(f0 / sr * 512))
    assert np.all(feat[5:-5].argmax(axis=1) == fb[:, k].argmax())               # the tone lands in its mel filter
    m = FO.mk_mfb(x)
    assert np.allclose(m.mean(axis=0), 0.0, atol=1e-9) and m.shape == feat.shape


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", [(48000, 1), (16000, 2), (399, 3), (400, 4), (401, 5), (560, 6), (100003, 7)])
def test_gpu_fbank_matches_the_oracle(cuda_dev, n, seed):
    from deepspeaker_pytorch_b200 import frontend

    x = synth(n, seed)
    ref_lin, _ = FO.fbank(x, samplerate=16000, nfilt=64)
    got_lin = frontend.mk_mfb(torch.from_numpy(x).cuda(), use_logscale=False, subtract_mean=False).cpu().numpy()
    assert got_lin.shape == ref_lin.shape
    # fp32 FFT on the device against float64 numpy: relative to each frame's largest filter output
    scale = np.maximum(ref_lin.max(axis=1, keepdims=True), 1e-12)
    assert np.abs(got_lin - ref_lin).max() / 1.0 <= 1e-4 * scale.max() + 1e-12
    assert (np.abs(got_lin - ref_lin) / scale).max() < 2e-5
    ref = FO.mk_mfb(x)
    got = frontend.mk_mfb(torch.from_numpy(x).cuda()).cpu().numpy()
    loud = ref_lin > 1e-3 * scale                       # bins above the fp32 noise floor of their frame: 0.01 dB
    assert np.abs(got - ref)[loud].max() < 1e-2
    assert np.abs(got - ref).max() < 0.5                 # everywhere else (round-off-dominated filters): half a dB
    assert abs(got.mean(axis=0)).max() < 1e-3            # per-bin mean removed
    with pytest.raises(RuntimeError):
        frontend.mk_mfb(torch.from_numpy(x))


@pytest.mark.gpu
def test_gpu_fbank_feeds_the_network_layout(cuda_dev):
    """(frames, 64) is the (T, 64) layout the model's (B, 1, T, 64) input is cropped from (constants.py: 32-frame crops)."""
    import deepspeaker_pytorch_b200 as dsk
    from deepspeaker_pytorch_b200 import frontend
    from oracle import rescnn_oracle as O

    x = synth(3 * 16000, 11)
    feat = frontend.mk_mfb(torch.from_numpy(x).cuda())
    crops = torch.stack([feat[j - 9:j + 23] for j in (9, 100, 200, feat.shape[0] - 23)]).unsqueeze(1).contiguous()   # NUM_PREVIOUS_FRAME / NUM_NEXT_FRAME
    assert crops.shape == (4, 1, 32, 64)
    sd = O.make_state_dict(0, 16)
    m = dsk.DeepSpeakerModel(512, 16).cuda().eval()
    m.load_state_dict(sd)
    with torch.no_grad():
        e = m(crops)
        ref = O.forward(sd, torch.from_numpy(np.stack([FO.mk_mfb(x)[j - 9:j + 23] for j in (9, 100, 200, feat.shape[0] - 23)])).float().unsqueeze(1))
    assert ((e.cpu() - ref).norm(dim=1) / ref.norm(dim=1)).max().item() < 1e-3      # audio -> embedding, GPU vs oracle end to end
