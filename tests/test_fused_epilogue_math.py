"""CPU: the decomposition behind the fused epilogues of the tensor-core kernels (csrc/tc_conv.cu: contour_reduce_store,
reduce_store) and the tap-sum kernel (csrc/cnn.cu: halo_tapsum_kernel), restated in NumPy with the same index
arithmetic and compared with the direct convolution of the oracle.

Per frequency tile of FLT bins the epilogue reduces relu(conv1) over channels and frequency taps into KH time-tap planes
of J = FLT + 2*HALO output offsets,   Q[ft][dt][j][t] = sum_{c,df} x[c][t][FLT*ft + j + df - 2*HALO] * w2[c][dt][df]
(output bin f = FLT*ft + j - HALO; the HALO outer columns on either side belong to the neighbouring tiles), and the
tap-sum kernel computes   out[t][f] = sigmoid(b + sum_dt (Q[ft][dt][r + HALO][t + dt - PT] + neighbour halo term)).
This pins the halo bookkeeping, the time-tap stride and the zero padding; the GPU tests then only have to prove the
kernels."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

CASES = {  # name: (weights key, C_in, KH, KW, FLT, HALO, W)
    "contour2": ("contour2_w", 8, 5, 5, 16, 2, 264),
    "note2": ("note2_w", 32, 7, 3, 4, 1, 88),
    "onset2": ("onset2_w", 32, 3, 3, 4, 1, 88),  # channels 1..32 of the 33-channel conv; channel 0 is the note input
}


def _epilogue_planes(x, w2, KH, KW, FLT, HALO, W):
    """x: relu(conv1) [C][T][W]; w2 [C][KH][KW] -> Q [tiles][KH][J][T] exactly as the epilogue accumulates it."""
    C, T, _ = x.shape
    n_tiles = (W + FLT - 1) // FLT
    J = FLT + 2 * HALO
    Q = np.zeros((n_tiles, KH, J, T))
    for ft in range(n_tiles):
        for fl in range(FLT):
            g = FLT * ft + fl
            if g >= W:  # bins past the image contribute zero (masked in the epilogue's first pass)
                continue
            for df in range(KW):
                j = fl - df + 2 * HALO  # input bin fl feeds output offset j (KW = 2*HALO + 1)
                for dt in range(KH):
                    Q[ft, dt, j, :] += np.einsum("ct,c->t", x[:, :, g], w2[:, dt, df])
    return Q


def _tapsum(Q, bias, KH, PT, FLT, HALO, W, extra=None):
    n_tiles, _, J, T = Q.shape
    out = np.full((T, W), float(bias))
    for f in range(W):
        ft, r = divmod(f, FLT)
        terms = [(ft, r + HALO)]
        if r < HALO and ft - 1 >= 0:
            terms.append((ft - 1, r + HALO + FLT))
        if r >= FLT - HALO and ft + 1 < n_tiles:
            terms.append((ft + 1, r + HALO - FLT))
        for t in range(T):
            for dt in range(KH):
                tt = t + dt - PT
                if 0 <= tt < T:
                    out[t, f] += sum(Q[a, dt, j, tt] for a, j in terms)
    if extra is not None:
        out += extra
    return 1.0 / (1.0 + np.exp(-out))


@pytest.mark.parametrize("name", list(CASES))
def test_fused_second_conv_decomposition(weights_np, name):
    key, C, KH, KW, FLT, HALO, W = CASES[name]
    assert KW == 2 * HALO + 1
    rng = np.random.default_rng(len(name))
    T = 23
    x = np.maximum(rng.standard_normal((C, T, W)), 0.0)  # relu(conv1)
    w_full = weights_np[key].astype(np.float64)  # [1][C(+1)][KH][KW]
    bias = float(weights_np[key[:-2] + "_b"].reshape(-1)[0])
    PT = KH // 2
    if name == "onset2":
        note = rng.random((T, W))
        w2 = w_full[0, 1:]
        wx = w_full[0, 0]
        extra = F.conv2d(torch.from_numpy(note)[None, None], torch.from_numpy(wx)[None, None], padding=(1, 1))[0, 0].numpy()
        full_in = np.concatenate([note[None], x])
    else:
        w2 = w_full[0]
        extra = None
        full_in = x
    Q = _epilogue_planes(x, w2, KH, KW, FLT, HALO, W)
    got = _tapsum(Q, bias, KH, PT, FLT, HALO, W, extra)
    ref = torch.sigmoid(F.conv2d(torch.from_numpy(full_in)[None], torch.from_numpy(w_full), torch.tensor([bias], dtype=torch.float64),
                                 padding=(PT, HALO)))[0, 0].numpy()
    assert got.shape == ref.shape == (T, W)
    assert np.abs(got - ref).max() < 1e-12


def test_packed_tap_pairs_of_the_contour_epilogue(weights_np):
    """The contour epilogue pairs output offsets for the packed FMAs: an input bin at even offset bl feeds the pairs
    (bl,bl+1), (bl+2,bl+3), (bl+4,bl+5) with (w4,w3), (w2,w1), (w0,0); at odd bl the pairs (bl-1,bl), (bl+1,bl+2), (bl+3,bl+4)
    with (0,w4), (w3,w2), (w1,w0)  [output offset j = bl + 4 - df] — same sums as the plain five taps."""
    w = weights_np["contour2_w"].astype(np.float64)[0, 3, 2]  # one (channel, dt) row of 5 frequency taps
    for bl in range(16):
        plain = np.zeros(22)
        for df in range(5):
            plain[bl + 4 - df] += w[df]
        packed = np.zeros(22)
        j2 = bl >> 1
        pairs = [(w[4], w[3]), (w[2], w[1]), (w[0], 0.0)] if bl % 2 == 0 else [(0.0, w[4]), (w[3], w[2]), (w[1], w[0])]
        for k, (a, b) in enumerate(pairs):
            packed[2 * (j2 + k)] += a
            packed[2 * (j2 + k) + 1] += b
        np.testing.assert_array_equal(plain, packed)
