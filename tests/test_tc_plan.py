"""CPU: emulate the tensor-core program of the contour conv (csrc/tc_contour.cu, built by the host code inside
libbp_b200.so) in NumPy and compare with the direct convolution of the oracle.  This pins the Toeplitz/aligned-chunk
decomposition, the harmonic-stack folding and the edge gating; the GPU test then only has to prove the MMA mechanics."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import model_ref


def _bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def _plan(w):
    from basic_pitch_b200 import _lib

    lib = _lib.load()
    sizes = np.zeros(3, np.int32)
    wc = np.ascontiguousarray(w, np.float32)
    lib.bp_debug_tc_plan(wc.ctypes.data, sizes.ctypes.data, None, None, None, None, None)
    n_tiles, n_steps, n_uses = (int(x) for x in sizes)
    tiles = np.zeros((n_tiles, 2, 2, 128, 8), np.uint16)
    tile_seq = np.zeros(n_steps, np.int32)
    step_use_off = np.zeros(n_steps + 1, np.int32)
    use_words = np.zeros(n_uses, np.uint32)
    group_off = np.zeros(6, np.int32)
    lib.bp_debug_tc_plan(wc.ctypes.data, sizes.ctypes.data, tiles.ctypes.data, tile_seq.ctypes.data,
                         step_use_off.ctypes.data, use_words.ctypes.data, group_off.ctypes.data)
    return tiles, tile_seq, step_use_off, use_words, group_off


def test_tc_program_reproduces_contour_conv(weights_np):
    w = weights_np["contour1_w"]
    tiles, tile_seq, step_use_off, use_words, group_off = _plan(w)
    assert group_off[0] == 0 and group_off[5] == len(tile_seq) and len(tiles) < 400
    # weight tiles as float64 [tile][k 16][n 128] from hi + lo
    tf = _bf16_to_f32(tiles).astype(np.float64)
    t_full = tf[:, 0] + tf[:, 1]  # [tile][kchunk][n][8]
    t_full = t_full.transpose(0, 1, 3, 2).reshape(len(tiles), 16, 128)  # k = kchunk*8 + j
    # split-bf16 representation error of the weights is tiny
    rng = np.random.default_rng(0)
    n_t = 40
    y = rng.standard_normal((n_t, 309))
    ypad = np.zeros((n_t + 2, 320))
    ypad[1:-1, :309] = y  # data rows d = t + 1
    out = np.zeros((n_t, 17 * 16, 8))
    for g in range(5):
        seen = set()
        for s in range(group_off[g], group_off[g + 1]):
            for k in range(step_use_off[s], step_use_off[s + 1]):
                wd = int(use_words[k])
                ftl, q, dt, first = wd & 3, (wd >> 2) & 31, (wd >> 7) & 3, (wd >> 9) & 1
                assert bool(first) == (ftl not in seen)
                seen.add(ftl)
                ft = 4 * g + ftl
                a = ypad[dt : dt + n_t, 16 * q : 16 * q + 16]  # rows m + dt
                d = a @ t_full[tile_seq[s]]  # [n_t][128]
                out[:, 16 * ft : 16 * ft + 16, :] += d.reshape(n_t, 16, 8)
        assert seen == set(range(1 if g == 4 else 4))
    out = out[:, :264]
    h = model_ref.harmonic_stack(torch.from_numpy(y)[None])  # (1,8,T,264)
    ref = F.conv2d(F.pad(h, (19, 19, 1, 1)), torch.from_numpy(w.astype(np.float64)))[0].numpy()  # (8,T,264)
    err = np.abs(out - ref.transpose(1, 2, 0)).max()
    assert err < 2e-4, err  # limited by the 16-bit (hi+lo) weights


def test_tc_program_size_is_bounded(weights_np):
    tiles, tile_seq, step_use_off, use_words, group_off = _plan(weights_np["contour1_w"])
    uses_per_group = [int(step_use_off[group_off[g + 1]] - step_use_off[group_off[g]]) for g in range(5)]
    # <= 5 aligned 16-bin chunks per (frequency tile, channel, time tap)
    assert all(u <= 4 * 8 * 3 * 5 for u in uses_per_group), uses_per_group
    assert (tile_seq >= 0).all() and (tile_seq < len(tiles)).all()
