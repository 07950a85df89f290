"""CPU: the column-parallel schedule of the sequential decode loops (csrc/decode.cu: decode_seq_kernel) restated in
NumPy and compared with the sequential restatement of the reference (oracle/decode_ref.py, reference:
basic_pitch/note_creation.py:409-509).

Both greedy loops of the reference read only the pitch column f of a step and zero columns f-1 .. f+1, so two steps in
columns more than one apart commute.  The kernel therefore lets every column run ahead of its neighbours:
  onset loop    a column processes its next candidate (t0, f) as long as it precedes (t descending, then f descending)
                the next unprocessed candidates of columns f-1 and f+1,
  melodia loop  a column picks its maximum (v, t) as long as it precedes (v descending, t ascending, f ascending) the
                maxima its neighbours had at the start of the round,
rounds separated by barriers; notes are emitted in any order and sorted by the sequential order key at the end.  This
test pins that the schedule reproduces the sequential result exactly (start, end, pitch, and the order of the list)."""
import numpy as np
import pytest

from oracle import decode_ref


def _scan(E, f, i, step, limit_ok, tol, thresh):
    quiet = 0
    while limit_ok(i) and quiet < tol:
        quiet = quiet + 1 if E[i, f] < thresh else 0
        i += step
    return i, quiet


def parallel_decode(frames, cand, frame_thresh, min_note_len, energy_tol=11, melodia=True):
    """frames (T,88) float32; cand (T,88) bool onset candidates -> [(t0, t1, pitch_idx)] in the reference's order."""
    n_t, n_f = frames.shape
    E = np.array(frames, dtype=np.float64)

    def wipe(t0, t1, f):
        E[t0:t1, f] = 0
        if f < n_f - 1:
            E[t0:t1, f + 1] = 0
        if f > 0:
            E[t0:t1, f - 1] = 0

    # ---------------- onset loop ----------------
    lists = [list(np.nonzero(cand[:, f])[0][::-1]) for f in range(n_f)]  # candidate times per column, descending
    pos = [0] * n_f
    onset_notes = []
    rounds = 0
    while True:
        front = [lists[f][pos[f]] if pos[f] < len(lists[f]) else -1 for f in range(n_f)]
        if all(t < 0 for t in front):
            break
        rounds += 1

        def precedes(t, f, g):  # (t, f) before column g's front?
            if g < 0 or g >= n_f or front[g] < 0:
                return True
            return t > front[g] or (t == front[g] and f > g)

        fired = False
        for f in range(n_f):  # (concurrently in the kernel: firing columns are never adjacent)
            if front[f] < 0 or not (precedes(front[f], f, f - 1) and precedes(front[f], f, f + 1)):
                continue
            while pos[f] < len(lists[f]) and precedes(lists[f][pos[f]], f, f - 1) and precedes(lists[f][pos[f]], f, f + 1):
                t0 = int(lists[f][pos[f]])
                pos[f] += 1
                fired = True
                if t0 >= n_t - 1:
                    continue
                i, quiet = _scan(E, f, t0 + 1, +1, lambda i: i < n_t - 1, energy_tol, frame_thresh)
                i -= quiet
                if i - t0 <= min_note_len:
                    continue
                wipe(t0, i, f)
                onset_notes.append((t0, i, f))
        assert fired, "no column could fire: the schedule would deadlock"
    onset_notes.sort(key=lambda n: (-n[0], -n[2]))

    # ---------------- melodia loop ----------------
    mel = []
    if melodia:
        while True:
            cmax = [(float(E[:, f].max()), int(np.argmax(E[:, f]))) if n_t else (0.0, 0) for f in range(n_f)]
            if not any(v > frame_thresh for v, _ in cmax):
                break

            def beats(v, t, f, g):
                if g < 0 or g >= n_f:
                    return True
                vg, tg = cmax[g]
                return v > vg or (v == vg and (t < tg or (t == tg and f < g)))

            fired = False
            for f in range(n_f):
                v, tm = cmax[f]
                if not (v > frame_thresh and beats(v, tm, f, f - 1) and beats(v, tm, f, f + 1)):
                    continue
                while v > frame_thresh and beats(v, tm, f, f - 1) and beats(v, tm, f, f + 1):
                    fired = True
                    E[tm, f] = 0
                    i, quiet = tm + 1, 0
                    while i < n_t - 1 and quiet < energy_tol:
                        quiet = quiet + 1 if E[i, f] < frame_thresh else 0
                        wipe(i, i + 1, f)
                        i += 1
                    t_end = i - 1 - quiet
                    i, quiet = tm - 1, 0
                    while i > 0 and quiet < energy_tol:
                        quiet = quiet + 1 if E[i, f] < frame_thresh else 0
                        wipe(i, i + 1, f)
                        i -= 1
                    t_start = i + 1 + quiet
                    if t_end - t_start > min_note_len:
                        mel.append((v, tm, f, t_start, t_end))
                    v, tm = float(E[:, f].max()), int(np.argmax(E[:, f]))  # own column only: the neighbours keep their
                    # round-start maxima, which are upper bounds of their current ones
            assert fired
    mel.sort(key=lambda n: (-n[0], n[1], n[2]))
    return onset_notes + [(a, b, f) for _v, _t, f, a, b in mel]


def _smooth_field(rng, n_t, density, length):
    x = np.zeros((n_t, 88), np.float32)
    for _ in range(int(density * n_t)):
        f = rng.integers(0, 88)
        t0 = rng.integers(0, n_t)
        ln = int(rng.integers(3, length))
        amp = rng.uniform(0.2, 1.0)
        seg = amp * np.exp(-np.arange(ln) / rng.uniform(5, 60))
        x[t0 : t0 + ln, f] = np.maximum(x[t0 : t0 + ln, f], seg[: max(0, min(ln, n_t - t0))].astype(np.float32))
    return np.clip(x + 0.05 * rng.random((n_t, 88)).astype(np.float32), 0, 1)


@pytest.mark.parametrize("seed,n_t,density,quant", [(0, 300, 0.5, None), (1, 865, 0.3, None), (2, 400, 2.0, None),
                                                  (3, 200, 1.0, 8), (4, 173, 0.2, None), (5, 500, 0.8, 4), (6, 30, 1.0, None)])
def test_column_parallel_schedule_equals_sequential(seed, n_t, density, quant):
    rng = np.random.default_rng(seed)
    frames = _smooth_field(rng, n_t, density, 80)
    onsets = _smooth_field(rng, n_t, density * 0.7, 6)
    if quant:  # coarse values: many exact ties in the melodia arg-max and equal candidate times across columns
        frames = (np.round(frames * quant) / quant).astype(np.float32)
        onsets = (np.round(onsets * quant) / quant).astype(np.float32)
    for melodia in (True, False):
        exp = decode_ref.output_to_notes_polyphonic(frames.copy(), onsets.copy(), 0.5, 0.3, 11, True, None, None, melodia)
        inf = decode_ref.infer_onsets(onsets, frames)
        pk = decode_ref.strict_time_peaks(inf)
        with np.errstate(invalid="ignore"):
            cand = np.where(pk, inf, 0.0) >= 0.5
        got = parallel_decode(frames, cand, 0.3, 11, melodia=melodia)
        assert got == [(a, b, p - 21) for a, b, p, _amp in exp], f"seed {seed} melodia {melodia}"


def test_dense_same_time_candidates_chain():
    """All 88 columns strike at the same frames (the dense-chord stress case): the schedule degenerates to a chain in
    pitch order and must still terminate with the sequential result."""
    n_t = 120
    frames = np.zeros((n_t, 88), np.float32)
    onsets = np.zeros((n_t, 88), np.float32)
    for t0 in (5, 45, 85):
        frames[t0 : t0 + 30] = 0.8
        onsets[t0] = 0.9
    exp = decode_ref.output_to_notes_polyphonic(frames.copy(), onsets.copy(), 0.5, 0.3, 11, False, None, None, True)
    pk = decode_ref.strict_time_peaks(onsets.astype(np.float64))
    got = parallel_decode(frames, pk & (onsets >= 0.5), 0.3, 11)
    assert got == [(a, b, p - 21) for a, b, p, _amp in exp]


def test_block_maxima_track_the_column_argmax():
    """decode_seq_kernel keeps, per pitch column, the maxima of 256-frame blocks (value + lowest frame) so that a melodia
    iteration rescans only the blocks it changed and reduces T / 256 entries (csrc/decode.cu: refresh_column).  NumPy model of
    that bookkeeping: after arbitrary range zeroings the reduced (max, first index) equals np.max / np.argmax of the
    column, ties included (the reference takes the FIRST maximum, note_creation.py:449-452)."""
    rng = np.random.default_rng(3)
    BLK = 256
    for T in (1, 255, 256, 257, 865, 4000):
        col = np.round(rng.random(T) * 8) / 8  # coarse values: many exact ties
        nblk = (T + BLK - 1) // BLK
        bmax = np.full(nblk, -np.inf)
        barg = np.zeros(nblk, np.int64)

        def refresh(b_lo, b_hi):
            for b in range(b_lo, b_hi + 1):
                seg = col[b * BLK : min(T, (b + 1) * BLK)]
                bmax[b] = seg.max()
                barg[b] = b * BLK + int(np.argmax(seg))  # first maximum inside the block
            best_v, best_t = -np.inf, 2**31 - 1
            for b in range(nblk):  # ascending blocks, strict '>' keeps the lower block on ties
                if bmax[b] > best_v:
                    best_v, best_t = bmax[b], barg[b]
            return best_v, best_t

        v, t = refresh(0, nblk - 1)
        assert v == col.max() and t == int(np.argmax(col))
        for _ in range(60):
            tm = int(rng.integers(0, T))
            lo = max(0, tm - int(rng.integers(0, 40)))
            hi = min(T - 1, tm + int(rng.integers(0, 40)))
            col[lo : hi + 1] = 0.0
            v, t = refresh(lo // BLK, hi // BLK)
            assert v == col.max() and t == int(np.argmax(col)), (T, lo, hi)
