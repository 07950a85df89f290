"""GPU tests for the BASELINE.json configurations that round 1 left unpinned (run with `-m gpu`):

  configs[1]  one 180 s clip end to end: posteriorgrams of all 110 windows vs the oracle, decode bit-identical on the
              GPU posteriorgrams (incl. pitch bends), end-to-end event agreement vs the CPU oracle
              (reference analogue: tests/test_inference.py:43-76)
  configs[2]  1024 x 2 s windows, HCQT + CNN only: sampled windows vs the oracle
  configs[3]  a multi-chunk batch of distinct 10 s clips through the batched entry point: sampled clips vs the oracle
  configs[4]  dense polyphony (88-voice chords) at batch scale: decode bit-exact incl. pitch bends

All calls go through the Python mirror -> ctypes -> C ABI -> CUDA kernels; the oracle is only the checker.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POST_TOL = 1e-4  # vs the fp32 oracle on identical 22 050 Hz input (north-star bar: 1e-3)


@pytest.fixture(scope="module")
def model():
    from basic_pitch_b200 import ICASSP_2022_MODEL_PATH
    from basic_pitch_b200.inference import Model

    return Model(ICASSP_2022_MODEL_PATH)


def test_config1_180s_clip_end_to_end(model):
    from basic_pitch_b200 import synth
    from bench import host_threads, parity_report

    clip = synth.random_notes_clip(180.0, seed=1)
    rep = parity_report(model, [clip], [0], host_threads())
    print("config[1] parity:", rep)
    for k, v in rep["post_max_abs"].items():
        assert v < POST_TOL, f"{k}: max-abs {v:.3e} over the 110 windows of the 180 s clip"
    assert rep["decode_bit_identical_on_gpu_posteriorgrams"]
    assert rep["e2e_events"]["gpu"] > 500
    assert rep["e2e_event_agreement"] >= 0.99, rep["e2e_events"]


def test_config2_1024_windows_sampled_vs_oracle(model, weights_np):
    from basic_pitch_b200 import synth
    from oracle import model_ref

    x = synth.window_batch(1024, seed=2)
    # make every window distinct: the generator tiles 64 windows, so scale each tile copy differently
    gain = (0.5 + 0.5 * np.arange(1024) / 1023.0).astype(np.float32)[:, None]
    x = np.ascontiguousarray(x * gain)
    got = model.predict(x[:, :, None])
    idx = np.unique(np.concatenate([np.arange(0, 1024, 17), [215, 216, 217, 431, 432, 1023]]))  # incl. chunk borders
    assert len(idx) >= 64
    ref = model_ref.forward_batched(x[idx], weights_np, batch=33)
    for k in ("note", "onset", "contour"):
        assert got[k].shape[0] == 1024
        err = float(np.abs(got[k][idx] - ref[k]).max())
        assert err < POST_TOL, f"{k}: max-abs {err:.3e} on {len(idx)} sampled windows of the 1024-window batch"


def test_config3_multichunk_batch_sampled_vs_oracle(model):
    from bench import host_threads, make_clips, parity_report

    clips = make_clips(260, seed0=3)  # 1820 windows = 9 internal chunks, 3 sub-batches of the host entry point
    rep = parity_report(model, clips, [0, 37, 129, 200, 259], host_threads())
    print("config[3] parity:", rep)
    for k, v in rep["post_max_abs"].items():
        assert v < POST_TOL, f"{k}: max-abs {v:.3e}"
    assert rep["decode_bit_identical_on_gpu_posteriorgrams"]
    assert rep["e2e_event_agreement"] >= 0.97, rep["e2e_events"]
    # the batch result must not depend on the batch composition: clip 129 alone == clip 129 inside the batch
    outs_b, res_b, _ = model.transcribe_arrays(clips)
    outs_1, res_1, _ = model.transcribe_arrays([clips[129]])
    for k in ("note", "onset", "contour"):
        assert np.array_equal(outs_b[129][k], outs_1[0][k])
    for k in ("start", "end", "pitch", "amp", "bends"):
        assert np.array_equal(res_b[129][k], res_1[0][k])


def test_config4_dense_polyphony_batch_decode_bit_exact(model):
    from basic_pitch_b200 import synth
    from oracle import decode_ref

    clips = [synth.dense_chords_clip(10.0, seed=7 + i) for i in range(12)]
    outs, res, _frames = model.transcribe_arrays(clips)
    total = 0
    for i in range(len(clips)):
        with np.errstate(all="ignore"):
            wb, _ = decode_ref.model_output_to_note_events({k: np.array(v) for k, v in outs[i].items()}, 0.5, 0.3)
        r = res[i]
        got = [(int(a), int(b), int(p), np.float32(x).tobytes(), [int(v) for v in r["bends"][r["bend_off"][j] : r["bend_off"][j + 1]]])
               for j, (a, b, p, x) in enumerate(zip(r["start"], r["end"], r["pitch"], r["amp"]))]
        exp = [(int(a), int(b), int(p), np.float32(x).tobytes(), [int(v) for v in bd]) for a, b, p, x, bd in wb]
        assert got == exp, f"dense clip {i}: {len(got)} vs {len(exp)} notes"
        total += len(got)
    assert total > 12 * 100
