"""Host-side breakdown of predict_batch (the Python drop-in of the batch path): where the wall time goes."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np


def main() -> None:
    import bench
    from basic_pitch_b200 import ICASSP_2022_MODEL_PATH
    from basic_pitch_b200 import note_creation as infer
    from basic_pitch_b200.inference import Model, predict_batch

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
    model = Model(ICASSP_2022_MODEL_PATH)
    clips = bench.make_clips(n, seed0=3)
    for rep in range(3):
        t0 = time.perf_counter()
        outs, arrs, frames = model.transcribe_arrays(clips, min_note_len=11, split_notes=False)
        t1 = time.perf_counter()
        ev = infer.note_events_batch(arrs, n)
        t2 = time.perf_counter()
        midis = [infer.LazyPrettyMIDI(e) for e in ev]
        t3 = time.perf_counter()
        full = [e.to_list() for e in ev]
        t4 = time.perf_counter()
        for m in midis[:50]:
            m.instruments
        t5 = time.perf_counter()
        print(f"rep {rep}: transcribe_arrays {1e3*(t1-t0):.1f} ms | lazy events {1e3*(t2-t1):.1f} | lazy midi {1e3*(t3-t2):.1f} | "
              f"materialise events {1e3*(t4-t3):.1f} | 50 midi objects {1e3*(t5-t4):.1f}", flush=True)
        del outs, arrs, ev, midis, full
    t0 = time.perf_counter()
    r = predict_batch(clips, model)
    print(f"predict_batch (lazy) {1e3*(time.perf_counter()-t0):.1f} ms")
    del r


if __name__ == "__main__":
    main()
