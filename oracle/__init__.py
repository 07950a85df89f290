"""ORACLE — test infrastructure only.

CPU restatements of the reference hot path used to CHECK the CUDA path:
  model_ref.py   the deployed graph (CQT -> log-norm -> BN -> harmonic stack -> CNN), torch-CPU f32/f64
  decode_ref.py  note decode (note_creation.py) in NumPy
  host_ref.py    windowing / unwrap index arithmetic (inference.py)
  ref_shims/     stand-ins for librosa / pretty_midi / mir_eval / resampy / onnxruntime so that the
                 UNMODIFIED reference modules under /root/reference import in the build container
                 (used only by make_golden.py to generate tests/golden/*; never on the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything from here.  The product package never does.
"""
