"""Host-side mirror of the reference's `evaluation` package (evaluation/tools.py, evaluation/eval_ullava.py:validate)
with the per-sentence metrics computed on the device."""
