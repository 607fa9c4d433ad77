"""Re-export of tools/synth.py (the seeded SURVEY.md 8(d) input generators) for the oracle-side
scripts and tests.  The generators themselves are plain torch and live outside oracle/ so that
bench.py's product arm never imports anything from this package."""
from tools.synth import synth_gt, synth_rois, synth_rpn  # noqa: F401
