"""Re-export of the package's synthetic-data generators for the tests (the generators are data only and live in
lemo_b200/synth.py so that the product arm of bench.py never imports oracle/)."""
from lemo_b200.synth import *          # noqa: F401,F403
from lemo_b200.synth import _rodrigues_np, GOLDEN   # noqa: F401
