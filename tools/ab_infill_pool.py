"""GPU measurement: AE fine-tune + infill throughput of InfillPool against the number of clips in flight (one stage + stream per clip)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lemo_b200 import _lib, synth                      # noqa: E402
_lib.build()
from lemo_b200.infill import InfillPool, body_repr, load_infill_prior, load_infill_stats   # noqa: E402

dev = torch.device('cuda', 0)
body68, con68 = synth.synth_marker_clip(5, T=120)
st64 = load_infill_stats()
clip, rot0 = body_repr(torch.from_numpy(body68).to(dev), torch.from_numpy(con68).to(dev), stats=st64, device=dev)
c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for S in [int(x) for x in sys.argv[1:]] or [4, 8, 12, 16, 24]:
    pool = InfillPool(load_infill_prior(), n_streams=S, device=dev, stats=st64)
    pool.run_many([clip] * S, [rot0] * S)
    torch.cuda.synchronize(dev)
    c0.record()
    pool.run_many([clip] * S, [rot0] * S)
    c1.record()
    torch.cuda.synchronize(dev)
    print('clips in flight %2d: %.1f ms per clip (%.0f ms for the batch)' % (S, c0.elapsed_time(c1) / S, c0.elapsed_time(c1)), flush=True)
    del pool
