"""Golden vectors for lemo_b200/temp_prox/windows.py produced by EXECUTING the reference's own lines (build container only).

temp_prox/data_parser_slide.py cannot be imported here (it needs cv2 and PROX data at construction), so the window-building statements
(:200-210) are read from the reference file at run time, dedented, and executed against a stand-in `self` whose `img_paths` is a list
of frame indices; torch's DataLoader(batch_size, drop_last=True) then batches the result as main_slide.py:142-149 does.  No reference
source is stored in this repository: only the resulting integer tables (tests/golden/reference_golden_windows.npz)."""
import os, sys, textwrap, types
import numpy as np
import torch

REF = '/root/reference/temp_prox/data_parser_slide.py'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'reference_golden_windows.npz')


def reference_windows(n_frames, batch_size):
    lines = open(REF).read().split('\n')
    first = next(i for i, l in enumerate(lines) if 'slide_window_size = int(self.batch_size * 0.7)' in l)
    last = next(i for i, l in enumerate(lines) if i > first and 'self.seq_marker_mask_slide = np.append' in l and not l.strip().startswith('#'))
    src = textwrap.dedent('\n'.join(lines[first:last + 1]))
    me = types.SimpleNamespace(batch_size=batch_size, img_paths=list(range(n_frames)), seq_marker_mask=np.zeros((n_frames, 67)))
    exec(src, {'np': np, 'self': me, 'int': int, 'min': min, 'len': len, 'range': range})
    loader = torch.utils.data.DataLoader(me.img_paths_slide, batch_size=batch_size, shuffle=False, drop_last=True)
    return [b.numpy().astype(np.int64) for b in loader]


def main():
    gold = {}
    for n, B in ((100, 100), (101, 100), (170, 100), (171, 100), (350, 100), (1000, 100), (57, 10), (23, 10), (200, 30), (305, 100), (99, 100)):
        w = reference_windows(n, B)
        gold['n%d_B%d' % (n, B)] = np.stack(w, 0) if w else np.zeros((0, B), np.int64)
        print(n, B, len(w), [int(x[0]) for x in w][:6])
    np.savez_compressed(OUT, **gold)


if __name__ == '__main__':
    main()
