#!/usr/bin/env python
"""Debug aid: push one seeded batch through the duo kernel and through the serial kernel, print the first mismatching soft bit per channel."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiosonde_auto_rx_b200 import fsk, siggen
nchan, nsamp = 5, 150_000
iq = siggen.fsk_batch(104, nchan, nsamp)
iq[2] = siggen.noise_iq(104 + 99, nsamp)
def run(mode):
    os.environ["B200FSK_KERNEL"] = mode
    b = fsk.FskBatch(48000, 4800, 2, 5, 300, 1000, 5000, -5000, 5000, 1, nchan=nchan)
    out, frames = b.push(iq, "cs16", soft=True)
    return out.view(np.uint32), frames, b.last_kernel()
ref, fr, _ = run("serial")
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out, f2, kn = run("duo")
    msg = []
    for c in range(nchan):
        n = fr[c] * 300
        bad = np.nonzero(out[c, :n] != ref[c, :n])[0]
        msg.append(f"ch{c}:" + ("ok" if not len(bad) and f2[c] == fr[c] else f"{len(bad)} bad, first bit {bad[0] if len(bad) else -1} (frame {bad[0] // 300 if len(bad) else -1}) frames {f2[c]}/{fr[c]}"))
    print(kn, " ".join(msg))
