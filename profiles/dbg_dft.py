import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from radiosonde_auto_rx_b200 import dft, siggen
from oracle import pyoracle as po
nsamp = 1 << 18
for c, kind in ((9, "MRZ"), (7, "MEISEI")):
    iq = siggen.fm_header_iq(9000 + c, nsamp, kind, snr_db=(10, 14, 20)[c % 3], invert=(c % 5 == 4))
    rc, dets, blks = po.dft_oracle_run(iq, 48000, 15.0, True, True, want_blocks=True)
    g_rc, g = dft.DftDetectBatch(48000, 15.0, True, nchan=1).run(iq[None], cont=True)[0]
    print(kind, "oracle", [(d[0], round(d[1], 4), d[2]) for d in dets])
    print(kind, "gpu   ", [(d.type, round(d.score, 4), d.pos) for d in g])
    j = {"MRZ": 9, "MEISEI": 7}[kind]
    print(kind, "oracle blocks hdr", j, [(b[1], round(b[2], 4), b[3]) for b in blks if b[0] == j])
