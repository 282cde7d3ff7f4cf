import ctypes as C, os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from oracle import pyoracle as po
from radiosonde_auto_rx_b200 import siggen, fsk
cfg = po.RS41
iq = siggen.fsk_iq(23, 40_000, ebno_db=11.0, ppm=-30.0)
ref, _, recs = po.Oracle(cfg).run_stream(iq)
b = fsk.FskBatch(cfg.Fs, cfg.Rs, cfg.M, cfg.P, cfg.Nsym, 1000, cfg.fs_tx, cfg.lo, cfg.hi, 1, nchan=1)
b.enable_stats(True)
pos = 0; k = 0
x = (iq.astype(np.float32) / np.float32(1000))
while True:
    nin = int(b.nin()[0])
    if pos + nin > len(iq): break
    out, fr = b.push(x[None, pos:pos + nin].reshape(1, -1), "f32", soft=True, max_frames=1)
    st = b.stats(0)
    print(k, fr[0], "ppm", st.ppm, float(recs[k]["ppm"]), "nrt", st.norm_rx_timing, float(recs[k]["norm_rx_timing"]))
    pos += nin; k += 1
