import json, os, subprocess, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from oracle import pyoracle as po
from radiosonde_auto_rx_b200 import siggen
BIN="/root/repo/radiosonde_auto_rx_b200/bin"
args=["--cs16", "-b", "-20000", "-u", "20000", "-s", "--stats=5", "2", "48000", "4800", "-", "-"]
iq = siggen.fsk_iq(17, 150_000, M=2, ebno_db=11.0, ppm=20.0)
def run(exe): 
    r=subprocess.run([exe]+args,input=iq.tobytes(),capture_output=True); return r.stdout,r.stderr
ro,re_=run(os.path.join(po.REF_DIR,"fsk_demod")); go,ge=run(os.path.join(BIN,"fsk_demod"))
rl=re_.decode().splitlines(); gl=ge.decode().splitlines()
print(len(rl),len(gl), ro==go)
for i,(a,b) in enumerate(zip(rl,gl)):
    if a!=b:
        k=[q for q in range(min(len(a),len(b))) if a[q]!=b[q]][0]
        print("line",i,"col",k); print("REF:",a[:110]); print("GOT:",b[:110]); print("REF@:",a[max(0,k-80):k+40]); print("GOT@:",b[max(0,k-80):k+40])
        break
