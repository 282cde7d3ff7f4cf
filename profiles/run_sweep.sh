B200FSK_DUO_G=1 timeout 300 python profiles/sweep.py 20 duo:128 duo:148 2>&1 | tail -2
timeout 300 python profiles/sweep.py 20 pipe:128 pipe:148 duo:256 2>&1 | tail -3
B200FSK_DUO_G=1 timeout 300 python profiles/duo_timing.py 148 262144 2>&1 | tail -8
B200FSK_DUO_G=1 timeout 300 python -m pytest tests/test_fsk_gpu.py -q -m gpu -k duo 2>&1 | tail -2
