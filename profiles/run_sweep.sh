timeout 600 python -m pytest tests/test_fsk_gpu.py -x -q -m gpu -k "duo" 2>&1 | tail -3
python profiles/duo_timing.py 256 262144 2>&1 | tail -8
python profiles/sweep.py 20 duo:256 duo:296 duo:148 2>&1 | tail -3
