timeout 900 python -m pytest tests/test_fsk_gpu.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -2
python profiles/sweep.py 20 pipe:256 2>&1 | tail -1
python profiles/sweep.py 18 pipe:148 2>&1 | tail -1
python profiles/pipe_timing.py 256 262144 2>&1 | tail -5
