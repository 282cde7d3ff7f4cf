timeout 900 python -m pytest tests/test_fsk_gpu.py tests/test_golden.py tests/test_fsk_capi.py tests/test_cli_gpu.py -x -q -m gpu 2>&1 | tail -2
python profiles/sweep.py 20 pipe:256 pipe:128 pipe:148 2>&1 | tail -3
