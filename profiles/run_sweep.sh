timeout 600 python -m pytest tests/test_fsk_gpu.py -x -q -k "wpc or more_channels" 2>&1 | tail -3
python profiles/sweep.py 18 wpc:1024 wpc:2048 wpc:2368 wpc:4096 2>&1 | tail -4
