timeout 900 python -m pytest tests/test_fsk_gpu.py -x -q -m gpu -k "wpc or full_ctas or more_channels" 2>&1 | tail -2
python profiles/sweep.py 18 wpc:1024 wpc:2220 wpc:4096 2>&1 | tail -3
timeout 800 ncu --set full --clock-control none --import-source on -k regex:fsk_wpc -s 2 -c 1 -o gpurun_out/r02_wpc_g7 python profiles/sweep.py 16 wpc:1036 2>&1 | tail -1
