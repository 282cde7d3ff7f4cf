timeout 900 python -m pytest tests/test_fsk_gpu.py -x -q -m gpu 2>&1 | tail -2
echo "== prof"; timeout 300 python profiles/duo_timing.py 256 262144 2>&1 | tail -8
timeout 300 python profiles/sweep.py 20 duo:256 duo:296 pipe:128 pipe:256 2>&1 | tail -4
B200FSK_LIB=$PWD/radiosonde_auto_rx_b200/libb200fsk_dm.so timeout 300 python profiles/sweep.py 20 duo:256 2>&1 | tail -1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:fsk_duo -s 2 -c 1 -o gpurun_out/r02_duo python profiles/sweep.py 18 duo:256 2>&1 | tail -1
